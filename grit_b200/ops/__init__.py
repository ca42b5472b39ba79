"""Mirror of the reference's ``models/ops`` package layout (functions/, modules/)."""
from .functions import MSDeformAttnFunction, ms_deform_attn_core_pytorch, set_deterministic  # noqa: F401
from .modules import MSDeformAttn, hoisted_value_proj  # noqa: F401
