"""Mirror of the reference's ``models/ops`` package layout (functions/, modules/)."""
from .functions import (MSDeformAttnFunction, ms_deform_attn_core_pytorch, pack_levels,  # noqa: F401
                        set_deterministic)
from .modules import MSDeformAttn, hoisted_value_proj  # noqa: F401
