from .ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction,  # noqa: F401
                                  ms_deform_attn_core_pytorch, set_deterministic)
