from .ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction, PackLevelsFunction,  # noqa: F401
                                  ms_deform_attn_core_pytorch, pack_levels, set_deterministic)
