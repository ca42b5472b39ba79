from .ms_deform_attn_func import (AddDropoutLayerNormFunction, MSDeformAttnFunction,  # noqa: F401
                                  MSDeformAttnFusedFunction, PackLevelsFunction, PackLevelsGroupNormFunction,
                                  add_dropout_layer_norm, ms_deform_attn_core_pytorch, pack_levels,
                                  pack_levels_groupnorm, set_deterministic)
