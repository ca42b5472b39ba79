from .ms_deform_attn import MSDeformAttn, hoisted_value_proj  # noqa: F401
