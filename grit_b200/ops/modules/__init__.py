from .ms_deform_attn import MSDeformAttn, hoisted_value_proj  # noqa: F401
from .decoder_layer import (DeformableTransformerDecoderLayer, GraphedDecoder, graphed_training_decoder, extract_region_features,  # noqa: F401
                            run_decoder)
