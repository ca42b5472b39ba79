"""grit_b200 -- B200 (sm_100a) multi-scale deformable attention, drop-in for davidnvq/grit's ``models/ops``.

    from grit_b200 import MSDeformAttn, MSDeformAttnFunction          # same API as models.ops.{modules,functions}
    grit_b200.install_as_reference_ops()                               # make the reference's own imports resolve here

The package holds only the hot path: CUDA kernels + C ABI (csrc/, include/msda.h), the ctypes binding (_lib.py) and the
host-side mirror of the reference interface (ops/).
"""
import sys

from .ops.functions import (MSDeformAttnFunction, add_dropout_layer_norm, ms_deform_attn_core_pytorch,  # noqa: F401
                            pack_levels, pack_levels_groupnorm, set_deterministic)
from .ops.modules import (DeformableTransformerDecoderLayer, GraphedDecoder, MSDeformAttn,  # noqa: F401
                          extract_region_features, graphed_training_decoder, hoisted_value_proj, run_decoder)

__all__ = ["MSDeformAttn", "MSDeformAttnFunction", "ms_deform_attn_core_pytorch", "install_as_reference_ops",
           "set_deterministic", "hoisted_value_proj", "pack_levels", "pack_levels_groupnorm", "add_dropout_layer_norm",
           "DeformableTransformerDecoderLayer", "graphed_training_decoder", "run_decoder", "GraphedDecoder", "extract_region_features"]


def install_as_reference_ops(alias_models_ops: bool = True):
    """Register the compat modules under the names the reference imports.

    * ``MultiScaleDeformableAttention`` -> grit_b200.MultiScaleDeformableAttention (the pybind replacement)
    * ``models.ops``, ``models.ops.functions``, ``models.ops.modules`` -> grit_b200.ops.* when those names are not
      already importable (GRIT does ``from models.ops.modules import MSDeformAttn``, models/detection/det_module.py).
    """
    from . import MultiScaleDeformableAttention as msda_mod
    from . import ops

    sys.modules["MultiScaleDeformableAttention"] = msda_mod
    if alias_models_ops:
        sys.modules.setdefault("models.ops", ops)
        sys.modules.setdefault("models.ops.functions", ops.functions)
        sys.modules.setdefault("models.ops.modules", ops.modules)
    return msda_mod
