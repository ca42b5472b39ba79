"""Stand-in for the reference's compiled extension module ``MultiScaleDeformableAttention``
(pybind surface: models/ops/src/vision.cpp:13-16; signatures: models/ops/src/ms_deform_attn.h:20-61).

``grit_b200.install_as_reference_ops()`` registers this module under that name so the reference's own
``models/ops/functions/ms_deform_attn_func.py`` (``import MultiScaleDeformableAttention as MSDA``) and
``models/ops/test.py`` run unchanged on top of the sm_100a kernels.
"""
from . import _lib


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    return _lib.forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    return list(_lib.backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output))
