"""Autograd boundary of multi-scale deformable attention -- drop-in for the reference's
``models/ops/functions/ms_deform_attn_func.py``.

``MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index, sampling_locations,
attention_weights, im2col_step)`` keeps the reference signature, return shape ``(N, Lq, M*D)`` and backward
tuple ``(grad_value, None, None, grad_sampling_loc, grad_attn_weight, None)`` (reference :21-38).  Forward and
backward call the sm_100a kernels through the C ABI (grit_b200/_lib.py -> include/msda.h); there is no other
implementation behind this class.

``im2col_step`` is accepted for signature compatibility.  The reference uses it to cut the batch into chunks with
one launch each and requires ``batch % min(batch, im2col_step) == 0`` (ms_deform_attn_cuda.cu:50-72); the B200
kernels cover the whole batch in one launch, so the value is only validated to be a positive integer.

bf16: ``value`` may be bf16 (output and grad_value are then bf16, accumulation is fp32).  ``sampling_locations`` /
``attention_weights`` are used in fp32; bf16 ones are up-cast here (bf16 cannot resolve pixel coordinates).
"""
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _lib

_deterministic = os.environ.get("GRIT_B200_DETERMINISTIC", "0") not in ("", "0", "false", "False")


def set_deterministic(enabled: bool) -> bool:
    """Select the bit-reproducible backward (MSDA_FLAG_DETERMINISTIC: grad_value accumulated as power-of-two scaled
    int64 fixed point with integer atomics; fp32 / bf16 values).  Also switched on by
    ``torch.use_deterministic_algorithms(True)`` or ``GRIT_B200_DETERMINISTIC=1``.  Returns the previous setting.
    The reference backward (float atomicAdd, ms_deform_im2col_cuda.cuh:87-159) has no such mode."""
    global _deterministic
    prev, _deterministic = _deterministic, bool(enabled)
    return prev


def _backward_flags(dtype) -> int:
    if dtype != torch.float64 and (_deterministic or torch.are_deterministic_algorithms_enabled()):
        return _lib.FLAG_DETERMINISTIC
    return 0


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        if int(im2col_step) <= 0:
            raise RuntimeError(f"im2col_step must be positive, got {im2col_step}")
        ctx.im2col_step = im2col_step
        ctx.loc_dtype = sampling_locations.dtype
        ctx.attn_dtype = attention_weights.dtype
        if value.dtype == torch.bfloat16:
            sampling_locations = sampling_locations.float()
            attention_weights = attention_weights.float()
        output = _lib.forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = _lib.backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), _backward_flags(value.dtype))
        if grad_sampling_loc.dtype != ctx.loc_dtype:
            grad_sampling_loc = grad_sampling_loc.to(ctx.loc_dtype)
        if grad_attn_weight.dtype != ctx.attn_dtype:
            grad_attn_weight = grad_attn_weight.to(ctx.attn_dtype)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Debug/test helper with the reference's name and signature (reference :41-61): the same function written in
    plain differentiable PyTorch (explicit four-tap gathers instead of ``F.grid_sample``).  Works on any device and
    dtype.  It is NOT used by ``MSDeformAttnFunction`` or ``MSDeformAttn``; it exists because the reference exports
    it and ``models/ops/test.py`` imports it to check the CUDA path.
    """
    n, _, m, d = value.shape
    _, lq, _, n_levels, n_points, _ = sampling_locations.shape
    out = value.new_zeros((n, lq, m, d))
    batch_ix = torch.arange(n, device=value.device).view(n, 1, 1, 1)
    head_ix = torch.arange(m, device=value.device).view(1, 1, m, 1)
    start = 0
    for lvl in range(n_levels):
        h, w = int(value_spatial_shapes[lvl][0]), int(value_spatial_shapes[lvl][1])
        x = sampling_locations[:, :, :, lvl, :, 0] * w - 0.5  # (N, Lq, M, P)
        y = sampling_locations[:, :, :, lvl, :, 1] * h - 0.5
        inside = (y > -1) & (x > -1) & (y < h) & (x < w)
        r0 = torch.floor(y.detach())
        c0 = torch.floor(x.detach())
        ly, lx = y - r0, x - c0
        for dr, dc, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
            r, c = r0 + dr, c0 + dc
            ok = inside & (r >= 0) & (r <= h - 1) & (c >= 0) & (c <= w - 1)
            pix = start + (r.clamp(0, h - 1) * w + c.clamp(0, w - 1)).long()
            pix = torch.where(ok, pix, torch.zeros_like(pix))  # NaN coordinates never index out of range
            tap = value[batch_ix, pix, head_ix]  # (N, Lq, M, P, D)
            coef = torch.where(ok, wgt, torch.zeros_like(wgt)) * attention_weights[:, :, :, lvl]
            out = out + (tap * coef.unsqueeze(-1)).sum(dim=3)
        start += h * w
    return out.reshape(n, lq, m * d)
