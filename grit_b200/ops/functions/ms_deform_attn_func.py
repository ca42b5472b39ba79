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


_warned_fp64_det = False


def _backward_flags(dtype) -> int:
    global _warned_fp64_det
    if _deterministic or torch.are_deterministic_algorithms_enabled():
        if dtype != torch.float64:
            return _lib.FLAG_DETERMINISTIC
        if not _warned_fp64_det:  # int64 fixed point cannot carry fp64 precision: say so instead of silently ignoring it
            _warned_fp64_det = True
            import warnings
            warnings.warn("grit_b200: the deterministic backward covers fp32 / bf16; fp64 grad_value is accumulated "
                          "with floating-point atomics and is not bit-reproducible run to run")
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


class MSDeformAttnFusedFunction(Function):
    """The core op with the module's pre-op arithmetic inside the kernels: takes the RAW outputs of
    the ``sampling_offsets`` / ``attention_weights`` Linears plus the reference points; softmax, ``offsets/normaliser +
    reference`` and the padding-mask fill (reference modules/ms_deform_attn.py:96-111) never touch HBM as separate
    passes.  ``value`` is the value_proj output (N, S, M, D).  When ``padding_mask`` is given its masked rows are zeroed
    IN PLACE (and the same rows of grad_value in backward); the tensor's autograd version counter is bumped, so any
    other autograd node that saved ``value`` raises instead of silently reading modified data.  Pass a mask only for a
    ``value`` you own -- ``MSDeformAttn`` does so for its own ``value_proj`` output and masks a caller-supplied
    ``value=`` out of place.  Not part of the reference API: ``MSDeformAttn`` uses it when ``module.fused`` is on and
    ``_lib.fused_supported`` says yes."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_offsets, attn_logits,
                reference_points, padding_mask, valid_ratios=None, validated=False):
        # validated: the caller has just run _lib.fused_supported on exactly these tensors (MSDeformAttn does), so the
        # ~40 attribute checks are not repeated here and in backward -- they are a third of the op's host time at GRIT's
        # decoder sizes
        if padding_mask is not None:
            # in-place write: `value` is the value_proj output, whose producer (addmm) does not need its own output
            # in backward; the matching rows of grad_value are zeroed below
            _lib.mask_rows_(value, padding_mask)
            torch.autograd.graph.increment_version(value)
        ctx.has_mask = padding_mask is not None
        ctx.has_vr = valid_ratios is not None
        sampling_offsets = sampling_offsets.contiguous()
        attn_logits = attn_logits.contiguous()
        reference_points = reference_points.contiguous()
        if valid_ratios is not None:
            valid_ratios = valid_ratios.contiguous()
        output = _lib.fused_forward(value, value_spatial_shapes, value_level_start_index, sampling_offsets,
                                    attn_logits, reference_points, valid_ratios, _validated=validated)
        saved = [value, value_spatial_shapes, value_level_start_index, sampling_offsets, attn_logits, reference_points]
        if padding_mask is not None:
            saved.append(padding_mask)
        if valid_ratios is not None:
            saved.append(valid_ratios)
        ctx.save_for_backward(*saved)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, offsets, logits, ref = ctx.saved_tensors[:6]
        vr = ctx.saved_tensors[-1] if ctx.has_vr else None
        # the saved tensors passed the checks in forward; only grad_output is new
        grad_value, grad_offs, grad_logits = _lib.fused_backward(value, shapes, lsi, offsets, logits, ref,
                                                                 grad_output.contiguous(), _backward_flags(value.dtype),
                                                                 valid_ratios=vr, _validated=True)
        if ctx.has_mask:
            _lib.mask_rows_(grad_value, ctx.saved_tensors[6])
        grad_ref = None
        if ctx.needs_input_grad[5]:
            n_points = offsets.shape[4]
            # per-level reference points as the kernels used them: (N, Lq, L, 2|4)
            if vr is None:
                ref_l = ref
            elif ref.shape[-1] == 2:
                ref_l = ref[:, :, None] * vr[:, None]
            else:
                ref_l = ref[:, :, None] * torch.cat([vr, vr], -1)[:, None]
            if ref.shape[-1] == 2:  # loc = ref + off / (W, H)  =>  d loc/d ref = 1, grad_loc = grad_off * (W, H)
                normalizer = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(grad_offs.dtype)
                grad_ref_l = (grad_offs * normalizer[None, None, None, :, None, :]).sum(dim=(2, 4))
            else:  # loc = ref.xy + off / P * ref.wh * 0.5
                scale = ref_l[:, :, None, :, None, 2:] * (0.5 / n_points)
                # (a zero-size reference box collapses every sample onto its centre and carries no offset gradient to
                #  recover grad_loc from; its reference-point gradient is reported as 0 instead of NaN)
                grad_loc = torch.where(scale != 0, grad_offs / torch.where(scale != 0, scale, torch.ones_like(scale)),
                                       torch.zeros_like(grad_offs))
                grad_xy = grad_loc.sum(dim=(2, 4))
                grad_wh = (grad_loc * offsets * (0.5 / n_points)).sum(dim=(2, 4))
                grad_ref_l = torch.cat([grad_xy, grad_wh], -1)
            if vr is None:
                grad_ref = grad_ref_l
            elif ref.shape[-1] == 2:  # ref_l = ref * vr[l]  =>  grad_ref = sum_l grad_ref_l * vr[l]
                grad_ref = (grad_ref_l * vr[:, None]).sum(2)
            else:
                grad_ref = (grad_ref_l * torch.cat([vr, vr], -1)[:, None]).sum(2)
        return grad_value, None, None, grad_offs, grad_logits, grad_ref, None, None, None


class AddDropoutLayerNormFunction(Function):
    """``LayerNorm(x + dropout(z))`` in one kernel each way (include/msda.h: msda_add_dropout_ln_*) --
    the epilogue that follows every attention / FFN block of the reference's DeformableTransformerDecoderLayer
    (models/detection/det_module.py:316-318, 331-333, 337-339).  ``keep`` is the dropout mask (bool, same shape) or
    None; ``keep_scale`` = 1/(1-p)."""

    @staticmethod
    def forward(ctx, x, z, keep, keep_scale, weight, bias, eps):
        need_grad = any(ctx.needs_input_grad[i] for i in (0, 1, 4, 5))
        y, h, mean, rstd = _lib.add_dropout_ln_forward(x, z, keep, keep_scale, weight, bias, eps, need_grad)
        ctx.keep_scale = keep_scale
        ctx.has_keep = keep is not None
        if need_grad:
            ctx.save_for_backward(*([h, mean, rstd, weight] + ([keep] if keep is not None else [])))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        h, mean, rstd, weight = ctx.saved_tensors[:4]
        keep = ctx.saved_tensors[4] if ctx.has_keep else None
        grad_x, grad_z, grad_w, grad_b = _lib.add_dropout_ln_backward(grad_y.contiguous(), h, mean, rstd, keep,
                                                                      ctx.keep_scale, weight)
        return grad_x, grad_z, None, None, grad_w, grad_b, None


def add_dropout_layer_norm(x, z, norm, p=0.0, training=False):
    """``norm(x + F.dropout(z, p, training))`` for an ``nn.LayerNorm`` ``norm`` -- one fused kernel when the tensors
    qualify (CUDA fp32, contiguous, 16-byte aligned, channels in {128, 256, 384, 512}, affine LayerNorm over the last
    dimension), the plain PyTorch composition otherwise.  In training the dropout mask is drawn from torch's own
    generator exactly as ``nn.Dropout`` would draw it for a tensor of this shape (``F.dropout`` on ones), so a seeded
    run reproduces the reference's mask."""
    import torch.nn.functional as F
    use_dropout = training and p > 0.0
    if norm.weight is not None and norm.bias is not None and tuple(norm.normalized_shape) == (x.shape[-1],):
        xc, zc = x.contiguous(), z.contiguous()
        keep = None
        if use_dropout:
            # the mask nn.Dropout would draw for this tensor: F.dropout on CUDA IS native_dropout, which returns its mask,
            # so one launch yields it (ones_like + dropout + compare were three)
            keep = torch.native_dropout(zc, p, True)[1]
        if _lib.add_dropout_ln_supported(xc, zc, norm.weight, norm.bias, keep):
            scale = 1.0 / (1.0 - p) if use_dropout else 1.0
            return AddDropoutLayerNormFunction.apply(xc, zc, keep, scale, norm.weight, norm.bias, norm.eps)
        if use_dropout:  # same mask, unfused
            return norm(x + z * keep.to(z.dtype) * (1.0 / (1.0 - p)))
    return norm(x + F.dropout(z, p, training))


class PackLevelsFunction(Function):
    """Multi-level NCHW feature maps -> the (N, S, C) memory the op reads, in one tiled-transpose launch
    (SURVEY.md 8f-3).  Same result as the reference's ``torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)``
    (models/detection/det_module.py:146-155); the backward is the adjoint scatter back to per-level NCHW gradients."""

    @staticmethod
    def forward(ctx, *levels):
        levels = [t.contiguous() for t in levels]
        ctx.shapes = [tuple(t.shape) for t in levels]
        return _lib.pack_levels(levels)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_memory):
        grads = [torch.empty(shape, dtype=grad_memory.dtype, device=grad_memory.device) for shape in ctx.shapes]
        _lib.pack_levels(grads, grad_memory.contiguous(), unpack=True)
        return tuple(grads)


class PackLevelsGroupNormFunction(Function):
    """GroupNorm + re-layout in one pass (second half): the conv outputs of Detector.input_proj
    (models/detection/detector.py:39-44, 64) go straight to the (N, S, C) memory the op reads, normalised with each
    level's GroupNorm affine, optionally as bf16 -- instead of GroupNorm writing N*C*H*W and prepare_od_inputs
    (det_module.py:146-155) reading and writing it again.  Backward: the packing's adjoint (msda_pack_levels, unpack)
    followed by torch's own GroupNorm backward on the statistics saved here."""

    @staticmethod
    def forward(ctx, num_groups, eps, out_dtype, n_levels, *tensors):
        levels = [t.contiguous() for t in tensors[:n_levels]]
        weights = [t.contiguous() for t in tensors[n_levels:2 * n_levels]]
        biases = [t.contiguous() for t in tensors[2 * n_levels:]]
        memory, stats = _lib.pack_levels_groupnorm(levels, weights, biases, num_groups, eps, out_dtype)
        ctx.num_groups, ctx.n_levels = num_groups, n_levels
        ctx.save_for_backward(stats, *levels, *weights)
        return memory

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_memory):
        stats = ctx.saved_tensors[0]
        levels = ctx.saved_tensors[1:1 + ctx.n_levels]
        weights = ctx.saved_tensors[1 + ctx.n_levels:]
        grads = [torch.empty(t.shape, dtype=torch.float32, device=t.device) for t in levels]
        _lib.pack_levels(grads, grad_memory.float().contiguous(), unpack=True)
        g_levels, g_weights, g_biases = [], [], []
        for l, (x, w, dy) in enumerate(zip(levels, weights, grads)):
            n, c, h, wd = x.shape
            mean, rstd = stats[l, :, :, 0].contiguous(), stats[l, :, :, 1].contiguous()
            gx, gw, gb = torch.ops.aten.native_group_norm_backward(dy, x, mean, rstd, w, n, c, h * wd, ctx.num_groups,
                                                                   [True, True, True])
            g_levels.append(gx), g_weights.append(gw), g_biases.append(gb)
        return (None, None, None, None, *g_levels, *g_weights, *g_biases)


def pack_levels_groupnorm(conv_outputs, group_norms, out_dtype=None):
    """``memory, spatial_shapes, level_start_index`` from the per-level CONV outputs (N, C, H_l, W_l) and the
    ``nn.GroupNorm`` modules that follow them in Detector.input_proj -- equals
    ``torch.cat([gn(x).flatten(2).transpose(1, 2) for x, gn in zip(conv_outputs, group_norms)], 1)`` in fp32 (or rounded
    to ``out_dtype`` = torch.bfloat16)."""
    gn0 = group_norms[0]
    for gn in group_norms:
        if gn.num_groups != gn0.num_groups or gn.eps != gn0.eps or gn.weight is None or gn.bias is None:
            raise RuntimeError("pack_levels_groupnorm needs affine GroupNorms with the same num_groups / eps on every level")
    n_levels = len(conv_outputs)
    memory = PackLevelsGroupNormFunction.apply(gn0.num_groups, gn0.eps, out_dtype or torch.float32, n_levels,
                                               *[x.float() for x in conv_outputs], *[gn.weight for gn in group_norms],
                                               *[gn.bias for gn in group_norms])
    hw = [(int(t.shape[2]), int(t.shape[3])) for t in conv_outputs]
    spatial_shapes = torch.as_tensor(hw, dtype=torch.long, device=memory.device)
    starts = [0]
    for h, w in hw[:-1]:
        starts.append(starts[-1] + h * w)
    level_start_index = torch.as_tensor(starts, dtype=torch.long, device=memory.device)
    return memory, spatial_shapes, level_start_index


def pack_levels(levels):
    """``memory, spatial_shapes, level_start_index`` from a list of (N, C, H_l, W_l) feature maps -- the tensors
    prepare_od_inputs builds (models/detection/det_module.py:146-158); shapes/starts are int64 tensors on the device."""
    memory = PackLevelsFunction.apply(*levels)
    hw = [(int(t.shape[2]), int(t.shape[3])) for t in levels]
    spatial_shapes = torch.as_tensor(hw, dtype=torch.long, device=memory.device)
    starts = [0]
    for h, w in hw[:-1]:
        starts.append(starts[-1] + h * w)
    level_start_index = torch.as_tensor(starts, dtype=torch.long, device=memory.device)
    return memory, spatial_shapes, level_start_index


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
