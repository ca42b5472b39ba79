"""``MSDeformAttn`` -- drop-in for the reference's ``models/ops/modules/ms_deform_attn.py:22-119``.

Same constructor, attributes (``im2col_step = 64``), parameter names and shapes (so published GRIT / Deformable-DETR
checkpoints load: ``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` ``.weight/.bias``), same
initialisation and the same ``forward`` signature and exceptions.  The four projections stay ``nn.Linear`` (cuBLAS);
the sampling core is ``MSDeformAttnFunction`` -> sm_100a kernels.

One knob is added: ``validate_shapes`` (default True, the reference behaviour).  The reference asserts
``sum(H_l*W_l) == Len_in`` on device tensors every call (reference :93), which costs a device->host sync; callers that
construct ``input_spatial_shapes`` from the same Python ints as ``input_flatten`` may set it to False.
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from ... import _lib
from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        """
        :param d_model   hidden dimension
        :param n_levels  number of feature levels
        :param n_heads   number of attention heads
        :param n_points  number of sampling points per attention head per feature level
        """
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, but got {} and {}'.format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a "
                          "power of 2 which is more efficient in our CUDA implementation.")

        self.im2col_step = 64
        self.validate_shapes = True
        self.fused = True

        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)

        self._reset_parameters()

    def _reset_parameters(self):
        """Reference init (:56-71): zero offset weights, offset bias on a ring of directions (one per head, scaled by
        point index), zero attention logits, Xavier projections."""
        constant_(self.sampling_offsets.weight.data, 0.)
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        ring = torch.stack([angle.cos(), angle.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        scale = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        bias = ring.view(self.n_heads, 1, 1, 2) * scale.expand(self.n_heads, self.n_levels, self.n_points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, value=None, valid_ratios=None):
        """
        :param query                    (N, Length_{query}, C)
        :param reference_points         (N, Length_{query}, n_levels, 2) in [0, 1], top-left (0,0), bottom-right (1,1),
                                        or (N, Length_{query}, n_levels, 4): (cx, cy, w, h) reference boxes
        :param input_flatten            (N, sum_l H_l*W_l, C)
        :param input_spatial_shapes     (n_levels, 2), [(H_0, W_0), ..., (H_{L-1}, W_{L-1})]
        :param input_level_start_index  (n_levels,), [0, H_0*W_0, H_0*W_0+H_1*W_1, ...]
        :param input_padding_mask       (N, sum_l H_l*W_l), True for padding elements
        :param value                    optional (N, sum_l H_l*W_l, C): this layer's ``value_proj(input_flatten)`` computed
                                        elsewhere (see ``hoisted_value_proj``: GRIT's six decoder layers project the SAME
                                        memory, det_module.py:191-198, so one batched GEMM can serve all of them).  Not in
                                        the reference signature; ``None`` gives the reference behaviour.  A supplied
                                        tensor is never modified (the padding mask is applied to a copy).
        :param valid_ratios             optional (N, n_levels, 2) [w-ratio, h-ratio]: ``reference_points`` is then the
                                        UN-EXPANDED (N, Length_{query}, 2|4) tensor and level l uses
                                        ``reference_points * valid_ratios[:, l]`` -- the scaling the reference's decoder
                                        layer does before calling this module (det_module.py:323-328), done inside the
                                        fused kernels instead of materialising (N, Lq, n_levels, 2|4).  Not in the
                                        reference signature.
        :return output                  (N, Length_{query}, C)
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if self.validate_shapes:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

        if reference_points.shape[-1] not in (2, 4):
            raise ValueError('Last dim of reference_points must be 2 or 4, but get {} instead.'.format(
                reference_points.shape[-1]))
        # In-place masking is only done on a tensor this module produced itself with a plain nn.Linear (addmm does not
        # need its output in backward); a caller-supplied `value=` or a wrapped value_proj is masked out of place.
        own_value = value is None and type(self.value_proj) is nn.Linear
        if value is None:
            value = self.value_proj(input_flatten)
        sampling_offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if self.fused and query.is_cuda:
            # fused path: softmax, offsets/normaliser + reference points and the mask fill happen
            # inside the gather kernels; falls through to the reference-shaped path when no specialisation exists or
            # the tensors are not laid out the way the kernels index them (_lib.fused_supported checks everything)
            mask = input_padding_mask
            if mask is not None and not own_value:
                value, mask = value.masked_fill(mask[..., None], float(0)), None
            value4 = value.contiguous().view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
            if valid_ratios is None:
                ref32 = reference_points.float().expand(N, Len_q, self.n_levels,
                                                        reference_points.shape[-1]).contiguous()
                vr32 = None
            else:
                ref32, vr32 = reference_points.float().contiguous(), valid_ratios.float().contiguous()
            offs32, logits32 = sampling_offsets.float().contiguous(), attention_weights.float().contiguous()
            if _lib.fused_supported(value4, input_spatial_shapes, input_level_start_index, offs32, logits32, ref32,
                                    mask, vr32):
                output = MSDeformAttnFusedFunction.apply(value4, input_spatial_shapes, input_level_start_index,
                                                         offs32, logits32, ref32, mask, vr32, True)
                return self.output_proj(output)
            if mask is None:
                input_padding_mask = None  # already applied above
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.contiguous().view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        attention_weights = F.softmax(attention_weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if valid_ratios is not None:  # the reference decoder layer's expansion (det_module.py:323-328)
            vr = valid_ratios if reference_points.shape[-1] == 2 else torch.cat([valid_ratios, valid_ratios], -1)
            reference_points = reference_points[:, :, None] * vr[:, None]

        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, sampling_locations,
                                            attention_weights, self.im2col_step)
        return self.output_proj(output)


def hoisted_value_proj(modules, input_flatten, padding_mask=None):
    """``value_proj`` of several MSDeformAttn layers that read the SAME memory, as ONE GEMM.

    GRIT's decoder runs six layers over an unchanged ``src`` (models/detection/det_module.py:191-198); each layer's
    ``value_proj`` is an (N*S, C) x (C, C) GEMM that re-reads ``src``.  Concatenating the weight matrices gives one
    batched GEMM (n launches -> 1).  Returns a tuple of per-layer ``value`` tensors (N, S, C), each contiguous.
    Pass ``value=vals[i]`` to layer i.  Gradients reach every layer's ``value_proj.weight/bias`` and ``input_flatten``
    through autograd.  With ``padding_mask`` (N, S) the masked pixels are zeroed here, once for all layers (the
    reference's ``value.masked_fill``, modules/ms_deform_attn.py:96-97); pass ``input_padding_mask=None`` to the layers then.
    """
    modules = list(modules)
    n_layers, c = len(modules), modules[0].d_model
    n, s_len, _ = input_flatten.shape
    # One batched GEMM whose output is already layer-major and contiguous per layer -- (n_layers, N*S, C) -- which is the
    # (N, S, M, D) layout the kernels index; a single (N*S, C) x (C, n*C) GEMM would need a strided re-layout copy of
    # all n outputs afterwards (measured: slower than n separate GEMMs at batch 64).
    weight_t = torch.stack([m.value_proj.weight.t() for m in modules], 0)  # (n, C_in, C_out)
    bias = torch.stack([m.value_proj.bias for m in modules], 0).unsqueeze(1)  # (n, 1, C_out)
    src = input_flatten.reshape(1, n * s_len, c).expand(n_layers, -1, -1)
    out = torch.baddbmm(bias, src, weight_t).view(n_layers, n, s_len, c)
    if padding_mask is not None:
        out = out.masked_fill(padding_mask[None, ..., None], float(0))
    return out.unbind(0)
