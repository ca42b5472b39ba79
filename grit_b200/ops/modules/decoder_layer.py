"""The decoder layer around the hot path and its consumers.

``DeformableTransformerDecoderLayer`` is a drop-in for the reference's class of the same name
(models/detection/det_module.py:272-349): same constructor arguments, same sub-module names (``cross_attn``,
``dropout1``, ``norm1``, ``self_attn``, ``dropout2``, ``norm2``, ``linear1``, ``dropout3``, ``linear2``, ``dropout4``,
``norm3``) so reference checkpoints load key for key, same ``forward`` signature and arithmetic.  What differs is how the
arithmetic is launched:

* the three ``tgt = tgt + dropout(tgt2); tgt = norm(tgt)`` epilogues are one fused kernel each
  (``add_dropout_layer_norm`` -> msda_add_dropout_ln_*), instead of dropout + add + LayerNorm;
* the reference-point x valid-ratio scaling (:323-328) happens inside the fused deformable-attention kernels
  (``MSDeformAttn.forward(..., valid_ratios=...)``) instead of materialising (N, Lq, L, 2|4);
* the cross-attention takes an optional pre-projected ``value`` (``hoisted_value_proj``: the six layers project the same
  memory, :191-198).

Self-attention (nn.MultiheadAttention) and the FFN Linears stay PyTorch / cuBLAS: dense GEMM work, out of scope.

``run_decoder`` walks a stack of such layers the way DetectionModule.forward does (:181-211, without the box-refinement
heads, which belong to the detector), optionally under ONE CUDA graph (``GraphedDecoder``): at GRIT's operating point
(150 queries) a layer's kernels take tens of microseconds, comparable to their launch overhead.
``extract_region_features`` is the op-side pipeline of tools/extract_features.py:48-155: batch 64, forward only,
``reg_feat`` per layer, fp32, ready for the caller's HDF5 writer.
"""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from ..functions import add_dropout_layer_norm
from .ms_deform_attn import MSDeformAttn, hoisted_value_proj


def _get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError(F"activation should be relu/gelu, not {activation}.")


class _DropPath(nn.Module):
    """Stochastic depth per sample (what timm's DropPath does; timm is not a dependency of this package)."""

    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x.div(keep) * mask


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 drop_path=0.):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)

        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

        self.drop_path = _DropPath(drop_path) if drop_path > 0. else None
        self.fused_epilogue = True  # False: the reference's dropout + add + LayerNorm launches (for A/B and tests)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _add_norm(self, x, z, dropout, norm):
        if self.fused_epilogue:
            return add_dropout_layer_norm(x, z, norm, dropout.p, self.training)
        return norm(x + dropout(z))

    def forward_ffn(self, tgt):
        tgt2 = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
        return self._add_norm(tgt, tgt2, self.dropout4, self.norm3)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, src_level_start_index,
                src_valid_ratios, src_padding_mask=None, value=None):
        if reference_points.shape[-1] not in (2, 4):
            raise AssertionError("reference_points must be (N, Lq, 2) or (N, Lq, 4)")

        q = k = self.with_pos_embed(tgt, query_pos)
        tgt2 = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = self._add_norm(tgt, tgt2, self.dropout2, self.norm2)

        # the valid-ratio scaling of the reference points (reference :323-328) is done inside the attention kernels
        tgt2 = self.cross_attn(self.with_pos_embed(tgt, query_pos), reference_points, src, src_spatial_shapes,
                               src_level_start_index, src_padding_mask, value=value, valid_ratios=src_valid_ratios)

        if self.drop_path is None:
            tgt = self._add_norm(tgt, tgt2, self.dropout1, self.norm1)
            tgt = self.forward_ffn(tgt)
        else:
            tgt = tgt + self.drop_path(self.dropout1(tgt2))
            tgt2 = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
            tgt = tgt + self.drop_path(self.dropout4(tgt2))
            tgt = self.norm3(tgt)
        return tgt


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


# Hoisting the value projections saves launches, not traffic (every layer's value is written once and read once either way),
# and the batched GEMM is no faster than six plain ones: measured on B200 (scripts/decoder_bench.py) it wins only while the
# decoder is launch-bound -- 4.4 vs 4.6 ms at N=4, 384x640 -- and loses above that (fp32 GEMMs: 34.0 vs 32.2 ms at N=64; with
# TF32 GEMMs 11.5 vs 5.3 ms), besides keeping n_layers values alive at once.  "auto" hoists below this many bytes of values.
HOIST_AUTO_MAX_BYTES = 256 << 20


def run_decoder(layers, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index, valid_ratios,
                padding_mask=None, hoist_value_proj="auto", return_intermediate=True):
    """The decoder loop of DetectionModule.forward (det_module.py:191-211) over ``layers`` with fixed reference points
    (box refinement belongs to the detection heads).  Returns (n_layers, N, Lq, C) when ``return_intermediate`` else the
    last layer's (N, Lq, C).  ``hoist_value_proj``: one batched value_proj GEMM (+ one mask fill) for all layers --
    True, False, or "auto" (only while all the layers' values together stay below HOIST_AUTO_MAX_BYTES)."""
    values = [None] * len(layers)
    mask = padding_mask
    if hoist_value_proj == "auto":
        hoist_value_proj = src.numel() * src.element_size() * len(layers) <= HOIST_AUTO_MAX_BYTES
    if hoist_value_proj and len(layers) > 1:
        values = hoisted_value_proj([layer.cross_attn for layer in layers], src, padding_mask)
        mask = None  # applied once, inside hoisted_value_proj
    outs = []
    for layer, value in zip(layers, values):
        tgt = layer(tgt, query_pos, reference_points, src, spatial_shapes, level_start_index, valid_ratios, mask,
                    value=value)
        if return_intermediate:
            outs.append(tgt)
    return torch.stack(outs) if return_intermediate else tgt


class GraphedDecoder:
    """``run_decoder`` (forward only, eval mode) captured in ONE CUDA graph -- ``graphed_layers`` of VERDICT r1 item 3.

    The C ABI is capture-safe (no allocation, no synchronisation, device-resident level metadata), so the whole stack --
    cuBLAS GEMMs, attention kernels, fused epilogues -- replays with a single launch.  Static input buffers are owned by
    this object; ``__call__`` copies the caller's tensors in, replays, and returns the static output (clone it if you keep
    it across calls).  Shapes are fixed at construction, as with any CUDA graph."""

    def __init__(self, layers, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index, valid_ratios,
                 padding_mask=None, hoist_value_proj="auto", return_intermediate=True, warmup=3):
        self.layers = layers
        for layer in layers:
            layer.eval()
            layer.cross_attn.validate_shapes = False  # the shape assert is a device->host sync: illegal under capture
        self.static = dict(tgt=tgt.clone(), query_pos=query_pos.clone(), reference_points=reference_points.clone(),
                           src=src.clone(), valid_ratios=valid_ratios.clone(),
                           padding_mask=None if padding_mask is None else padding_mask.clone())
        self.shapes, self.lsi = spatial_shapes.clone(), level_start_index.clone()
        self.kw = dict(hoist_value_proj=hoist_value_proj, return_intermediate=return_intermediate)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        s = self.static
        return run_decoder(self.layers, s["tgt"], s["query_pos"], s["reference_points"], s["src"], self.shapes, self.lsi,
                           s["valid_ratios"], s["padding_mask"], **self.kw)

    def __call__(self, tgt, query_pos, reference_points, src, valid_ratios, padding_mask=None):
        s = self.static
        s["tgt"].copy_(tgt), s["query_pos"].copy_(query_pos), s["reference_points"].copy_(reference_points)
        s["src"].copy_(src), s["valid_ratios"].copy_(valid_ratios)
        if s["padding_mask"] is not None and padding_mask is not None:
            s["padding_mask"].copy_(padding_mask)
        self.graph.replay()
        return self.out


class _DecoderStack(nn.Module):
    """``run_decoder`` as a module (what ``torch.cuda.make_graphed_callables`` wants): tensors in, tensor out."""

    def __init__(self, layers, hoist_value_proj, return_intermediate):
        super().__init__()
        self.layers = layers if isinstance(layers, nn.ModuleList) else nn.ModuleList(layers)
        self.kw = dict(hoist_value_proj=hoist_value_proj, return_intermediate=return_intermediate)

    def forward(self, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index, valid_ratios,
                padding_mask=None):
        return run_decoder(self.layers, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index,
                           valid_ratios, padding_mask, **self.kw)


def graphed_training_decoder(layers, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index,
                             valid_ratios, padding_mask=None, hoist_value_proj="auto", return_intermediate=True):
    """The decoder stack for TRAINING with its forward and its backward each replayed from a CUDA graph
    (``torch.cuda.make_graphed_callables``).  GRIT trains its detector at batch 4 per GPU (configs/detection/
    train_config.yaml:70), where the six layers are ~900 kernel launches per step and the step is bound by the host, not by
    the GPU; the C ABI is capture-safe and the fused autograd functions allocate through torch, so the whole stack --
    cuBLAS GEMMs, sampling kernels, epilogues, dropout -- captures.  The arguments are sample tensors of the shapes that
    will be used (their ``requires_grad`` flags must match the later calls); returns a callable with ``run_decoder``'s
    positional signature.  Parameters receive ``.grad`` as usual; shapes are fixed, as with any CUDA graph."""
    for layer in layers:
        layer.cross_attn.validate_shapes = False  # the shape assert is a device->host sync: illegal under capture
    stack = _DecoderStack(layers, hoist_value_proj, return_intermediate)
    sample = (tgt, query_pos, reference_points, src, spatial_shapes, level_start_index, valid_ratios)
    if padding_mask is not None:
        sample = sample + (padding_mask,)
    return torch.cuda.make_graphed_callables(stack, sample)


def extract_region_features(layers, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index,
                            valid_ratios, padding_mask=None, graphed=None):
    """Region features of a batch, the op-side pipeline of the reference's tools/extract_features.py:80-119 (batch 64,
    ``model.eval()``, ``torch.no_grad()``): the six decoder layers forward only, one CUDA graph when ``graphed`` (a
    ``GraphedDecoder`` built for these shapes) is given.  Returns ``reg_feat`` (n_layers, N, Lq, C) float32 -- the
    reference stores the last layer's (N, Lq, C) slice as HDF5 dataset ``reg_feat`` (:85, :112); writing the file is the
    caller's business (h5py is not a dependency of this package)."""
    with torch.no_grad():
        if graphed is not None:
            out = graphed(tgt, query_pos, reference_points, src, valid_ratios, padding_mask)
        else:
            was_training = [layer.training for layer in layers]
            for layer in layers:
                layer.eval()
            out = run_decoder(layers, tgt, query_pos, reference_points, src, spatial_shapes, level_start_index,
                              valid_ratios, padding_mask)
            for layer, flag in zip(layers, was_training):
                layer.train(flag)
    return out.float()
