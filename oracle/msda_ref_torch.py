"""TEST INFRASTRUCTURE ONLY -- the reference's CPU path, restated for timing and cross-checks.

The reference has no CPU kernel (models/ops/src/cpu/ms_deform_attn_cpu.cpp:26,39 raise); what
runs on a host is ``ms_deform_attn_core_pytorch`` (models/ops/functions/ms_deform_attn_func.py:41-61):
one ``F.grid_sample`` per level on (N*M, D, H, W) maps with grid = 2*loc - 1, bilinear, zero
padding, align_corners=False, then a weighted sum over the L*P samples.  This file restates that
composition (same torch ops, so the same CPU cost profile) for bench.py's ``cpu_baseline`` and
``--impl reference`` legs and for tests.  It is never imported by grit_b200/.

Parity status: PINNED (tests/test_oracle_golden.py compares it with the golden vectors made
from the reference function itself).
"""
import torch
import torch.nn.functional as F


def grid_sample_reference(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); spatial_shapes iterable of (H,W); loc (N,Lq,M,L,P,2); attn (N,Lq,M,L,P)
    -> (N, Lq, M*D).  Differentiable through autograd."""
    n, _, m, d = value.shape
    lq, nl, npts = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    grids = sampling_locations * 2 - 1
    per_level = []
    for lvl, (chunk, (h, w)) in enumerate(zip(value.split(sizes, dim=1), spatial_shapes)):
        fmap = chunk.permute(0, 2, 3, 1).reshape(n * m, d, int(h), int(w))
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * m, lq, npts, 2)
        per_level.append(F.grid_sample(fmap, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
    sampled = torch.cat(per_level, dim=-1)  # (N*M, D, Lq, L*P), level-major like attn's (L, P)
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, nl * npts)
    out = (sampled * weights).sum(-1)  # (N*M, D, Lq)
    return out.view(n, m * d, lq).transpose(1, 2).contiguous()


def forward_backward(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """One fwd + autograd bwd pass; returns (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().requires_grad_(True)
    s = sampling_locations.detach().requires_grad_(True)
    a = attention_weights.detach().requires_grad_(True)
    out = grid_sample_reference(v, spatial_shapes, s, a)
    out.backward(grad_output)
    return out.detach(), v.grad, s.grad, a.grad
