"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (ctypes over oracle/libmsda_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under grit_b200/ does: the product path is the CUDA library and it
raises when that library is missing.

Parity status: PINNED against golden vectors generated from the reference's own
``ms_deform_attn_core_pytorch`` (/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61)
by tests/golden/gen_golden.py; see tests/test_oracle_golden.py.

Three restatements live here:
  * ``forward`` / ``backward``      -- the C tap-level oracle (msda_oracle_impl.h), fp64 or fp32,
                                       OpenMP-threaded; follows ms_deform_im2col_cuda.cuh:33-159,272-296.
  * ``forward_numpy``               -- an independent, loop-free numpy restatement (fp64) used to
                                       cross-check the C code on small cases.
  * ``module_forward_numpy``        -- the MSDeformAttn module arithmetic
                                       (models/ops/modules/ms_deform_attn.py:93-117) in numpy fp64:
                                       value_proj, mask fill, offsets, softmax, sampling locations.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/libmsda_oracle.so with the committed Makefile (gcc only, no GPU needed)."""
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("msda_oracle.c", "msda_oracle_impl.h"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_mtime:
        subprocess.run(["make", "-C", _HERE, "-B", "libmsda_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i64 = ctypes.c_int64
        vp = ctypes.c_void_p
        for sfx in ("f64", "f32"):
            fwd = getattr(_lib, f"msda_oracle_forward_{sfx}")
            fwd.restype = None
            fwd.argtypes = [vp] * 6 + [i64] * 7
            bwd = getattr(_lib, f"msda_oracle_backward_{sfx}")
            bwd.restype = None
            bwd.argtypes = [vp] * 9 + [i64] * 7
    return _lib


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _prep(value, shapes, level_start, loc, attn, dtype):
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    attn = np.ascontiguousarray(attn, dtype=dtype)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    level_start = np.ascontiguousarray(level_start, dtype=np.int64)
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert M2 == M and two == 2 and attn.shape == (N, Lq, M, L, P)
    assert shapes.shape == (L, 2) and level_start.shape == (L,)
    return value, shapes, level_start, loc, attn, (N, S, M, D, L, Lq, P)


def forward(value, shapes, level_start, loc, attn, dtype=np.float64) -> np.ndarray:
    """C oracle forward.  Returns (N, Lq, M*D) like MSDeformAttnFunction.forward."""
    lib = _load()
    value, shapes, level_start, loc, attn, dims = _prep(value, shapes, level_start, loc, attn, dtype)
    N, S, M, D, L, Lq, P = dims
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f64 if dtype == np.float64 else lib.msda_oracle_forward_f32
    fn(_ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(attn), _ptr(out), *dims)
    return out


def backward(value, shapes, level_start, loc, attn, grad_out, dtype=np.float64):
    """C oracle backward.  Returns (grad_value, grad_sampling_loc, grad_attn_weight)."""
    lib = _load()
    value, shapes, level_start, loc, attn, dims = _prep(value, shapes, level_start, loc, attn, dtype)
    grad_out = np.ascontiguousarray(grad_out, dtype=dtype)
    gv = np.zeros_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = lib.msda_oracle_backward_f64 if dtype == np.float64 else lib.msda_oracle_backward_f32
    fn(_ptr(value), _ptr(shapes), _ptr(level_start), _ptr(loc), _ptr(attn), _ptr(grad_out),
       _ptr(gv), _ptr(gl), _ptr(ga), *dims)
    return gv, gl, ga


def forward_numpy(value, shapes, level_start, loc, attn) -> np.ndarray:
    """Vectorised numpy fp64 restatement (independent of the C code; small cases only).

    Same rules as msda_oracle_impl.h: pixel coords = loc*size - 0.5, the (-1, size) validity
    window, four taps each zeroed when outside the map
    (reference: ms_deform_im2col_cuda.cuh:33-84, 285-292).
    """
    value = np.asarray(value, dtype=np.float64)
    loc = np.asarray(loc, dtype=np.float64)
    attn = np.asarray(attn, dtype=np.float64)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = np.zeros((N, Lq, M, D))
    b_idx = np.arange(N)[:, None, None, None]
    m_idx = np.arange(M)[None, None, :, None]
    for l in range(L):
        H, W = int(shapes[l][0]), int(shapes[l][1])
        start = int(level_start[l])
        x = loc[:, :, :, l, :, 0] * W - 0.5  # (N, Lq, M, P)
        y = loc[:, :, :, l, :, 1] * H - 0.5
        with np.errstate(invalid="ignore"):
            live = (y > -1) & (x > -1) & (y < H) & (x < W)
        xs = np.where(live, x, 0.0)
        ys = np.where(live, y, 0.0)
        r0 = np.floor(ys).astype(np.int64)
        c0 = np.floor(xs).astype(np.int64)
        lh, lw = ys - r0, xs - c0
        for dr, dc, wt in ((0, 0, (1 - lh) * (1 - lw)), (0, 1, (1 - lh) * lw),
                           (1, 0, lh * (1 - lw)), (1, 1, lh * lw)):
            r, c = r0 + dr, c0 + dc
            ok = live & (r >= 0) & (r <= H - 1) & (c >= 0) & (c <= W - 1)
            pix = start + np.clip(r, 0, H - 1) * W + np.clip(c, 0, W - 1)
            tap = value[b_idx, pix, m_idx, :]  # (N, Lq, M, P, D)
            coef = np.where(ok, wt, 0.0) * attn[:, :, :, l, :]
            out += (tap * coef[..., None]).sum(axis=3)
    return out.reshape(N, Lq, M * D)


def softmax_np(x, axis=-1):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def module_forward_numpy(params, query, reference_points, input_flatten, shapes, level_start,
                         padding_mask=None, n_heads=8, n_levels=4, n_points=4, return_parts=False):
    """MSDeformAttn.forward in numpy fp64 (models/ops/modules/ms_deform_attn.py:93-117).

    ``params`` maps the reference's state_dict keys (``sampling_offsets.weight`` ... ``output_proj.bias``)
    to arrays.  Uses the C oracle for the core op.
    """
    f8 = lambda k: np.asarray(params[k], dtype=np.float64)
    query = np.asarray(query, dtype=np.float64)
    ref = np.asarray(reference_points, dtype=np.float64)
    src = np.asarray(input_flatten, dtype=np.float64)
    N, Lq, C = query.shape
    S = src.shape[1]
    M, L, P = n_heads, n_levels, n_points
    value = src @ f8("value_proj.weight").T + f8("value_proj.bias")
    if padding_mask is not None:
        value = np.where(np.asarray(padding_mask, dtype=bool)[..., None], 0.0, value)
    value = value.reshape(N, S, M, C // M)
    offs = (query @ f8("sampling_offsets.weight").T + f8("sampling_offsets.bias")).reshape(N, Lq, M, L, P, 2)
    logits = (query @ f8("attention_weights.weight").T + f8("attention_weights.bias")).reshape(N, Lq, M, L * P)
    attn = softmax_np(logits, -1).reshape(N, Lq, M, L, P)
    shp = np.asarray(shapes, dtype=np.float64)
    if ref.shape[-1] == 2:
        normalizer = np.stack([shp[:, 1], shp[:, 0]], -1)  # (W, H) per level
        loc = ref[:, :, None, :, None, :] + offs / normalizer[None, None, None, :, None, :]
    elif ref.shape[-1] == 4:
        loc = ref[:, :, None, :, None, :2] + offs / P * ref[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")
    core = forward(value, shapes, level_start, loc, attn)
    out = core @ f8("output_proj.weight").T + f8("output_proj.bias")
    if return_parts:
        return out, dict(value=value, loc=loc, attn=attn, core=core)
    return out
