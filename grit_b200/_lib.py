"""ctypes binding of the C ABI in include/msda.h (the thin layer north_star asks for).

This is the ONLY route from Python to the kernels.  There is no CPU or pure-PyTorch fallback: if
``libmsda_b200.so`` is missing the import of anything that needs it raises, and CPU tensors are
rejected the way the reference rejects them (models/ops/src/ms_deform_attn.h:38,60).
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import build as _build

MSDA_F32, MSDA_F64, MSDA_BF16 = 0, 1, 2
FLAG_ZERO_GRAD_VALUE = 1 << 0
FLAG_DETERMINISTIC = 1 << 1
FLAG_FORCE_GENERIC = 1 << 2
FLAG_ALIGNED16 = 1 << 3  # msda_backward_workspace_bytes only: every tensor of the coming call is 16-byte aligned

_DTYPE_CODE = {torch.float32: MSDA_F32, torch.float64: MSDA_F64, torch.bfloat16: MSDA_BF16}


class MsdaDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in
                ("batch", "spatial_size", "num_heads", "channels", "num_levels", "num_query", "num_point")]


_lib = None
_lock = threading.Lock()
_NVTX = os.environ.get("GRIT_B200_NVTX", "0") not in ("", "0")


class _nvtx_range:
    """NVTX range around a library call when GRIT_B200_NVTX=1 (shows up in nsys / ncu --nvtx timelines); a no-op
    otherwise.  The reference has no tracing hooks at all."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def library_path() -> str:
    """The in-tree library; GRIT_B200_LIB names another build of the same ABI (same-box A/B of two kernel versions)."""
    return os.environ.get("GRIT_B200_LIB") or _build.LIB_PATH


def load():
    """dlopen the CUDA library; raise loudly when it is absent (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: the sm_100a CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `python -m grit_b200.build`). "
                "grit_b200 has no CPU fallback.")
        lib = ctypes.CDLL(path)
        vp, i64p, dimsp = ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(MsdaDims)
        lib.msda_abi_version.restype = ctypes.c_int
        lib.msda_last_error.restype = ctypes.c_char_p
        lib.msda_last_kernel.restype = ctypes.c_char_p
        lib.msda_launch_count.restype = ctypes.c_int64
        lib.msda_launch_count.argtypes = [ctypes.c_int]
        lib.msda_set_tuning.restype = ctypes.c_int
        lib.msda_set_tuning.argtypes = [ctypes.c_char_p, ctypes.c_int]
        lib.msda_forward.restype = ctypes.c_int
        lib.msda_forward.argtypes = [vp, i64p, i64p, vp, vp, vp, dimsp, ctypes.c_int, ctypes.c_uint, vp]
        lib.msda_backward_strategy.restype = ctypes.c_int
        lib.msda_backward_strategy.argtypes = [dimsp, ctypes.c_int, ctypes.c_uint]
        lib.msda_backward_workspace_bytes.restype = ctypes.c_size_t
        lib.msda_backward_workspace_bytes.argtypes = [dimsp, ctypes.c_int, ctypes.c_uint]
        lib.msda_backward.restype = ctypes.c_int
        lib.msda_backward.argtypes = [vp, i64p, i64p, vp, vp, vp, vp, vp, vp, dimsp, ctypes.c_int, ctypes.c_uint,
                                      vp, ctypes.c_size_t, vp]
        lib.msda_pack_levels.restype = ctypes.c_int
        lib.msda_pack_levels.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
                                         ctypes.c_int64, ctypes.c_int64, vp, ctypes.c_int, ctypes.c_int, vp]
        lib.msda_pack_levels_groupnorm.restype = ctypes.c_int
        lib.msda_pack_levels_groupnorm.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                                   ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                                   ctypes.c_float, vp, ctypes.c_int, vp, vp]
        lib.msda_probe_ceiling.restype = ctypes.c_int
        lib.msda_probe_ceiling.argtypes = [ctypes.c_int, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int64), vp]
        lib.msda_fused_supported.restype = ctypes.c_int
        lib.msda_fused_supported.argtypes = [dimsp, ctypes.c_int, ctypes.c_int]
        lib.msda_mask_rows.restype = ctypes.c_int
        lib.msda_mask_rows.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int64, vp]
        lib.msda_fused_forward.restype = ctypes.c_int
        lib.msda_fused_forward.argtypes = [vp, i64p, i64p, vp, vp, vp, ctypes.c_int, vp, dimsp, ctypes.c_int,
                                           ctypes.c_uint, vp]
        lib.msda_fused_backward.restype = ctypes.c_int
        lib.msda_fused_backward.argtypes = [vp, i64p, i64p, vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, dimsp,
                                            ctypes.c_int, ctypes.c_uint, vp, ctypes.c_size_t, vp]
        lib.msda_fused_forward_vr.restype = ctypes.c_int
        lib.msda_fused_forward_vr.argtypes = [vp, i64p, i64p, vp, vp, vp, vp, ctypes.c_int, vp, dimsp, ctypes.c_int,
                                              ctypes.c_uint, vp]
        lib.msda_fused_backward_vr.restype = ctypes.c_int
        lib.msda_fused_backward_vr.argtypes = [vp, i64p, i64p, vp, vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, dimsp,
                                               ctypes.c_int, ctypes.c_uint, vp, ctypes.c_size_t, vp]
        lib.msda_add_dropout_ln_supported.restype = ctypes.c_int
        lib.msda_add_dropout_ln_supported.argtypes = [ctypes.c_int64]
        lib.msda_add_dropout_ln_workspace_bytes.restype = ctypes.c_size_t
        lib.msda_add_dropout_ln_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
        lib.msda_add_dropout_ln_forward.restype = ctypes.c_int
        lib.msda_add_dropout_ln_forward.argtypes = [vp, vp, vp, ctypes.c_float, vp, vp, ctypes.c_float, vp, vp, vp, vp,
                                                    ctypes.c_int64, ctypes.c_int64, vp]
        lib.msda_add_dropout_ln_backward.restype = ctypes.c_int
        lib.msda_add_dropout_ln_backward.argtypes = [vp, vp, vp, vp, vp, ctypes.c_float, vp, vp, vp, vp, vp, vp,
                                                     ctypes.c_size_t, ctypes.c_int64, ctypes.c_int64, vp]
        lib.msda_host_session_create.restype = ctypes.c_int
        lib.msda_host_session_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), dimsp, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int]
        lib.msda_host_session_destroy.restype = None
        lib.msda_host_session_destroy.argtypes = [vp]
        lib.msda_host_forward_backward.restype = ctypes.c_int
        lib.msda_host_forward_backward.argtypes = [vp, vp, i64p, i64p, vp, vp, vp, vp, vp, vp, vp, dimsp,
                                                   ctypes.c_uint]
        lib.msda_host_submit.restype = ctypes.c_int
        lib.msda_host_submit.argtypes = lib.msda_host_forward_backward.argtypes
        lib.msda_host_wait.restype = ctypes.c_int
        lib.msda_host_wait.argtypes = [vp]
        if lib.msda_abi_version() != 1:
            raise RuntimeError(f"{path}: ABI version {lib.msda_abi_version()} != 1")
        _lib = lib
    return _lib


def last_kernel() -> str:
    return load().msda_last_kernel().decode()


def launch_count(reset: bool = False) -> int:
    return int(load().msda_launch_count(1 if reset else 0))


def set_tuning(key: str, value: int) -> int:
    """A/B knob of the library (include/msda.h: msda_set_tuning: "variant", "hoist", "warps", "v3_threads",
    "bwd_mode", "bin_min_rows", "owned_max_taps"); returns the previous value."""
    prev = load().msda_set_tuning(key.encode(), int(value))
    if prev < 0:
        raise KeyError(key)
    return prev


def _raise(lib, rc, what):
    raise RuntimeError(f"{what} failed (code {rc}): {lib.msda_last_error().decode()}")


def _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output=None):
    """The reference's host-side preconditions (ms_deform_attn_cuda.cu:28-38, 93-105; ms_deform_attn.h:38,60)."""
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)]
    if grad_output is not None:
        named.append(("grad_output", grad_output))
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if t.device != value.device:
            raise RuntimeError(f"{name} is on {t.device}, value is on {value.device}")
    if value.dtype not in _DTYPE_CODE:
        raise RuntimeError(f'"ms_deform_attn" not implemented for \'{value.dtype}\'')
    loc_dtype = torch.float64 if value.dtype == torch.float64 else torch.float32
    for name, t in (("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        if t.dtype != loc_dtype:
            raise RuntimeError(f"expected {name} of dtype {loc_dtype} for value of dtype {value.dtype}, "
                               f"but found {t.dtype}")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 (torch.long) tensors")
    if grad_output is not None and grad_output.dtype != value.dtype:
        raise RuntimeError(f"grad_output dtype {grad_output.dtype} does not match value dtype {value.dtype}")
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5 or sampling_loc.shape[-1] != 2:
        raise RuntimeError("expected value (N,S,M,D), sampling_loc (N,Lq,M,L,P,2), attn_weight (N,Lq,M,L,P)")
    n, s, m, d = value.shape
    n2, lq, m2, l, p, _ = sampling_loc.shape
    if (n2, m2) != (n, m) or tuple(attn_weight.shape) != (n, lq, m, l, p):
        raise RuntimeError("value / sampling_loc / attn_weight shapes disagree")
    if tuple(spatial_shapes.shape) != (l, 2) or level_start_index.numel() != l:
        raise RuntimeError("spatial_shapes must be (L,2) and level_start_index (L,) with L = sampling_loc.size(3)")
    return MsdaDims(n, s, m, d, l, lq, p)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_CTX = _NullCtx()


def _device_ctx(device):
    """`torch.cuda.device(device)` only when `device` is not already current: the context manager and
    `torch.cuda.current_stream()` (which builds a Stream object) were ~25 us of host time per library call, a tenth of
    GRIT's batch-4 decoder step (scripts/host_overhead_profile.py)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL_CTX
    return torch.cuda.device(device)


def _raw_stream():
    """cudaStream_t of torch's current stream on the current device, as an integer."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # very old / future torch without the private accessor
        return _raw_stream()


_warned_generic = set()


def _warn_if_generic(lib, value, dims):
    """One warning per (D, L, P, dtype) when a large fp32/bf16 problem lands on the scalar generic kernels (the
    reference warns similarly about non-power-of-two head widths, modules/ms_deform_attn.py:37-40)."""
    if value.dtype == torch.float64 or dims.batch * dims.num_query * dims.num_heads < 100_000:
        return
    key = (dims.channels, dims.num_levels, dims.num_point, value.dtype)
    if key in _warned_generic or not lib.msda_last_kernel().startswith(b"fwd_generic"):
        return
    _warned_generic.add(key)
    import warnings
    warnings.warn(f"grit_b200: no specialised kernel for D={dims.channels}, L={dims.num_levels}, P={dims.num_point}, "
                  f"{value.dtype} (or the tensors are not 16-byte aligned); using the generic path, which is several "
                  "times slower.  Specialised: D in {16,32,64,128} with (L,P) in {(4,4),(4,8),(1,4),(1,8)}, "
                  "D in {32,64} with (L,P) in {(3,4),(5,4)}.")


def forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, flags: int = 0):
    """-> output (N, Lq, M*D).  Mirrors ms_deform_attn_cuda_forward (ms_deform_attn_cuda.cu:20-80)."""
    lib = load()
    dims = _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    out = torch.empty((dims.batch, dims.num_query, dims.num_heads * dims.channels), dtype=value.dtype,
                      device=value.device)
    with _device_ctx(value.device), _nvtx_range("msda_forward"):
        stream = _raw_stream()
        rc = lib.msda_forward(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc),
                              _ptr(attn_weight), _ptr(out), ctypes.byref(dims), _DTYPE_CODE[value.dtype], flags,
                              ctypes.c_void_p(stream))
    if rc:
        _raise(lib, rc, "msda_forward")
    if not flags & FLAG_FORCE_GENERIC:
        _warn_if_generic(lib, value, dims)
    return out


def backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, flags: int = 0):
    """-> (grad_value, grad_sampling_loc, grad_attn_weight).  Mirrors ms_deform_attn_cuda_backward (:83-153)."""
    lib = load()
    dims = _check_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output)
    code = _DTYPE_CODE[value.dtype]
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    # grad_value is allocated uninitialised: with FLAG_ZERO_GRAD_VALUE the library zero-fills it on the stream when its
    # strategy accumulates into it (row kernels), overwrites it from the workspace (bf16 / deterministic fold) or writes
    # every line exactly once (owned backward: no fill at all)
    grad_value = torch.empty_like(value)
    flags |= FLAG_ZERO_GRAD_VALUE
    aligned = all(t.data_ptr() % 16 == 0 for t in (value, sampling_loc, attn_weight, grad_output, grad_value, grad_loc))
    ws_bytes = lib.msda_backward_workspace_bytes(ctypes.byref(dims), code, flags | (FLAG_ALIGNED16 if aligned else 0))
    workspace = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=value.device) if ws_bytes else None
    with _device_ctx(value.device), _nvtx_range("msda_backward"):
        stream = _raw_stream()
        rc = lib.msda_backward(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc),
                               _ptr(attn_weight), _ptr(grad_output), _ptr(grad_value), _ptr(grad_loc),
                               _ptr(grad_attn), ctypes.byref(dims), code, flags,
                               _ptr(workspace) if workspace is not None else None, ws_bytes,
                               ctypes.c_void_p(stream))
    if rc:
        _raise(lib, rc, "msda_backward")
    return grad_value, grad_loc, grad_attn


def pack_levels(levels, memory=None, unpack: bool = False):
    """levels: list of contiguous CUDA (N, C, H_l, W_l) tensors; memory: (N, sum H_l*W_l, C).
    unpack=False writes `memory` from the levels (allocating it when None); unpack=True writes the levels from it."""
    lib = load()
    n, c = levels[0].shape[:2]
    hw = [int(t.shape[2] * t.shape[3]) for t in levels]
    for t in levels:
        if not (t.is_cuda and t.is_contiguous() and t.dtype == levels[0].dtype and tuple(t.shape[:2]) == (n, c)):
            raise RuntimeError("pack_levels expects contiguous CUDA (N, C, H, W) tensors of one dtype")
    if levels[0].dtype not in _DTYPE_CODE:
        raise RuntimeError(f"pack_levels not implemented for {levels[0].dtype}")
    if memory is None:
        memory = torch.empty((n, sum(hw), c), dtype=levels[0].dtype, device=levels[0].device)
    if not memory.is_contiguous() or tuple(memory.shape) != (n, sum(hw), c) or memory.dtype != levels[0].dtype:
        raise RuntimeError("memory must be a contiguous (N, S, C) tensor of the levels' dtype")
    ptrs = (ctypes.c_void_p * len(levels))(*[t.data_ptr() for t in levels])
    hws = (ctypes.c_int64 * len(levels))(*hw)
    with _device_ctx(memory.device):
        rc = lib.msda_pack_levels(ptrs, hws, len(levels), n, c, _ptr(memory), _DTYPE_CODE[memory.dtype],
                                  1 if unpack else 0, ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_pack_levels")
    return memory


def pack_levels_groupnorm(levels, weights, biases, num_groups, eps, out_dtype=torch.float32):
    """levels: contiguous CUDA fp32 (N, C, H_l, W_l) conv outputs; weights / biases: each level's GroupNorm affine (C,).
    -> (memory (N, S, C) of out_dtype, stats (L, N, G, 2) fp32 = mean, rstd)."""
    lib = load()
    n, c = levels[0].shape[:2]
    hw = [int(t.shape[2] * t.shape[3]) for t in levels]
    for t, w, b in zip(levels, weights, biases):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape[:2]) == (n, c)):
            raise RuntimeError("pack_levels_groupnorm expects contiguous CUDA fp32 (N, C, H, W) tensors")
        for p in (w, b):
            if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32 and tuple(p.shape) == (c,) and
                    p.device == t.device):
                raise RuntimeError("GroupNorm weight / bias must be contiguous CUDA fp32 (C,) tensors")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"pack_levels_groupnorm writes fp32 or bf16 memory, not {out_dtype}")
    memory = torch.empty((n, sum(hw), c), dtype=out_dtype, device=levels[0].device)
    stats = torch.empty((len(levels), n, num_groups, 2), dtype=torch.float32, device=levels[0].device)
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    hws = (ctypes.c_int64 * len(levels))(*hw)
    with _device_ctx(memory.device):
        rc = lib.msda_pack_levels_groupnorm(arr(levels), hws, len(levels), n, c, int(num_groups), arr(weights),
                                            arr(biases), float(eps), _ptr(memory), _DTYPE_CODE[out_dtype], _ptr(stats),
                                            ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_pack_levels_groupnorm")
    return memory, stats


def probe_ceiling(which: str, scratch, iters: int = 5):
    """G lines/s of the on-chip ceiling microbenchmark (include/msda.h: msda_probe_ceiling) over `scratch`
    (a contiguous fp32 CUDA tensor that is overwritten when which == 'red')."""
    lib = load()
    code = {"gather": 0, "red": 1}[which]
    lines = ctypes.c_int64(0)
    best = None
    with _device_ctx(scratch.device):
        st = ctypes.c_void_p(_raw_stream())
        for i in range(iters + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.msda_probe_ceiling(code, _ptr(scratch), scratch.numel() * scratch.element_size(),
                                        ctypes.byref(lines), st)
            e1.record()
            if rc:
                _raise(lib, rc, "msda_probe_ceiling")
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if i > 0:
                best = ms if best is None else min(best, ms)
    return lines.value / best / 1e6  # G lines/s (a line = 128 bytes)


def fused_dims(value, sampling_offsets, reference_points):
    n, s, m, d = value.shape
    _, lq, _, l, p, _ = sampling_offsets.shape
    return MsdaDims(n, s, m, d, l, lq, p)


def _fused_problem(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                   padding_mask=None, grad_output=None, valid_ratios=None):
    """Why the fused kernels cannot take these tensors (a string), or None when they can.  The fused entry points do
    raw pointer arithmetic on every argument, so shapes, dtypes, devices, contiguity and alignment are all checked
    here -- the same role _check_inputs plays for msda_forward/backward."""
    if not value.is_cuda:
        return "Not implemented on the CPU"
    if value.dtype not in (torch.float32, torch.bfloat16):
        return f"fused kernels take fp32 / bf16 values, got {value.dtype}"
    if value.dim() != 4 or sampling_offsets.dim() != 6 or sampling_offsets.shape[-1] != 2:
        return "expected value (N,S,M,D) and sampling_offsets (N,Lq,M,L,P,2)"
    n, s, m, d = value.shape
    n2, lq, m2, l, p, _ = sampling_offsets.shape
    if (n2, m2) != (n, m):
        return "value / sampling_offsets shapes disagree"
    if tuple(attn_logits.shape) != (n, lq, m, l * p):
        return f"attn_logits must be (N,Lq,M,L*P) = {(n, lq, m, l * p)}, got {tuple(attn_logits.shape)}"
    if valid_ratios is None:
        if reference_points.dim() != 4 or tuple(reference_points.shape[:3]) != (n, lq, l) or \
                reference_points.shape[-1] not in (2, 4):
            return f"reference_points must be (N,Lq,L,2|4) = {(n, lq, l)} + (2|4,), got {tuple(reference_points.shape)}"
    else:  # un-expanded reference points, scaled per level by the valid ratios inside the kernels
        if reference_points.dim() != 3 or tuple(reference_points.shape[:2]) != (n, lq) or \
                reference_points.shape[-1] not in (2, 4):
            return f"with valid_ratios, reference_points must be (N,Lq,2|4), got {tuple(reference_points.shape)}"
        if tuple(valid_ratios.shape) != (n, l, 2) or valid_ratios.dtype != torch.float32:
            return f"valid_ratios must be a float32 (N,L,2) = {(n, l, 2)} tensor"
    for name, t in (("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits),
                    ("reference_points", reference_points)):
        if t.dtype != torch.float32:
            return f"{name} must be float32, got {t.dtype}"
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        return "spatial_shapes and level_start_index must be int64 (torch.long) tensors"
    if tuple(spatial_shapes.shape) != (l, 2) or level_start_index.numel() != l:
        return "spatial_shapes must be (L,2) and level_start_index (L,)"
    named = [("value", value), ("sampling_offsets", sampling_offsets), ("attn_logits", attn_logits),
             ("reference_points", reference_points), ("spatial_shapes", spatial_shapes),
             ("level_start_index", level_start_index)]
    if grad_output is not None:
        if grad_output.dtype != value.dtype or grad_output.numel() != n * lq * m * d:
            return "grad_output must have value's dtype and N*Lq*M*D elements"
        named.append(("grad_output", grad_output))
    if padding_mask is not None:
        if padding_mask.dtype != torch.bool or tuple(padding_mask.shape) != (n, s):
            return f"padding_mask must be a bool (N,S) = {(n, s)} tensor"
        named.append(("padding_mask", padding_mask))
    if valid_ratios is not None:
        named.append(("valid_ratios", valid_ratios))
    for name, t in named:
        if t.device != value.device:
            return f"{name} is on {t.device}, value is on {value.device}"
        if not t.is_contiguous():
            return f"{name} tensor has to be contiguous"
    for name, t in named[:4] + named[6:7]:
        if t.data_ptr() % 16:
            return f"{name} storage is not 16-byte aligned"
    return None


def fused_supported(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                    padding_mask=None, valid_ratios=None) -> bool:
    """True when the fused kernels (include/msda.h: msda_fused_*) cover this problem AND the tensors are laid out the
    way the kernels index them; the module falls back to the validated reference-shaped path otherwise."""
    if _fused_problem(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                      padding_mask, valid_ratios=valid_ratios) is not None:
        return False
    dims = fused_dims(value, sampling_offsets, reference_points)
    return bool(load().msda_fused_supported(ctypes.byref(dims), _DTYPE_CODE[value.dtype],
                                            int(reference_points.shape[-1])))


def mask_rows_(data, mask):
    """In place: data[mask] = 0 for data (..., R) contiguous and mask (...) bool -- value.masked_fill(mask[..., None], 0)."""
    lib = load()
    row_bytes = data.shape[-1] * data.element_size() if mask.dim() == data.dim() - 1 else \
        data[(0,) * mask.dim()].numel() * data.element_size()
    if not (data.is_cuda and data.is_contiguous() and mask.is_contiguous() and mask.dtype == torch.bool and
            mask.device == data.device and tuple(data.shape[:mask.dim()]) == tuple(mask.shape)):
        raise RuntimeError("mask_rows_ expects contiguous CUDA data (..., R) and a contiguous bool mask (...) on the "
                           "same device")
    with _device_ctx(data.device):
        rc = lib.msda_mask_rows(_ptr(data), _ptr(mask), mask.numel(), row_bytes,
                                ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_mask_rows")
    return data


def fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                  valid_ratios=None, _validated=False):
    lib = load()
    if not _validated:  # (True only from callers that have just run fused_supported on these very tensors)
        why = _fused_problem(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                             valid_ratios=valid_ratios)
        if why:
            raise RuntimeError(why)
    dims = fused_dims(value, sampling_offsets, reference_points)
    out = torch.empty((dims.batch, dims.num_query, dims.num_heads * dims.channels), dtype=value.dtype,
                      device=value.device)
    with _device_ctx(value.device), _nvtx_range("msda_fused_forward"):
        rc = lib.msda_fused_forward_vr(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index),
                                       _ptr(sampling_offsets), _ptr(attn_logits), _ptr(reference_points),
                                       _ptr(valid_ratios) if valid_ratios is not None else None,
                                       int(reference_points.shape[-1]), _ptr(out), ctypes.byref(dims),
                                       _DTYPE_CODE[value.dtype], 0,
                                       ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_fused_forward")
    return out


def fused_backward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                   grad_output, flags: int = 0, valid_ratios=None, _validated=False):
    lib = load()
    if _validated:  # the other tensors were checked in forward (autograd saved them): only grad_output is new
        n, s_, m, d = value.shape
        if not (grad_output.dtype == value.dtype and grad_output.numel() == n * sampling_offsets.shape[1] * m * d and
                grad_output.device == value.device and grad_output.is_contiguous() and grad_output.data_ptr() % 16 == 0):
            raise RuntimeError("grad_output must be a contiguous, 16-byte aligned tensor of value's dtype and device "
                               "with N*Lq*M*D elements")
    else:
        why = _fused_problem(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                             grad_output=grad_output, valid_ratios=valid_ratios)
        if why:
            raise RuntimeError(why)
    dims = fused_dims(value, sampling_offsets, reference_points)
    code = _DTYPE_CODE[value.dtype]
    grad_offs = torch.empty_like(sampling_offsets)
    grad_logits = torch.empty_like(attn_logits)
    ws_bytes = lib.msda_backward_workspace_bytes(ctypes.byref(dims), code, flags)
    if ws_bytes:
        grad_value = torch.empty_like(value)
        flags |= FLAG_ZERO_GRAD_VALUE
    else:
        grad_value = torch.zeros_like(value)
    workspace = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=value.device) if ws_bytes else None
    with _device_ctx(value.device), _nvtx_range("msda_fused_backward"):
        rc = lib.msda_fused_backward_vr(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index),
                                     _ptr(sampling_offsets), _ptr(attn_logits), _ptr(reference_points),
                                     _ptr(valid_ratios) if valid_ratios is not None else None,
                                     int(reference_points.shape[-1]), _ptr(grad_output), _ptr(grad_value),
                                     _ptr(grad_offs), _ptr(grad_logits), ctypes.byref(dims), code, flags,
                                     _ptr(workspace) if workspace is not None else None, ws_bytes,
                                     ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_fused_backward")
    return grad_value, grad_offs, grad_logits


def add_dropout_ln_supported(x, z, weight, bias, keep=None) -> bool:
    """True when msda_add_dropout_ln_* (include/msda.h) can take these tensors as they are."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 2):
        return False
    c = x.shape[-1]
    tensors = [x, z, weight, bias] + ([keep] if keep is not None else [])
    if z.shape != x.shape or z.dtype != x.dtype or weight is None or bias is None:
        return False
    if tuple(weight.shape) != (c,) or tuple(bias.shape) != (c,) or weight.dtype != x.dtype or bias.dtype != x.dtype:
        return False
    if keep is not None and (keep.dtype != torch.bool or keep.shape != x.shape):
        return False
    if any(t.device != x.device or not t.is_contiguous() or t.data_ptr() % 16 for t in tensors):
        return False
    return bool(load().msda_add_dropout_ln_supported(c))


def add_dropout_ln_forward(x, z, keep, keep_scale, weight, bias, eps, save_for_backward):
    """y = LayerNorm(x + keep*keep_scale*z); returns (y, h, mean, rstd) -- the last three None in inference."""
    lib = load()
    c = x.shape[-1]
    rows = x.numel() // c
    y = torch.empty_like(x)
    h = torch.empty_like(x) if save_for_backward else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_for_backward else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_for_backward else None
    opt = lambda t: _ptr(t) if t is not None else None
    with _device_ctx(x.device), _nvtx_range("msda_add_dropout_ln_forward"):
        rc = lib.msda_add_dropout_ln_forward(_ptr(x), _ptr(z), opt(keep), float(keep_scale), _ptr(weight), _ptr(bias),
                                             float(eps), _ptr(y), opt(h), opt(mean), opt(rstd), rows, c,
                                             ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_add_dropout_ln_forward")
    return y, h, mean, rstd


def add_dropout_ln_backward(grad_y, h, mean, rstd, keep, keep_scale, weight):
    """-> (grad_x, grad_z, grad_weight, grad_bias)."""
    lib = load()
    c = h.shape[-1]
    rows = h.numel() // c
    grad_x, grad_z = torch.empty_like(h), torch.empty_like(h)
    grad_w, grad_b = torch.empty_like(weight), torch.empty_like(weight)
    ws_bytes = lib.msda_add_dropout_ln_workspace_bytes(rows, c)
    ws = torch.empty(max(ws_bytes // 4, 4), dtype=torch.float32, device=h.device)
    with _device_ctx(h.device), _nvtx_range("msda_add_dropout_ln_backward"):
        rc = lib.msda_add_dropout_ln_backward(_ptr(grad_y), _ptr(h), _ptr(mean), _ptr(rstd),
                                              _ptr(keep) if keep is not None else None, float(keep_scale),
                                              _ptr(weight), _ptr(grad_x), _ptr(grad_z), _ptr(grad_w), _ptr(grad_b),
                                              _ptr(ws), ws_bytes, rows, c,
                                              ctypes.c_void_p(_raw_stream()))
    if rc:
        _raise(lib, rc, "msda_add_dropout_ln_backward")
    return grad_x, grad_z, grad_w, grad_b


class HostSession:
    """Host-buffer forward+backward (include/msda.h: msda_host_session); bench.py's e2e leg."""

    def __init__(self, max_dims: MsdaDims, dtype: torch.dtype, device: int = 0, images_per_chunk: int = 1):
        self._lib = load()
        self._handle = ctypes.c_void_p()
        self.dtype = dtype
        rc = self._lib.msda_host_session_create(ctypes.byref(self._handle), ctypes.byref(max_dims),
                                                _DTYPE_CODE[dtype], device, images_per_chunk)
        if rc:
            _raise(self._lib, rc, "msda_host_session_create")

    def _call(self, fn, name, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
              output, grad_value, grad_loc, grad_attn, flags):
        for t in (value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, output,
                  grad_value, grad_loc, grad_attn):
            if t.is_cuda or not t.is_contiguous():
                raise RuntimeError("HostSession expects contiguous host tensors")
        if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
            raise RuntimeError("spatial_shapes and level_start_index must be int64 (torch.long) tensors")
        n, s, m, d = value.shape
        _, lq, _, l, p, _ = sampling_loc.shape
        dims = MsdaDims(n, s, m, d, l, lq, p)
        rc = fn(self._handle, _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_loc),
                _ptr(attn_weight), _ptr(grad_output), _ptr(output), _ptr(grad_value), _ptr(grad_loc), _ptr(grad_attn),
                ctypes.byref(dims), flags)
        if rc:
            _raise(self._lib, rc, name)

    def forward_backward(self, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                         output, grad_value, grad_loc, grad_attn, flags: int = 0):
        """All arguments are HOST tensors (ideally pinned); outputs are written in place; returns when they have landed."""
        self._call(self._lib.msda_host_forward_backward, "msda_host_forward_backward", value, spatial_shapes,
                   level_start_index, sampling_loc, attn_weight, grad_output, output, grad_value, grad_loc, grad_attn,
                   flags)

    def submit(self, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
               output, grad_value, grad_loc, grad_attn, flags: int = 0):
        """Enqueue one forward+backward and return at once (msda_host_submit); call wait() before touching the buffers."""
        self._call(self._lib.msda_host_submit, "msda_host_submit", value, spatial_shapes, level_start_index,
                   sampling_loc, attn_weight, grad_output, output, grad_value, grad_loc, grad_attn, flags)

    def wait(self):
        rc = self._lib.msda_host_wait(self._handle)
        if rc:
            _raise(self._lib, rc, "msda_host_wait")

    def close(self):
        if self._handle:
            self._lib.msda_host_session_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
