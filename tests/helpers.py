"""Input builders and comparison helpers shared by the CPU and GPU tests."""
import numpy as np
import torch

PYRAMID_384x640 = [(48, 80), (24, 40), (12, 20), (6, 10)]      # S = 5100   
PYRAMID_800x1333 = [(100, 167), (50, 84), (25, 42), (13, 21)]  # S = 22223


def level_start(shapes):
    sizes = [h * w for h, w in shapes]
    return np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)


def make_inputs(N, Lq, M, D, shapes, P, seed=0, lo=-0.05, hi=1.05, dtype=np.float64):
    """Seeded synthetic problem: value ~ N(0,1), loc ~ U[lo,hi), attn = softmax(N(0,1)), grad_out ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    L = len(shapes)
    S = int(sum(h * w for h, w in shapes))
    value = rng.standard_normal((N, S, M, D))
    loc = rng.random((N, Lq, M, L, P, 2)) * (hi - lo) + lo
    logits = rng.standard_normal((N, Lq, M, L * P))
    e = np.exp(logits - logits.max(-1, keepdims=True))
    attn = (e / e.sum(-1, keepdims=True)).reshape(N, Lq, M, L, P)
    gout = rng.standard_normal((N, Lq, M * D))
    return dict(value=value.astype(dtype), loc=loc.astype(dtype), attn=attn.astype(dtype), grad_out=gout.astype(dtype),
                shapes=np.asarray(shapes, dtype=np.int64).reshape(L, 2), level_start=level_start(shapes))


def tie_mask(loc, shapes, eps=2e-4):
    """True where a sample sits within `eps` pixels of an integer pixel coordinate: floor() may legitimately resolve
    differently in fp32 and fp64 there and grad_sampling_loc is discontinuous, so those points are not compared."""
    loc = np.asarray(loc, dtype=np.float64)
    mask = np.zeros(loc.shape[:-1], dtype=bool)
    for l, (h, w) in enumerate(np.asarray(shapes).tolist()):
        x = loc[:, :, :, l, :, 0] * w - 0.5
        y = loc[:, :, :, l, :, 1] * h - 0.5
        mask[:, :, :, l] = (np.abs(x - np.round(x)) < eps) | (np.abs(y - np.round(y)) < eps)
    return mask


def to_cuda(case, dtype, device="cuda"):
    """numpy case dict -> torch CUDA tensors; loc/attn follow the library's dtype rule (fp32 unless fp64)."""
    loc_dtype = torch.float64 if dtype == torch.float64 else torch.float32
    t = {}
    for k in ("value", "grad_out"):
        if k in case:
            t[k] = torch.from_numpy(np.ascontiguousarray(case[k])).to(device=device, dtype=dtype)
    for k in ("loc", "attn"):
        t[k] = torch.from_numpy(np.ascontiguousarray(case[k])).to(device=device, dtype=loc_dtype)
    t["shapes"] = torch.from_numpy(case["shapes"]).to(device)
    t["level_start"] = torch.from_numpy(case["level_start"]).to(device)
    return t


def rounded_case(case, dtype):
    """The case as the kernel sees it after casting (so the fp64 oracle runs on identical numbers)."""
    loc_dtype = torch.float64 if dtype == torch.float64 else torch.float32
    out = dict(case)
    for k in ("value", "grad_out"):
        if k in case:
            out[k] = torch.from_numpy(np.asarray(case[k])).to(dtype).double().numpy()
    for k in ("loc", "attn"):
        out[k] = torch.from_numpy(np.asarray(case[k])).to(loc_dtype).double().numpy()
    return out
