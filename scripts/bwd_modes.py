"""GPU: time the backward strategies (msda_set_tuning "bwd_mode": 1 row reds | 2 row + binned coarse levels | 3 owned)
and the two forward variants on the bench workloads.  Development tool, not a benchmark of record.

    python scripts/bwd_modes.py [--workloads a,b,...] [--iters 10] [--loc-dist uniform|detector]

Every timing includes what the strategy needs around the kernels (grad_value zero-fill, bf16 workspace zero-fill and
fold), i.e. msda_backward with MSDA_FLAG_ZERO_GRAD_VALUE on preallocated buffers, CUDA events, inputs resident in HBM.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from grit_b200 import _lib  # noqa: E402

EXTRA = {
    "grit_decoder_384x640_f32": dict(N=64, shapes=[(48, 80), (24, 40), (12, 20), (6, 10)], Lq=150, M=8, D=64, P=4,
                                     dtype="f32", layers=6),
    "grit_decoder_800x1333_f32": dict(N=32, shapes=[(100, 167), (50, 84), (25, 42), (13, 21)], Lq=150, M=8, D=64, P=4,
                                      dtype="f32", layers=6),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16,"
                                           "grit_decoder_384x640_f32,grit_decoder_384x640_bf16,"
                                           "grit_decoder_800x1333_f32,grit_decoder_800x1333_bf16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--loc-dist", default="uniform")
    ap.add_argument("--modes", default="1,2,3")
    ap.add_argument("--out", default=None)
    ap.add_argument("--tuning", default="", help="extra msda_set_tuning knobs for the whole run: key=value,key=value")
    ap.add_argument("--skip-fwd", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    for kv in filter(None, args.tuning.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(k, int(v))
    dev = torch.device("cuda", 0)
    peak, _ = bench.hbm_peak()
    results = {}
    for name in args.workloads.split(","):
        cfg = dict(bench.WORKLOADS.get(name) or EXTRA[name])
        N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
        L = len(cfg["shapes"])
        S = sum(h * w for h, w in cfg["shapes"])
        Lq = cfg["Lq"] or S
        dt = {"f32": torch.float32, "bf16": torch.bfloat16}[cfg["dtype"]]
        ev = 4 if dt == torch.float32 else 2
        x = bench.make_layer_inputs(torch, cfg, dev, 1, args.loc_dist)
        shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        out = torch.empty(N, Lq, M * D, device=dev, dtype=dt)
        gv = torch.empty(N, S, M, D, device=dev, dtype=dt)
        gl = torch.empty(N, Lq, M, L, P, 2, device=dev)
        ga = torch.empty(N, Lq, M, L, P, device=dev)
        dims = _lib.MsdaDims(N, S, M, D, L, Lq, P)
        code = _lib._DTYPE_CODE[dt]
        wsb = lib.msda_backward_workspace_bytes(ctypes.byref(dims), code, 0)
        ws = torch.empty(max(wsb // 4, 4), dtype=torch.float32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P_ = _lib._ptr
        fb, bb = bench.algorithmic_bytes(N, S, Lq, M, D, L, P, ev)

        def fwd():
            rc = lib.msda_forward(P_(x["value"]), P_(shapes), P_(lsi), P_(x["loc"]), P_(x["attn"]), P_(out),
                                  ctypes.byref(dims), code, 0, st)
            assert rc == 0, lib.msda_last_error()

        def bwd():
            rc = lib.msda_backward(P_(x["value"]), P_(shapes), P_(lsi), P_(x["loc"]), P_(x["attn"]), P_(x["gout"]),
                                   P_(gv), P_(gl), P_(ga), ctypes.byref(dims), code, _lib.FLAG_ZERO_GRAD_VALUE, P_(ws),
                                   wsb, st)
            assert rc == 0, lib.msda_last_error()

        def timeit(fn):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / args.iters

        res = {"N": N, "Lq": Lq, "S": S, "D": D, "dtype": cfg["dtype"], "fwd_alg_MB": fb / 1e6, "bwd_alg_MB": bb / 1e6}
        for variant in (() if args.skip_fwd else (0, 5, 3)):
            prev = _lib.set_tuning("variant", variant)
            try:
                ms = timeit(fwd)
                res[f"fwd_variant{variant}"] = {"ms": ms, "kernel": _lib.last_kernel(), "hbm_frac": fb / ms / 1e6 / peak}
            finally:
                _lib.set_tuning("variant", prev)
        ref_gv = None
        for mode in [int(m) for m in args.modes.split(",")]:
            prev = _lib.set_tuning("bwd_mode", mode)
            try:
                ms = timeit(bwd)
                kern = _lib.last_kernel()
                entry = {"ms": ms, "kernel": kern, "hbm_frac": bb / ms / 1e6 / peak}
                if ref_gv is None:
                    ref_gv = gv.float().clone()
                else:
                    entry["max_err_vs_first_mode"] = float((gv.float() - ref_gv).abs().max() / ref_gv.abs().max())
                res[f"bwd_mode{mode}"] = entry
            finally:
                _lib.set_tuning("bwd_mode", prev)
        ms = timeit(bwd)
        res["bwd_auto"] = {"ms": ms, "kernel": _lib.last_kernel(), "hbm_frac": bb / ms / 1e6 / peak}
        results[name] = res
        print(name, json.dumps(res), flush=True)
        del x, out, gv, gl, ga, ws
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
