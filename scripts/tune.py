"""GPU: time kernel variants (msda_set_tuning knobs) on a bench workload.  Development tool, not a benchmark of record.

    python scripts/tune.py [--workload detr_encoder_800x1333] [--loc-dist uniform] [--iters 10]
"""
import argparse
import ctypes
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from grit_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="detr_encoder_800x1333")
    ap.add_argument("--loc-dist", default="uniform")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--sets", type=int, default=3)
    ap.add_argument("--configs", default="2:0:4,3:0:512,3:0:1024")
    args = ap.parse_args()
    cfg = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
    L = len(cfg["shapes"])
    S = sum(h * w for h, w in cfg["shapes"])
    Lq = cfg["Lq"] or S
    dt = {"f32": torch.float32, "bf16": torch.bfloat16}[cfg["dtype"]]
    ev = 4 if dt == torch.float32 else 2
    shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    sets = [bench.make_layer_inputs(torch, cfg, dev, i, args.loc_dist) for i in range(args.sets)]
    fwd_b, bwd_b = bench.algorithmic_bytes(N, S, Lq, M, D, L, P, ev)
    peak, _ = bench.hbm_peak()
    ref = None
    print(f"workload {args.workload} loc={args.loc_dist} N={N} S={S} Lq={Lq} D={D} dtype={cfg['dtype']}")
    print(f"{'variant:hm:warps':18s} {'fwd ms':>8s} {'frac':>6s} {'bwd ms':>8s} {'frac':>6s} {'Mq/s f+b':>9s}  kernels / max err vs first config")
    for spec in args.configs.split(","):
        variant, hm, warps = (int(x) for x in spec.split(":"))
        _lib.set_tuning("variant", variant), _lib.set_tuning("head_major", hm), _lib.set_tuning("hoist", hm)
        _lib.set_tuning("v3_threads" if variant == 3 else "warps", warps)
        s = sets[0]
        out = _lib.forward(s["value"], shapes, lsi, s["loc"], s["attn"])
        kf = _lib.last_kernel()
        gv, gl, ga = _lib.backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"])
        kb = _lib.last_kernel()
        torch.cuda.synchronize()
        res = [out.float(), gv.float(), gl, ga]
        if ref is None:
            ref = res
            err = "reference"
        else:
            err = " ".join(f"{float((a - b).abs().max() / b.abs().max()):.1e}" for a, b in zip(res, ref))
        del out, gv, gl, ga
        out = torch.empty(N, Lq, M * D, device=dev, dtype=dt)
        gvb = torch.empty(N, S, M, D, device=dev, dtype=dt)
        glb = torch.empty(N, Lq, M, L, P, 2, device=dev)
        gab = torch.empty(N, Lq, M, L, P, device=dev)
        dims = _lib.MsdaDims(N, S, M, D, L, Lq, P)
        code = _lib._DTYPE_CODE[dt]
        wsb = lib.msda_backward_workspace_bytes(ctypes.byref(dims), code, 0)
        ws = torch.empty(max(wsb // 4, 4), dtype=torch.float32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = _lib._ptr

        def run(kind, s):
            if kind == "fwd":
                rc = lib.msda_forward(p(s["value"]), p(shapes), p(lsi), p(s["loc"]), p(s["attn"]), p(out),
                                      ctypes.byref(dims), code, 0, st)
            else:
                rc = lib.msda_backward(p(s["value"]), p(shapes), p(lsi), p(s["loc"]), p(s["attn"]), p(s["gout"]),
                                       p(gvb), p(glb), p(gab), ctypes.byref(dims), code,
                                       _lib.FLAG_ZERO_GRAD_VALUE if dt == torch.bfloat16 else 0, p(ws), wsb, st)
            assert rc == 0, lib.msda_last_error()

        times = {}
        for kind in ("fwd", "bwd"):
            for i in range(3):
                run(kind, sets[i % len(sets)])
            tot = 0.0
            for i in range(args.iters):
                if kind == "bwd" and dt != torch.bfloat16:
                    gvb.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(kind, sets[i % len(sets)])
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            times[kind] = tot / args.iters
        f, b = times["fwd"], times["bwd"]
        print(f"{spec:18s} {f:8.3f} {fwd_b / f / 1e6 / peak:6.3f} {b:8.3f} {bwd_b / b / 1e6 / peak:6.3f} "
              f"{N * Lq / (f + b) / 1e3:9.1f}  {kf} | {kb} | {err}")
        del out, gvb, glb, gab, ws


if __name__ == "__main__":
    main()
