"""GPU: time the REFERENCE's own CUDA kernels (baseline/_ref, built by baseline/build_ref_cuda.py) next to grit_b200 on a
bench workload, and compare the two outputs element-wise at full size.  Writes gpurun_out/r2_ref_cuda_<workload>[_N<batch>].json.

    python scripts/ref_cuda_bench.py [--workload detr_encoder_800x1333] [--iters 10]
"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from grit_b200 import _lib


def load_ref():
    path = os.path.join(ROOT, "baseline", "_ref", "MultiScaleDeformableAttentionRef.so")
    spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttentionRef", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def timed(fn, iters, pre=None):
    for _ in range(3):
        if pre:
            pre()
        fn()
    tot = 0.0
    for _ in range(iters):
        if pre:
            pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="detr_encoder_800x1333")
    ap.add_argument("--loc-dist", default="uniform")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=None, help="override the workload's images per GPU")
    args = ap.parse_args()
    cfg = dict(bench.WORKLOADS[args.workload])
    if args.batch:
        cfg["N"] = args.batch
    if cfg["dtype"] != "f32":
        cfg["dtype"] = "f32"  # the reference kernels are float/double only
    dev = torch.device("cuda:0")
    ref = load_ref()
    shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    s = bench.make_layer_inputs(torch, cfg, dev, 0, args.loc_dist)
    N, Lq = s["loc"].shape[0], s["loc"].shape[1]
    step = 64

    out_ref = ref.ms_deform_attn_forward(s["value"], shapes, lsi, s["loc"], s["attn"], step)
    g_ref = ref.ms_deform_attn_backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"], step)
    out = _lib.forward(s["value"], shapes, lsi, s["loc"], s["attn"])
    kf = _lib.last_kernel()
    g = _lib.backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"])
    kb = _lib.last_kernel()
    torch.cuda.synchronize()
    nerr = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    agree = {"out": nerr(out, out_ref), "grad_value": nerr(g[0], g_ref[0]), "grad_attn": nerr(g[2], g_ref[2])}
    # grad_sampling_loc: exclude samples within 2e-4 px of an integer coordinate (floor ties, see tests/helpers.py)
    W = shapes[:, 1].float().view(1, 1, 1, -1, 1)
    H = shapes[:, 0].float().view(1, 1, 1, -1, 1)
    x, y = s["loc"][..., 0] * W - 0.5, s["loc"][..., 1] * H - 0.5
    keep = ((x - x.round()).abs() > 2e-4) & ((y - y.round()).abs() > 2e-4)
    d = (g[1] - g_ref[1]).abs() * keep.unsqueeze(-1)
    agree["grad_sampling_loc"] = float(d.max() / g_ref[1].abs().max())
    del out_ref, g_ref, out, g

    t = {}
    t["ref_fwd_ms"] = timed(lambda: ref.ms_deform_attn_forward(s["value"], shapes, lsi, s["loc"], s["attn"], step), args.iters)
    t["ref_bwd_ms"] = timed(lambda: ref.ms_deform_attn_backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"], step), args.iters)
    t["b200_fwd_ms"] = timed(lambda: _lib.forward(s["value"], shapes, lsi, s["loc"], s["attn"]), args.iters)
    t["b200_bwd_ms"] = timed(lambda: _lib.backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"]), args.iters)
    res = {"workload": args.workload, "loc_dist": args.loc_dist, "dtype": "f32", "N": N, "Lq": Lq,
           "note": "both sides timed through their Python entry points incl. output allocation and zero-fills "
                   "(reference: at::zeros x4; grit_b200: torch.empty x3 + zeros x1)",
           **{k: round(v, 4) for k, v in t.items()},
           "ref_queries_per_s": N * Lq / ((t["ref_fwd_ms"] + t["ref_bwd_ms"]) * 1e-3),
           "b200_queries_per_s": N * Lq / ((t["b200_fwd_ms"] + t["b200_bwd_ms"]) * 1e-3),
           "speedup_fwd": t["ref_fwd_ms"] / t["b200_fwd_ms"], "speedup_bwd": t["ref_bwd_ms"] / t["b200_bwd_ms"],
           "speedup_fwd_bwd": (t["ref_fwd_ms"] + t["ref_bwd_ms"]) / (t["b200_fwd_ms"] + t["b200_bwd_ms"]),
           "kernels": [kf, kb], "max_norm_diff_vs_reference_kernel": agree}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = f"_N{args.batch}" if args.batch else ""
    with open(os.path.join(ROOT, "gpurun_out", f"r2_ref_cuda_{args.workload}{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
