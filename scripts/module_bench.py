"""GPU: time MSDeformAttn (the module: 4 Linears + softmax + location arithmetic + core op) forward+backward and
break the time down by kernel with torch.profiler.  Development tool.

    python scripts/module_bench.py [--workload detr_encoder_800x1333] [--fused 0|1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from grit_b200 import MSDeformAttn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="detr_encoder_800x1333")
    ap.add_argument("--fused", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--ref-dim", type=int, default=2)
    ap.add_argument("--mask", type=int, default=1)
    args = ap.parse_args()
    cfg = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
    L = len(cfg["shapes"])
    S = sum(h * w for h, w in cfg["shapes"])
    Lq = cfg["Lq"] or S
    C = M * D
    torch.manual_seed(0)
    mod = MSDeformAttn(C, L, M, P).to(dev)
    if hasattr(mod, "fused"):
        mod.fused = bool(args.fused)
    mod.validate_shapes = False
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.01)
        mod.attention_weights.weight.normal_(0, 0.1)
    shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    query = torch.randn(N, Lq, C, device=dev, requires_grad=True)
    src = torch.randn(N, S, C, device=dev, requires_grad=True)
    ref = torch.rand(N, Lq, L, args.ref_dim, device=dev)
    if args.ref_dim == 4:
        ref[..., 2:] = ref[..., 2:] * 0.2 + 0.05
    mask = None
    if args.mask:
        mask = torch.zeros(N, S, dtype=torch.bool, device=dev)
        mask[:, ::10] = True
    gout = torch.randn(N, Lq, C, device=dev)

    def step():
        query.grad = src.grad = None
        for p in mod.parameters():
            p.grad = None
        out = mod(query, ref, src, shapes, lsi, mask)
        out.backward(gout)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"{args.workload} fused={args.fused} ref_dim={args.ref_dim} mask={args.mask}: module fwd+bwd {ms:.3f} ms "
          f"-> {N * Lq / ms / 1e3:.1f} Mq/s; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:22]
    tot = sum(e.device_time_total for e in prof.key_averages())
    for e in rows:
        print(f"  {e.device_time_total / 1e3:8.3f} ms {e.count:3d}x  {e.key[:110]}")
    print(f"  total device time {tot / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
