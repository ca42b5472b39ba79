"""GPU A/B of the staged forward's scheduling knobs against the row kernel, all in one process on one GPU.
    python scripts/staged_ab.py [--workload detr_encoder_800x1333] [--reps 3]"""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from grit_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="detr_encoder_800x1333")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--configs", default="all", help="all | row (row-kernel knobs only)")
args = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda", 0)
cfg = dict(bench.WORKLOADS[args.workload])
N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
L = len(cfg["shapes"]); S = sum(h * w for h, w in cfg["shapes"]); Lq = cfg["Lq"] or S
sets = [bench.make_layer_inputs(torch, cfg, dev, i, "uniform") for i in range(3)]  # rotate inputs >> L2
shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
dt = sets[0]["value"].dtype
out = torch.empty(N, Lq, M * D, device=dev, dtype=dt)
dims = _lib.MsdaDims(N, S, M, D, L, Lq, P); code = _lib._DTYPE_CODE[dt]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream); P_ = _lib._ptr

def fwd(x):
    rc = lib.msda_forward(P_(x["value"]), P_(shapes), P_(lsi), P_(x["loc"]), P_(x["attn"]), P_(out), ctypes.byref(dims), code, 0, st)
    assert rc == 0, lib.msda_last_error()

def timeit():
    for i in range(3): fwd(sets[i % 3])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.iters): fwd(sets[i % 3])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / args.iters

configs = [("row_v5", {"variant": 5})] + \
    [(f"staged_rows{r}", {"variant": 3, "staged_rows": r, "staged_persistent": 0}) for r in (256, 512, 1024, 2048, 4445)] + \
    [(f"staged_persistent_rows{r}", {"variant": 3, "staged_rows": r, "staged_persistent": 1}) for r in (512, 1024, 4445)] + \
    [("staged_768thr_rows1024", {"variant": 3, "staged_rows": 1024, "v3_threads": 768}),
     ("staged_512thr_rows1024", {"variant": 3, "staged_rows": 1024, "v3_threads": 512})]
if args.configs == "row":
    configs = [("row_v5", {"variant": 5}), ("row_v5_w8", {"variant": 5, "warps": 8}),
               ("row_v5_hoist", {"variant": 5, "hoist": 1})]
res = {}
for rep in range(args.reps):
    for name, knobs in configs:
        saved = {k: _lib.set_tuning(k, v) for k, v in knobs.items()}
        try:
            res.setdefault(name, []).append(round(timeit(), 4))
        finally:
            for k, v in saved.items(): _lib.set_tuning(k, v)
for k, v in res.items(): print(f"{k:32s} {v}")
print(json.dumps(res))
