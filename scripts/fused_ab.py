"""GPU A/B: the fused module kernels (softmax + location arithmetic inside the gather) against the plain kernels on the SAME
sampling problem -- logits / offsets / reference points drawn once, locations and attention weights derived from them in
PyTorch for the plain kernels.  800x1333 encoder shape, N=16, fp32, D=32.
    python scripts/fused_ab.py [--out gpurun_out/r2_fused_ab.json]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grit_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--shape", default="800x1333", choices=["800x1333", "384x640"])
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
args = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
shapes_l = [(100, 167), (50, 84), (25, 42), (13, 21)] if args.shape == "800x1333" else [(48, 80), (24, 40), (12, 20), (6, 10)]
N, M, D, L, P = args.n or (16 if args.shape == "800x1333" else 32), 8, 32, 4, 4
vdt = torch.float32 if args.dtype == "f32" else torch.bfloat16
S = sum(h * w for h, w in shapes_l); Lq = S
shapes = torch.tensor(shapes_l, dtype=torch.int64, device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
value = torch.randn(N, S, M, D, device=dev).to(vdt)
logits = torch.randn(N, Lq, M, L * P, device=dev)
ref = torch.rand(N, Lq, L, 2, device=dev) * 1.1 - 0.05          # uniform reference points: worst-case locality
offs = torch.randn(N, Lq, M, L, P, 2, device=dev) * 2.0
gout = torch.randn(N, Lq, M * D, device=dev).to(vdt)
norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
loc = (ref[:, :, None, :, None, :] + offs / norm[None, None, None, :, None, :]).contiguous()
attn = torch.softmax(logits, -1).view(N, Lq, M, L, P).contiguous()

def timeit(fn):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / args.iters

res = {}
prev = _lib.set_tuning("variant", 5)
res["fwd_plain_ms"] = timeit(lambda: _lib.forward(value, shapes, lsi, loc, attn)); res["fwd_plain_kernel"] = _lib.last_kernel()
_lib.set_tuning("variant", prev)
res["fwd_fused_ms"] = timeit(lambda: _lib.fused_forward(value, shapes, lsi, offs, logits, ref)); res["fwd_fused_kernel"] = _lib.last_kernel()
for tag, mode in (("", 1), ("_planes", 4)):  # 1 = row-style kernels, 4 = planes (the default for this dense shape)
    prev = _lib.set_tuning("bwd_mode", mode)
    res[f"bwd_plain{tag}_ms"] = timeit(lambda: _lib.backward(value, shapes, lsi, loc, attn, gout)); res[f"bwd_plain{tag}_kernel"] = _lib.last_kernel()
    res[f"bwd_fused{tag}_ms"] = timeit(lambda: _lib.fused_backward(value, shapes, lsi, offs, logits, ref, gout)); res[f"bwd_fused{tag}_kernel"] = _lib.last_kernel()
    _lib.set_tuning("bwd_mode", prev)
# what the fusion removes at this shape: softmax + location arithmetic as separate PyTorch kernels (forward only here)
res["pre_op_torch_ms"] = timeit(lambda: ((ref[:, :, None, :, None, :] + offs / norm[None, None, None, :, None, :]), torch.softmax(logits, -1)))
a = _lib.forward(value, shapes, lsi, loc, attn); b = _lib.fused_forward(value, shapes, lsi, offs, logits, ref)
res["fwd_max_diff"] = float((a.float() - b.float()).abs().max() / a.float().abs().max())
print(json.dumps(res))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
