"""GPU experiment: the forward is bound by the L2->SM gather path, the backward by the SM->L2 reduction path.  Do a
forward and a backward of INDEPENDENT inputs overlap when issued on two streams?"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from grit_b200 import _lib

cfg = bench.WORKLOADS["detr_encoder_800x1333"]
dev = torch.device("cuda:0")
lib = _lib.load()
N, M, D, P = cfg["N"], cfg["M"], cfg["D"], cfg["P"]
L = len(cfg["shapes"])
S = sum(h * w for h, w in cfg["shapes"])
Lq = S
shapes = torch.tensor(cfg["shapes"], dtype=torch.int64, device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
sets = [bench.make_layer_inputs(torch, cfg, dev, i, "uniform") for i in range(4)]
out = [torch.empty(N, Lq, M * D, device=dev) for _ in range(2)]
gv = torch.zeros(N, S, M, D, device=dev)
gl = torch.empty(N, Lq, M, L, P, 2, device=dev)
ga = torch.empty(N, Lq, M, L, P, device=dev)
dims = _lib.MsdaDims(N, S, M, D, L, Lq, P)
p = _lib._ptr


def fwd(s, o, st):
    assert lib.msda_forward(p(s["value"]), p(shapes), p(lsi), p(s["loc"]), p(s["attn"]), p(o), ctypes.byref(dims), 0, 0,
                            ctypes.c_void_p(st.cuda_stream)) == 0


def bwd(s, st):
    assert lib.msda_backward(p(s["value"]), p(shapes), p(lsi), p(s["loc"]), p(s["attn"]), p(s["gout"]), p(gv), p(gl),
                             p(ga), ctypes.byref(dims), 0, 0, None, 0, ctypes.c_void_p(st.cuda_stream)) == 0


sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
iters = 12


def run(overlap):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_stream(torch.cuda.current_stream())
    sb.wait_stream(torch.cuda.current_stream())
    for i in range(iters):
        fwd(sets[i % 4], out[i % 2], sa)
        bwd(sets[(i + 1) % 4], sb if overlap else sa)
    torch.cuda.current_stream().wait_stream(sa)
    torch.cuda.current_stream().wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for ov in (False, True, False, True):
    ms = run(ov)
    print(f"overlap={ov}: {ms:.3f} ms per (forward + backward) pair -> {N * Lq / ms / 1e3:.1f} Mq/s")
