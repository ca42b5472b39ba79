"""GPU: time the float and the deterministic backward on the headline shape (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from grit_b200 import _lib

cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "detr_encoder_800x1333"]
dev = "cuda"
s = bench.make_layer_inputs(torch, cfg, dev, 0, "uniform")
shapes = torch.tensor(cfg["shapes"], device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
for flags in (0, _lib.FLAG_DETERMINISTIC):
    for i in range(2):
        g = _lib.backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"], flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        g = _lib.backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"], flags)
    e1.record()
    torch.cuda.synchronize()
    print("flags", flags, "backward ms (incl. alloc / zero-fill / pre-pass / fold):", round(e0.elapsed_time(e1) / 5, 3))
