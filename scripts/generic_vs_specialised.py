import sys, torch
sys.path.insert(0, ".")
import bench
from grit_b200 import _lib
cfg = dict(bench.WORKLOADS["detr_encoder_800x1333"]); cfg["N"] = 4
dev = "cuda"
s = bench.make_layer_inputs(torch, cfg, dev, 0, "uniform")
shapes = torch.tensor(cfg["shapes"], device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / 5
for flags, name in ((0, "specialised"), (_lib.FLAG_FORCE_GENERIC, "generic")):
    f = t(lambda: _lib.forward(s["value"], shapes, lsi, s["loc"], s["attn"], flags))
    b = t(lambda: _lib.backward(s["value"], shapes, lsi, s["loc"], s["attn"], s["gout"], flags))
    print(name, "N=4 fwd ms", round(f, 3), "bwd ms", round(b, 3))
