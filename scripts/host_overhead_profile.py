"""GPU: where the HOST time of the eager fused decoder goes at GRIT's training batch (4): cProfile over run_decoder.
    python scripts/host_overhead_profile.py"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from grit_b200 import DeformableTransformerDecoderLayer, run_decoder

torch.manual_seed(0)
shapes_l = [(48, 80), (24, 40), (12, 20), (6, 10)]
C, M, L, P, Lq, N = 512, 8, 4, 4, 150, 4
S = sum(h * w for h, w in shapes_l)
dev = "cuda"
layers = torch.nn.ModuleList([DeformableTransformerDecoderLayer(C, 1024, 0.1, "relu", L, M, P) for _ in range(6)]).to(dev)
for layer in layers:
    layer.cross_attn.validate_shapes = False
shapes = torch.tensor(shapes_l, device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
tgt, pos = torch.randn(N, Lq, C, device=dev), torch.randn(N, Lq, C, device=dev)
ref, src = torch.rand(N, Lq, 2, device=dev), torch.randn(N, S, C, device=dev)
vr = torch.rand(N, L, 2, device=dev) * 0.2 + 0.8
mask = torch.zeros(N, S, dtype=torch.bool, device=dev); mask[:, ::10] = True


def step():
    for p_ in layers.parameters():
        p_.grad = None
    t = tgt.clone().requires_grad_(True)
    run_decoder(layers, t, pos, ref, src, shapes, lsi, vr, mask)[-1].sum().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
