"""Generate tests/golden/decoder_layer_ref{2,4}.npz from the UNMODIFIED reference decoder layer (build container only).

    python tests/golden/gen_golden_decoder.py            # needs /root/reference

The reference's ``DeformableTransformerDecoderLayer`` (models/detection/det_module.py:272-349) is loaded by file path.
Its module-level imports that are absent here are satisfied with stand-ins that the layer's arithmetic never reaches
with drop_path = 0: ``timm.models.layers.DropPath`` and ``utils.misc.inverse_sigmoid``; ``models.ops.modules.MSDeformAttn``
is the reference's own module routed to the reference's own pure-PyTorch oracle exactly as in gen_golden.py.  Every
number written is therefore computed by reference code, in fp64, eval mode (dropout off), on fp32-rounded parameters and
inputs (which are what is stored).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden  # noqa: E402

OUT_DIR = os.path.dirname(os.path.abspath(__file__))
REF_DET = "/root/reference/models/detection/det_module.py"


def load_reference_layer():
    _, ref_msda = gen_golden._load_reference()
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):  # never instantiated with drop_path = 0 (det_module.py:301)
        def __init__(self, p=0.):
            super().__init__()

    timm_layers.DropPath = DropPath
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    utils = types.ModuleType("utils")
    misc = types.ModuleType("utils.misc")
    misc.inverse_sigmoid = lambda x, eps=1e-5: torch.log(x.clamp(eps, 1 - eps) / (1 - x).clamp(eps, 1 - eps))
    sys.modules.update({"utils": utils, "utils.misc": misc})
    models = types.ModuleType("models")
    ops = types.ModuleType("models.ops")
    mods = types.ModuleType("models.ops.modules")
    mods.MSDeformAttn = ref_msda
    sys.modules.update({"models": models, "models.ops": ops, "models.ops.modules": mods})
    spec = importlib.util.spec_from_file_location("ref_det_module", REF_DET)
    det = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(det)
    return det.DeformableTransformerDecoderLayer


def main():
    Layer = load_reference_layer()
    for ref_dim in (2, 4):
        torch.manual_seed(200 + ref_dim)
        C, F_, M, L, P, N, Lq = 128, 64, 4, 4, 4, 2, 11
        shapes = torch.as_tensor([(10, 14), (5, 7), (3, 4), (2, 2)], dtype=torch.long)
        lsi = gen_golden._level_start(shapes)
        S = int(shapes.prod(1).sum())
        layer = Layer(d_model=C, d_ffn=F_, dropout=0.1, activation="relu", n_levels=L, n_heads=M, n_points=P).eval()
        with torch.no_grad():
            layer.cross_attn.sampling_offsets.weight.normal_(0, 0.05)
            layer.cross_attn.attention_weights.weight.normal_(0, 0.3)
            layer.cross_attn.attention_weights.bias.normal_(0, 0.3)
            for norm in (layer.norm1, layer.norm2, layer.norm3):
                norm.weight.normal_(1.0, 0.2)
                norm.bias.normal_(0, 0.2)
        layer = layer.float().double()  # fp32-rounded parameters, fp64 arithmetic
        r32 = lambda t: t.float().double()
        tgt = r32(torch.randn(N, Lq, C)).requires_grad_(True)
        pos = r32(torch.randn(N, Lq, C))
        src = r32(torch.randn(N, S, C)).requires_grad_(True)
        ref = r32(torch.rand(N, Lq, 2)) if ref_dim == 2 else \
            r32(torch.cat([torch.rand(N, Lq, 2), torch.rand(N, Lq, 2) * 0.5 + 0.05], -1))
        vr = r32(torch.rand(N, L, 2) * 0.4 + 0.6)
        mask = torch.zeros(N, S, dtype=torch.bool)
        mask[1, ::6] = True
        out = layer(tgt, pos, ref, src, shapes, lsi, vr, mask)
        gout = r32(torch.randn_like(out))
        out.backward(gout)
        arrays = dict(tgt=tgt.float(), query_pos=pos.float(), src=src.float(), reference_points=ref.float(),
                      valid_ratios=vr.float(), shapes=shapes, level_start=lsi, padding_mask=mask, grad_out=gout.float(),
                      out=out, grad_tgt=tgt.grad, grad_src=src.grad, d_model=C, d_ffn=F_, n_heads=M, n_levels=L,
                      n_points=P)
        for k, v in layer.state_dict().items():
            arrays["param." + k] = v.float()
        for k, p in layer.named_parameters():
            if k.startswith(("cross_attn.", "norm")):
                arrays["grad." + k] = p.grad.float()
        gen_golden._save(f"decoder_layer_ref{ref_dim}", **arrays)


if __name__ == "__main__":
    main()
