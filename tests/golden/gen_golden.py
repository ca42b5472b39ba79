"""Generate tests/golden/*.npz from the UNMODIFIED reference, run in the build container only.

    python tests/golden/gen_golden.py            # needs /root/reference (absent on the GPU box)

The reference's oracle ``ms_deform_attn_core_pytorch`` and its ``MSDeformAttn`` module are loaded
straight from /root/reference/models/ops by file path.  The reference module imports the compiled
extension ``MultiScaleDeformableAttention`` at import time (functions/ms_deform_attn_func.py:18);
no GPU exists here, so a stub module is registered whose forward routes to the reference's own
pure-PyTorch oracle -- every number written below is therefore computed by reference code.

Cases
  testpy_*        the reference test's recipe (models/ops/test.py:21-36, seed 3, same draw order)
  oob_small       out-of-range locations, a 1x1 level, odd sizes (N=2, M=3, D=5, L=3, P=2)
  det_small       detector-like 4-level pyramid, M=8, D=32, padding-style zeros
  module_ref2/4   MSDeformAttn.forward with 2-d / 4-d reference points and a padding mask,
                  incl. gradients w.r.t. inputs and all eight parameters
Only the committed .npz files travel; nothing in tests/ or bench.py reads /root/reference.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_OPS = "/root/reference/models/ops"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def _load_reference():
    stub = types.ModuleType("MultiScaleDeformableAttention")
    sys.modules["MultiScaleDeformableAttention"] = stub

    def load(name, path, is_pkg=False):
        spec = importlib.util.spec_from_file_location(
            name, path, submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg = types.ModuleType("refops")
    pkg.__path__ = [REF_OPS]
    sys.modules["refops"] = pkg
    load("refops.functions", os.path.join(REF_OPS, "functions", "__init__.py"), is_pkg=True)
    func = sys.modules["refops.functions.ms_deform_attn_func"]
    load("refops.modules", os.path.join(REF_OPS, "modules", "__init__.py"), is_pkg=True)
    modl = sys.modules["refops.modules.ms_deform_attn"]

    core = func.ms_deform_attn_core_pytorch

    class _OracleFunction:  # stands in for the CUDA-backed autograd Function on this GPU-less box
        @staticmethod
        def apply(value, shapes, level_start, loc, attn, im2col_step):
            return core(value, shapes, loc, attn)

    modl.MSDeformAttnFunction = _OracleFunction
    return core, modl.MSDeformAttn


def _level_start(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def _grads(core, value, shapes, loc, attn, gout):
    v = value.clone().requires_grad_(True)
    s = loc.clone().requires_grad_(True)
    a = attn.clone().requires_grad_(True)
    out = core(v, shapes, s, a)
    out.backward(gout)
    return out.detach(), v.grad, s.grad, a.grad


def _save(name, **arrays):
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path)} B)")


def main():
    core, MSDeformAttn = _load_reference()

    # ---- the reference test's own recipe, same RNG stream order as models/ops/test.py -------------
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = _level_start(shapes)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)

    def draw(channels):
        value = torch.rand(N, S, M, channels) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
        return value, loc, attn

    value, loc, attn = draw(D)  # check_forward_equal_with_pytorch_double
    _save("testpy_fwd_double", value=value, loc=loc, attn=attn, shapes=shapes, level_start=lsi,
          out=core(value.double(), shapes, loc.double(), attn.double()))
    value, loc, attn = draw(D)  # check_forward_equal_with_pytorch_float
    _save("testpy_fwd_float", value=value, loc=loc, attn=attn, shapes=shapes, level_start=lsi,
          out=core(value, shapes, loc, attn))
    gen = torch.Generator().manual_seed(1234)
    for channels in (30, 32, 64, 71):  # check_gradient_numerical(channels)
        value, loc, attn = draw(channels)
        gout = torch.randn(N, Lq, M * channels, generator=gen, dtype=torch.float64)
        out, gv, gl, ga = _grads(core, value.double(), shapes, loc.double(), attn.double(), gout)
        _save(f"testpy_grad_D{channels}", value=value, loc=loc, attn=attn, shapes=shapes, level_start=lsi,
              grad_out=gout, out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)

    # ---- out-of-range / degenerate-level case -----------------------------------------------------
    gen = torch.Generator().manual_seed(7)
    N, M, D, Lq, L, P = 2, 3, 5, 7, 3, 2
    shapes = torch.as_tensor([(5, 7), (1, 1), (2, 3)], dtype=torch.long)
    lsi = _level_start(shapes)
    S = int(shapes.prod(1).sum())
    value = torch.randn(N, S, M, D, generator=gen, dtype=torch.float64)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=gen, dtype=torch.float64) * 1.6 - 0.3
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=gen, dtype=torch.float64), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=gen, dtype=torch.float64)
    out, gv, gl, ga = _grads(core, value, shapes, loc, attn, gout)
    _save("oob_small", value=value, loc=loc, attn=attn, shapes=shapes, level_start=lsi,
          grad_out=gout, out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)

    # ---- detector-like pyramid (D=32, 8 heads, 4x4 points) ----------------------------------------
    gen = torch.Generator().manual_seed(11)
    N, M, D, Lq, L, P = 2, 8, 32, 24, 4, 4
    shapes = torch.as_tensor([(8, 12), (4, 6), (2, 3), (1, 2)], dtype=torch.long)
    lsi = _level_start(shapes)
    S = int(shapes.prod(1).sum())
    value = torch.randn(N, S, M, D, generator=gen, dtype=torch.float64)
    value[:, ::9] = 0  # padding-style zeroed pixels
    loc = torch.rand(N, Lq, M, L, P, 2, generator=gen, dtype=torch.float64) * 1.1 - 0.05
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=gen, dtype=torch.float64), -1).view(N, Lq, M, L, P)
    gout = torch.randn(N, Lq, M * D, generator=gen, dtype=torch.float64)
    # inputs are stored as fp32; the reference is evaluated in fp64 on the rounded values so the
    # stored outputs correspond exactly to the stored inputs
    v32, l32, a32, g32 = value.float().double(), loc.float().double(), attn.float().double(), gout.float().double()
    out, gv, gl, ga = _grads(core, v32, shapes, l32, a32, g32)
    _save("det_small", value=v32.float(), loc=l32.float(), attn=a32.float(), shapes=shapes, level_start=lsi,
          grad_out=g32.float(), out=out, grad_value=gv, grad_loc=gl, grad_attn=ga)

    # ---- the MSDeformAttn module (reference modules/ms_deform_attn.py:73-119) ---------------------
    for ref_dim in (2, 4):
        torch.manual_seed(100 + ref_dim)
        C, M, L, P, N, Lq = 32, 4, 3, 2, 2, 9
        shapes = torch.as_tensor([(6, 8), (3, 4), (2, 2)], dtype=torch.long)
        lsi = _level_start(shapes)
        S = int(shapes.prod(1).sum())
        mod = MSDeformAttn(d_model=C, n_levels=L, n_heads=M, n_points=P).double()
        init_state = {k: v.clone() for k, v in mod.state_dict().items()}
        with torch.no_grad():  # move off the all-zero init so every path carries signal
            mod.sampling_offsets.weight.normal_(0, 0.05)
            mod.attention_weights.weight.normal_(0, 0.3)
            mod.attention_weights.bias.normal_(0, 0.3)
            mod.value_proj.bias.normal_(0, 0.1)
            mod.output_proj.bias.normal_(0, 0.1)
        query = torch.randn(N, Lq, C, dtype=torch.float64, requires_grad=True)
        src = torch.randn(N, S, C, dtype=torch.float64, requires_grad=True)
        if ref_dim == 2:
            ref = torch.rand(N, Lq, L, 2, dtype=torch.float64)
        else:
            ref = torch.cat([torch.rand(N, Lq, L, 2, dtype=torch.float64),
                             torch.rand(N, Lq, L, 2, dtype=torch.float64) * 0.5 + 0.05], -1)
        mask = torch.zeros(N, S, dtype=torch.bool)
        mask[1, ::5] = True
        out = mod(query, ref, src, shapes, lsi, mask)
        gout = torch.randn_like(out)
        out.backward(gout)
        arrays = dict(query=query, reference_points=ref, input_flatten=src, shapes=shapes, level_start=lsi,
                      padding_mask=mask, out=out, grad_out=gout, grad_query=query.grad, grad_input_flatten=src.grad,
                      n_heads=M, n_levels=L, n_points=P, d_model=C)
        for k, v in mod.state_dict().items():
            arrays["param." + k] = v
            arrays["init." + k] = init_state[k]
        for k, p in mod.named_parameters():
            arrays["grad." + k] = p.grad
        _save(f"module_ref{ref_dim}", **arrays)


if __name__ == "__main__":
    main()
