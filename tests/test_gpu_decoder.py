"""GPU: the decoder layer around the hot path: fused residual + dropout + LayerNorm epilogue,
valid-ratio scaling inside the attention kernels, the drop-in DeformableTransformerDecoderLayer against golden vectors
made by the reference's own class, the CUDA-graphed six-layer decoder and the region-feature extraction helper."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from . import helpers
from .conftest import load_golden, max_norm_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from grit_b200 import _lib
    _lib.load()
    return _lib


@pytest.mark.parametrize("C", [128, 256, 384, 512])
@pytest.mark.parametrize("rows", [(3, 150), (1, 7), (2, 1031)])
def test_add_dropout_layer_norm_without_dropout_matches_torch(lib, C, rows):
    """y = LayerNorm(x + z) and all four gradients against torch's own LayerNorm (reference det_module.py:337-339 with
    dropout off).  fp32; tolerance = a few ulp of the normalised values (torch reduces with Welford, this kernel with
    two passes, so the last bit may differ)."""
    from grit_b200 import add_dropout_layer_norm
    torch.manual_seed(C + rows[1])
    norm = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.normal_(1.0, 0.3), norm.bias.normal_(0, 0.3)
    x = torch.randn(*rows, C, device="cuda", requires_grad=True)
    z = torch.randn(*rows, C, device="cuda", requires_grad=True)
    g = torch.randn(*rows, C, device="cuda")
    y = add_dropout_layer_norm(x, z, norm, 0.1, training=False)
    assert lib.last_kernel().startswith("add_dropout_ln_fwd")
    y.backward(g)
    got = [y.detach().clone(), x.grad.clone(), z.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone()]
    x.grad = z.grad = None
    norm.zero_grad(set_to_none=True)
    y_ref = norm(x + z)
    y_ref.backward(g)
    ref = [y_ref.detach(), x.grad, z.grad, norm.weight.grad, norm.bias.grad]
    for a, b, name in zip(got, ref, ("y", "dx", "dz", "dgamma", "dbeta")):
        assert max_norm_err(a.cpu().numpy(), b.cpu().numpy()) < 5e-6, name
    # second run of backward gives identical bits (two-stage fixed-order reduction, no atomics)
    x.grad = z.grad = None
    norm.zero_grad(set_to_none=True)
    add_dropout_layer_norm(x, z, norm, 0.1, training=False).backward(g)
    assert torch.equal(norm.weight.grad, got[3]) and torch.equal(norm.bias.grad, got[4]) and torch.equal(x.grad, got[1])


@pytest.mark.parametrize("C", [256, 512])
def test_add_dropout_layer_norm_seeded_mask_matches_nn_dropout(lib, C):
    """Training mode: with the same RNG state the fused op applies the mask nn.Dropout would have drawn for this tensor
    (reference: tgt + dropout1(tgt2), det_module.py:337), forward and backward."""
    from grit_b200 import add_dropout_layer_norm
    norm = torch.nn.LayerNorm(C).cuda()
    drop = torch.nn.Dropout(0.1)
    x = torch.randn(4, 150, C, device="cuda", requires_grad=True)
    z = torch.randn(4, 150, C, device="cuda", requires_grad=True)
    g = torch.randn(4, 150, C, device="cuda")
    torch.manual_seed(1234)
    y = add_dropout_layer_norm(x, z, norm, drop.p, training=True)
    y.backward(g)
    got = [y.detach().clone(), x.grad.clone(), z.grad.clone(), norm.weight.grad.clone()]
    x.grad = z.grad = None
    norm.zero_grad(set_to_none=True)
    torch.manual_seed(1234)
    y_ref = norm(x + drop(z))
    y_ref.backward(g)
    assert (z.grad == 0).float().mean().item() == pytest.approx(0.1, abs=0.02)  # the mask is really applied
    for a, b, name in zip(got, [y_ref.detach(), x.grad, z.grad, norm.weight.grad], ("y", "dx", "dz", "dgamma")):
        assert max_norm_err(a.cpu().numpy(), b.cpu().numpy()) < 5e-6, name
    assert torch.equal(got[2] == 0, z.grad == 0)  # identical mask, element for element


def test_add_dropout_layer_norm_falls_back_when_unsupported(lib):
    from grit_b200 import add_dropout_layer_norm
    norm = torch.nn.LayerNorm(200).cuda()  # 200 channels: no specialisation
    x, z = torch.randn(5, 200, device="cuda"), torch.randn(5, 200, device="cuda")
    assert torch.allclose(add_dropout_layer_norm(x, z, norm), norm(x + z))
    norm64 = torch.nn.LayerNorm(256).cuda().double()  # fp64: torch path
    x, z = torch.randn(5, 256, device="cuda", dtype=torch.float64), torch.randn(5, 256, device="cuda", dtype=torch.float64)
    assert torch.equal(add_dropout_layer_norm(x, z, norm64), norm64(x + z))


def _layer_from_golden(g, dtype):
    from grit_b200 import DeformableTransformerDecoderLayer
    layer = DeformableTransformerDecoderLayer(int(g["d_model"]), int(g["d_ffn"]), 0.1, "relu", int(g["n_levels"]),
                                              int(g["n_heads"]), int(g["n_points"]))
    missing = layer.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items()
                                     if k.startswith("param.")}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys  # reference checkpoints load key for key
    layer = layer.to("cuda", dtype).eval()
    layer.cross_attn.validate_shapes = False
    return layer


@pytest.mark.parametrize("name", ["decoder_layer_ref2", "decoder_layer_ref4"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_decoder_layer_matches_reference_layer_golden(lib, name, dtype):
    """DeformableTransformerDecoderLayer (fused epilogues + in-kernel valid-ratio scaling in fp32, generic kernels and
    torch epilogues in fp64) against vectors made by the reference's own class (tests/golden/gen_golden_decoder.py,
    det_module.py:313-349), incl. gradients w.r.t. tgt, src and the cross-attention / LayerNorm parameters."""
    g = load_golden(name)
    layer = _layer_from_golden(g, dtype)
    cu = lambda k: torch.from_numpy(g[k]).to("cuda")
    tgt = cu("tgt").to(dtype).requires_grad_(True)
    src = cu("src").to(dtype).requires_grad_(True)
    out = layer(tgt, cu("query_pos").to(dtype), cu("reference_points").to(dtype), src, cu("shapes"), cu("level_start"),
                cu("valid_ratios").to(dtype), cu("padding_mask"))
    out.backward(cu("grad_out").to(dtype))
    tol = 1e-9 if dtype == torch.float64 else 3e-4
    assert max_norm_err(out.detach().cpu().numpy(), g["out"]) < tol
    assert max_norm_err(tgt.grad.cpu().numpy(), g["grad_tgt"]) < tol
    assert max_norm_err(src.grad.cpu().numpy(), g["grad_src"]) < tol
    for k, p in layer.named_parameters():
        if "grad." + k in g:  # parameter gradients are stored rounded to fp32
            assert max_norm_err(p.grad.cpu().numpy(), g["grad." + k]) < max(tol, 2e-7), k


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_valid_ratio_scaling_inside_the_fused_kernels(lib, ref_dim):
    """MSDeformAttn(..., valid_ratios=vr) with un-expanded reference points == the reference's expansion
    reference_points[:, :, None] * valid_ratios[:, None] (det_module.py:323-328) fed to the plain module; gradients
    w.r.t. the reference points included."""
    from grit_b200 import MSDeformAttn
    torch.manual_seed(5 + ref_dim)
    N, Lq, C, M, L, P = 2, 40, 256, 8, 4, 4
    shapes_l = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes_l)
    mod = MSDeformAttn(C, L, M, P).cuda()
    mod.validate_shapes = False
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.2)
    shapes = torch.tensor(shapes_l, device="cuda")
    lsi = torch.from_numpy(helpers.level_start(shapes_l)).cuda()
    query, src = torch.randn(N, Lq, C, device="cuda"), torch.randn(N, S, C, device="cuda")
    ref = torch.rand(N, Lq, 2, device="cuda")
    if ref_dim == 4:
        ref = torch.cat([ref, torch.rand(N, Lq, 2, device="cuda") * 0.4 + 0.05], -1)
    vr = torch.rand(N, L, 2, device="cuda") * 0.4 + 0.6
    gout = torch.randn(N, Lq, C, device="cuda")
    r1 = ref.clone().requires_grad_(True)
    out = mod(query, r1, src, shapes, lsi, valid_ratios=vr)
    assert lib.last_kernel().startswith("fwd_fused")
    out.backward(gout)
    r2 = ref.clone().requires_grad_(True)
    expanded = r2[:, :, None] * (vr if ref_dim == 2 else torch.cat([vr, vr], -1))[:, None]
    out_ref = mod(query, expanded, src, shapes, lsi)
    out_ref.backward(gout)
    assert max_norm_err(out.detach().cpu().numpy(), out_ref.detach().cpu().numpy()) < 2e-6
    assert max_norm_err(r1.grad.cpu().numpy(), r2.grad.cpu().numpy()) < 2e-4
    mod.fused = False  # the unfused path expands like the reference
    out_unfused = mod(query, ref, src, shapes, lsi, valid_ratios=vr)
    assert max_norm_err(out_unfused.detach().cpu().numpy(), out_ref.detach().cpu().numpy()) < 1e-5


def _decoder_problem(n_layers=6, N=4, Lq=150, C=256, M=8, shapes_l=helpers.PYRAMID_384x640):
    from grit_b200 import DeformableTransformerDecoderLayer
    torch.manual_seed(0)
    L, P = 4, 4
    S = sum(h * w for h, w in shapes_l)
    layers = torch.nn.ModuleList([DeformableTransformerDecoderLayer(C, 4 * C, 0.1, "relu", L, M, P)
                                  for _ in range(n_layers)]).cuda().eval()
    for layer in layers:
        layer.cross_attn.validate_shapes = False
        with torch.no_grad():
            layer.cross_attn.sampling_offsets.weight.normal_(0, 0.02)
            layer.cross_attn.attention_weights.weight.normal_(0, 0.2)
    shapes = torch.tensor(shapes_l, device="cuda")
    lsi = torch.from_numpy(helpers.level_start(shapes_l)).cuda()
    args = dict(tgt=torch.randn(N, Lq, C, device="cuda"), query_pos=torch.randn(N, Lq, C, device="cuda"),
                reference_points=torch.rand(N, Lq, 2, device="cuda"), src=torch.randn(N, S, C, device="cuda"),
                valid_ratios=torch.rand(N, L, 2, device="cuda") * 0.3 + 0.7,
                padding_mask=torch.zeros(N, S, dtype=torch.bool, device="cuda"))
    args["padding_mask"][:, ::11] = True
    return layers, args, shapes, lsi


def test_hoisted_and_per_layer_decoder_agree_and_fused_epilogue_equals_reference_launches(lib):
    """run_decoder with one batched value_proj GEMM == every layer projecting for itself; fused epilogues == the
    reference's dropout + add + LayerNorm launches (eval mode), within fp32 rounding."""
    from grit_b200 import run_decoder
    layers, a, shapes, lsi = _decoder_problem(n_layers=3)
    with torch.no_grad():
        hoisted = run_decoder(layers, a["tgt"], a["query_pos"], a["reference_points"], a["src"], shapes, lsi,
                              a["valid_ratios"], a["padding_mask"], hoist_value_proj=True)
        per_layer = run_decoder(layers, a["tgt"], a["query_pos"], a["reference_points"], a["src"], shapes, lsi,
                                a["valid_ratios"], a["padding_mask"], hoist_value_proj=False)
        for layer in layers:
            layer.fused_epilogue = False
            layer.cross_attn.fused = False
        plain = run_decoder(layers, a["tgt"], a["query_pos"], a["reference_points"], a["src"], shapes, lsi,
                            a["valid_ratios"], a["padding_mask"], hoist_value_proj=False)
    assert tuple(hoisted.shape) == (3,) + tuple(a["tgt"].shape)
    assert max_norm_err(hoisted.cpu().numpy(), per_layer.cpu().numpy()) < 1e-5
    assert max_norm_err(hoisted.cpu().numpy(), plain.cpu().numpy()) < 2e-5


def test_graphed_decoder_and_region_feature_extraction(lib):
    """GraphedDecoder: six layers, forward only, one CUDA graph == eager run_decoder; replay with new inputs follows the
    inputs; extract_region_features returns (n_layers, N, Lq, C) fp32 (tools/extract_features.py:80-119)."""
    from grit_b200 import GraphedDecoder, extract_region_features, run_decoder
    layers, a, shapes, lsi = _decoder_problem(n_layers=6)
    with torch.no_grad():
        eager = run_decoder(layers, a["tgt"], a["query_pos"], a["reference_points"], a["src"], shapes, lsi,
                            a["valid_ratios"], a["padding_mask"])
    graphed = GraphedDecoder(layers, a["tgt"], a["query_pos"], a["reference_points"], a["src"], shapes, lsi,
                             a["valid_ratios"], a["padding_mask"])
    out = graphed(a["tgt"], a["query_pos"], a["reference_points"], a["src"], a["valid_ratios"], a["padding_mask"])
    assert max_norm_err(out.cpu().numpy(), eager.cpu().numpy()) < 1e-5
    b = {k: (torch.randn_like(v) if v.dtype.is_floating_point and k not in ("reference_points", "valid_ratios") else v)
         for k, v in a.items()}
    with torch.no_grad():
        eager_b = run_decoder(layers, b["tgt"], b["query_pos"], b["reference_points"], b["src"], shapes, lsi,
                              b["valid_ratios"], b["padding_mask"])
    feats = extract_region_features(layers, b["tgt"], b["query_pos"], b["reference_points"], b["src"], shapes, lsi,
                                    b["valid_ratios"], b["padding_mask"], graphed=graphed)
    assert feats.dtype == torch.float32 and tuple(feats.shape) == (6,) + tuple(a["tgt"].shape)
    assert max_norm_err(feats.cpu().numpy(), eager_b.cpu().numpy()) < 1e-5
    feats_eager = extract_region_features(layers, b["tgt"], b["query_pos"], b["reference_points"], b["src"], shapes, lsi,
                                          b["valid_ratios"], b["padding_mask"])
    assert max_norm_err(feats_eager.cpu().numpy(), eager_b.cpu().numpy()) < 1e-6


def test_graphed_training_decoder_matches_eager_gradients(lib):
    """graphed_training_decoder: forward and backward of the decoder stack replayed from CUDA graphs (GRIT trains at batch
    4, where the step is host-bound) == the eager stack: outputs, input gradients and parameter gradients, in eval mode
    (no dropout) so both are deterministic functions of the inputs; then a replay with new inputs follows the inputs."""
    import copy

    from grit_b200 import graphed_training_decoder, run_decoder
    layers, a, shapes, lsi = _decoder_problem(n_layers=3)
    for layer in layers:
        layer.eval()
    eager_layers = copy.deepcopy(layers)

    def step(fn, owner, tgt, src):
        for p_ in owner.parameters():
            p_.grad = None
        t, s_ = tgt.clone().requires_grad_(True), src.clone().requires_grad_(True)
        out = fn(t, a["query_pos"], a["reference_points"], s_, shapes, lsi, a["valid_ratios"], a["padding_mask"])
        out[-1].square().sum().backward()
        return out.detach().clone(), t.grad.clone(), s_.grad.clone(), [p_.grad.clone() for p_ in owner.parameters()]

    sample_t, sample_s = a["tgt"].clone().requires_grad_(True), a["src"].clone().requires_grad_(True)
    graphed = graphed_training_decoder(layers, sample_t, a["query_pos"], a["reference_points"], sample_s, shapes, lsi,
                                       a["valid_ratios"], a["padding_mask"])
    eager = lambda *args: run_decoder(eager_layers, *args)
    for scale in (1.0, 0.5):  # second pass: new inputs through the same graphs
        tgt, src = a["tgt"] * scale, a["src"] * scale + 0.1
        got = step(graphed, layers, tgt, src)
        ref = step(eager, eager_layers, tgt, src)
        assert max_norm_err(got[0].cpu().numpy(), ref[0].cpu().numpy()) < 1e-5
        assert max_norm_err(got[1].cpu().numpy(), ref[1].cpu().numpy()) < 2e-4
        assert max_norm_err(got[2].cpu().numpy(), ref[2].cpu().numpy()) < 2e-4
        for g_, r_ in zip(got[3], ref[3]):
            assert max_norm_err(g_.cpu().numpy(), r_.cpu().numpy()) < 5e-4


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,shapes_l", [(256, [(12, 20), (6, 10), (3, 5), (2, 3)]), (512, [(13, 21), (7, 11), (4, 6), (2, 3)])])
def test_groupnorm_epilogue_writes_packed_memory(lib, out_dtype, C, shapes_l):
    """GroupNorm(32, C) of every level's conv output written straight into the (N, S, C) memory layout
    == the reference's GroupNorm (detector.py:39-44, 64) + flatten/transpose/cat (det_module.py:146-155), forward and
    backward (gradients w.r.t. the conv outputs and every level's GroupNorm weight / bias)."""
    from grit_b200 import pack_levels_groupnorm
    torch.manual_seed(C)
    N = 3
    gns = [torch.nn.GroupNorm(32, C).cuda() for _ in shapes_l]
    for gn in gns:
        with torch.no_grad():
            gn.weight.normal_(1.0, 0.3), gn.bias.normal_(0, 0.3)
    xs = [(torch.randn(N, C, h, w, device="cuda") * 2 + 5).requires_grad_(True) for h, w in shapes_l]  # |mean| >> std
    memory, spatial_shapes, lsi = pack_levels_groupnorm(xs, gns, out_dtype)
    assert lib.last_kernel().startswith("pack_levels_gn")
    ref = torch.cat([gn(x).flatten(2).transpose(1, 2) for x, gn in zip(xs, gns)], 1)
    assert memory.dtype == out_dtype and tuple(memory.shape) == tuple(ref.shape)
    assert spatial_shapes.tolist() == [list(s) for s in shapes_l] and int(lsi[1]) == shapes_l[0][0] * shapes_l[0][1]
    tol = 1e-5 if out_dtype == torch.float32 else 1e-2
    assert max_norm_err(memory.float().detach().cpu().numpy(), ref.detach().cpu().numpy()) < tol
    g = torch.randn_like(ref)
    memory.backward(g.to(out_dtype))
    got = [x.grad.clone() for x in xs] + [gn.weight.grad.clone() for gn in gns] + [gn.bias.grad.clone() for gn in gns]
    for x in xs:
        x.grad = None
    for gn in gns:
        gn.zero_grad(set_to_none=True)
    ref.backward(g.to(out_dtype).float())
    want = [x.grad for x in xs] + [gn.weight.grad for gn in gns] + [gn.bias.grad for gn in gns]
    for a, b in zip(got, want):
        assert max_norm_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-4
