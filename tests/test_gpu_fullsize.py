"""GPU: element-wise parity at the FULL sizes of BASELINE.json's configs, through the C ABI with the
default (auto) kernel selection -- so the shapes the bench measures are the shapes that are checked:

  config 3  encoder 800x1333, Lq = S = 22223, D = 32     -> fwd_v5, bwd_planes
  config 2  encoder 384x640,  Lq = S = 5100,  D = 32     -> fwd_v5, bwd_planes
  config 4  decoder Lq = 150, D = 64, S = 22223 / 5100   -> fwd_v5, bwd_v5 + msda_bwd_owned (forced here: the auto rule
            picks it for bf16 problems whose grad_value is >= 64 MB, i.e. at the bench batch sizes, not at N = 2)
(the binned backward, the plain row backward and the staged forward, not defaults, are forced in further passes over the
two encoder shapes)

in fp32 and bf16, with uniform and detector-like sampling locations and a padding mask on the right/bottom 10 % of every
level.  N = 2 images: the OpenMP C oracle needs about a second per image at these sizes.  Tolerances as everywhere
(tests/test_gpu_parity.py: fp32 1e-5 forward / 1e-4 gradients, bf16 2e-2).
"""
import numpy as np
import pytest
import torch

from . import helpers
from .test_gpu_parity import assert_parity, lib, oracle_results, run_kernels  # noqa: F401  (lib is a fixture)

pytestmark = pytest.mark.gpu


def detector_like_loc(rng, N, Lq, M, shapes, P, encoder):
    """reference point = own pixel centre (encoder) or U[0,1) (decoder); offsets = the
    module-init ring pattern (p+1) * unit_dir(head) (reference modules/ms_deform_attn.py:58-63) + N(0, 2 px), divided
    by (W_l, H_l) (reference :106-108)."""
    L = len(shapes)
    if encoder:
        refs = []
        for h, w in shapes:
            ys, xs = np.meshgrid(np.arange(h) + 0.5, np.arange(w) + 0.5, indexing="ij")
            refs.append(np.stack([xs.reshape(-1) / w, ys.reshape(-1) / h], -1))
        ref = np.broadcast_to(np.concatenate(refs, 0)[None], (N, Lq, 2))
    else:
        ref = rng.random((N, Lq, 2))
    ang = np.arange(M) * (2.0 * np.pi / M)
    ring = np.stack([np.cos(ang), np.sin(ang)], -1)
    ring = ring / np.abs(ring).max(-1, keepdims=True)
    offs = ring.reshape(1, 1, M, 1, 1, 2) * np.arange(1, P + 1).reshape(1, 1, 1, 1, P, 1)
    offs = offs + 2.0 * rng.standard_normal((N, Lq, M, L, P, 2))
    norm = np.asarray([[w, h] for h, w in shapes], dtype=np.float64).reshape(1, 1, 1, L, 1, 2)
    return ref.reshape(N, Lq, 1, 1, 1, 2) + offs / norm


def padding_mask(N, shapes, frac=0.1):
    """True on the right / bottom `frac` of every level (padding-mask variant)."""
    parts = []
    for h, w in shapes:
        m = np.zeros((h, w), dtype=bool)
        m[int(round(h * (1 - frac))):, :] = True
        m[:, int(round(w * (1 - frac))):] = True
        parts.append(m.reshape(-1))
    return np.broadcast_to(np.concatenate(parts)[None], (N, sum(h * w for h, w in shapes))).copy()


FULL_SIZE = [
    # id,                 pyramid,                   Lq,   D,  expected backward kernel substring
    ("enc800x1333_d32_row", helpers.PYRAMID_800x1333, None, 32, "bwd_v5<"),
    ("enc384x640_d32_row", helpers.PYRAMID_384x640, None, 32, "bwd_v5<"),
    ("enc800x1333_d32_auto", helpers.PYRAMID_800x1333, None, 32, "bwd_planes"),   # levels 2 + 3 on chip
    ("enc384x640_d32_auto", helpers.PYRAMID_384x640, None, 32, "bwd_planes"),    # levels 1 - 3 on chip
    ("enc800x1333_d32_binned", helpers.PYRAMID_800x1333, None, 32, "+binned"),
    ("enc384x640_d32_binned", helpers.PYRAMID_384x640, None, 32, "+binned"),
    ("enc800x1333_d32_planes", helpers.PYRAMID_800x1333, None, 32, "bwd_planes"),
    ("enc384x640_d32_planes", helpers.PYRAMID_384x640, None, 32, "bwd_planes"),
    ("enc384x640_d64_planes", helpers.PYRAMID_384x640, None, 64, "bwd_planes"),
    ("dec800x1333_d64", helpers.PYRAMID_800x1333, 150, 64, "+owned"),
    ("dec384x640_d64", helpers.PYRAMID_384x640, 150, 64, "+owned"),
    ("dec800x1333_d32", helpers.PYRAMID_800x1333, 150, 32, "+owned"),
]


@pytest.mark.parametrize("dist", ["uniform", "detector"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name,shapes,Lq,D,bsub", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_elementwise_vs_oracle(lib, oracle, name, shapes, Lq, D, bsub, dtype, dist):
    N, M, P = 2, 8, 4
    S = sum(h * w for h, w in shapes)
    encoder = Lq is None
    Lq = Lq or S
    case = helpers.make_inputs(N, Lq, M, D, shapes, P, seed=len(name) + D, dtype=np.float32)
    rng = np.random.default_rng(7)
    if dist == "detector":
        case["loc"] = detector_like_loc(rng, N, Lq, M, shapes, P, encoder).astype(np.float32)
    # masked pixels are zero rows of value (reference modules/ms_deform_attn.py:96-97)
    case["value"][padding_mask(N, shapes)] = 0.0
    case = helpers.rounded_case(case, dtype)
    prev = lib.set_tuning("bwd_mode", 2 if name.endswith("_binned") else 4 if name.endswith("_planes") else
                          1 if name.endswith("_row") else 3 if bsub == "+owned" else 0)
    prev_v = lib.set_tuning("variant", 3 if name.endswith("_row") else 0)
    # the forced "_planes" passes use four small CTAs per SM (smallest levels on chip); auto = one 768-thread CTA per SM
    prev_t = lib.set_tuning("planes_threads", 256 if name.endswith("_planes") else 768)
    try:
        got = run_kernels(lib, case, dtype)
    finally:
        lib.set_tuning("bwd_mode", prev)
        lib.set_tuning("variant", prev_v)
        lib.set_tuning("planes_threads", prev_t)
    expect = "fwd_staged" if name.endswith("_row") else "fwd_v5"  # the staged forward is forced in the "_row" passes
    assert got["fwd_kernel"].startswith(expect), got["fwd_kernel"]
    assert bsub in got["bwd_kernel"], got["bwd_kernel"]
    assert_parity(got, oracle_results(oracle, case), case, dtype, f"{name} {dist}")


def test_full_size_backward_strategies_agree(lib):
    """800x1333 encoder shape, fp32: the row-only backward and the row + binned backward produce the same grad_value up
    to fp32 summation order, and bit-identical grad_sampling_loc / grad_attn_weight (same row kernel)."""
    from .test_gpu_parity import run_backward_mode
    from .conftest import max_norm_err
    shapes = helpers.PYRAMID_800x1333
    S = sum(h * w for h, w in shapes)
    case = helpers.make_inputs(2, S, 8, 32, shapes, 4, seed=11, dtype=np.float32)
    row = run_backward_mode(lib, case, torch.float32, "row")
    binned = run_backward_mode(lib, case, torch.float32, "binned")
    assert "+binned" in binned["bwd_kernel"] and "+binned" not in row["bwd_kernel"]
    assert torch.equal(row["grad_loc"], binned["grad_loc"]) and torch.equal(row["grad_attn"], binned["grad_attn"])
    assert max_norm_err(binned["grad_value"].double().cpu().numpy(), row["grad_value"].double().cpu().numpy()) < 2e-5


def test_fused_forward_full_size_error_vs_fp64_matches_the_unfused_path(lib, oracle):
    """800x1333, fp32: the fused forward takes raw offsets / logits / reference points and forms the sampling locations
    with reciprocal multiplies (no IEEE divisions in the kernel); the unfused path gets locations PyTorch computed with
    divisions.  Both hold fp32 locations -- one ulp of a location is ~1e-5 px on a 167-px-wide level -- so both sit a few
    1e-6 from the fp64 truth; the fused path must not be further from it than the unfused one by more than that noise."""
    from .conftest import max_norm_err
    rng = np.random.default_rng(5)
    shapes_l = helpers.PYRAMID_800x1333
    N, M, D, L, P = 1, 8, 32, 4, 4
    S = sum(h * w for h, w in shapes_l)
    Lq = S
    value = rng.standard_normal((N, S, M, D)).astype(np.float32)
    offs = (rng.standard_normal((N, Lq, M, L, P, 2)) * 2.0).astype(np.float32)
    logits = rng.standard_normal((N, Lq, M, L * P)).astype(np.float32)
    ref = (rng.random((N, Lq, L, 2)) * 1.1 - 0.05).astype(np.float32)
    shp = np.asarray(shapes_l, dtype=np.float64)
    norm = np.stack([shp[:, 1], shp[:, 0]], -1)[None, None, None, :, None, :]
    loc64 = ref.astype(np.float64)[:, :, None, :, None, :] + offs.astype(np.float64) / norm
    attn64 = oracle.softmax_np(logits.astype(np.float64), -1).reshape(N, Lq, M, L, P)
    shapes_np = np.asarray(shapes_l, dtype=np.int64)
    lsi = helpers.level_start(shapes_l)
    truth = oracle.forward(value, shapes_np, lsi, loc64, attn64)

    cu = lambda x: torch.from_numpy(x).cuda()
    fused = lib.fused_forward(cu(value), cu(shapes_np), cu(lsi), cu(offs), cu(logits), cu(ref))
    assert lib.last_kernel().startswith("fwd_fused")
    t_ref, t_off = cu(ref), cu(offs)
    loc32 = (t_ref[:, :, None, :, None, :] + t_off / cu(norm.astype(np.float32))).contiguous()
    attn32 = torch.softmax(cu(logits), -1).view(N, Lq, M, L, P).contiguous()
    plain = lib.forward(cu(value), cu(shapes_np), cu(lsi), loc32, attn32)
    e_fused = max_norm_err(fused.double().cpu().numpy().reshape(truth.shape), truth)
    e_plain = max_norm_err(plain.double().cpu().numpy().reshape(truth.shape), truth)
    assert e_fused < 1e-5 and e_plain < 1e-5, (e_fused, e_plain)
    assert e_fused <= 2.0 * e_plain + 1e-6, (e_fused, e_plain)


def test_module_on_a_dense_encoder_shape_takes_the_planes_backward_on_both_paths(lib):
    """384x640 encoder shape through MSDeformAttn: big enough for the auto rule, so the fused path runs msda_bwd_planes with
    the fused point source (softmax + location arithmetic in the kernel) and the reference-shaped path runs it with
    materialised locations / weights; both must give the same gradients (fp32 summation order apart)."""
    from grit_b200 import MSDeformAttn
    from .conftest import max_norm_err
    torch.manual_seed(11)
    shapes_l = helpers.PYRAMID_384x640
    N, M, D, L, P = 2, 8, 32, 4, 4
    C, S = M * D, sum(h * w for h, w in shapes_l)
    shapes = torch.tensor(shapes_l, device="cuda")
    lsi = torch.from_numpy(helpers.level_start(shapes_l)).cuda()
    mod = MSDeformAttn(C, L, M, P).cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.3 / C ** 0.5)
        mod.attention_weights.weight.normal_(0, 1.0 / C ** 0.5)
    query, src = torch.randn(N, S, C, device="cuda"), torch.randn(N, S, C, device="cuda")
    ref_pts = torch.rand(N, S, L, 2, device="cuda")
    gout = torch.randn(N, S, C, device="cuda")
    results = {}
    for fused in (True, False):
        mod.fused = fused
        mod.zero_grad(set_to_none=True)
        q, s_ = query.clone().requires_grad_(True), src.clone().requires_grad_(True)
        out = mod(q, ref_pts, s_, shapes, lsi, None)
        seen = []

        def note_kernel(grad):  # msda_last_kernel is thread-local: read it on the autograd thread that ran the op
            seen.append(lib.last_kernel())
            return grad
        s_.register_hook(note_kernel)
        out.backward(gout)
        kernel = seen[0]
        assert kernel.startswith("bwd_planes_fused" if fused else "bwd_planes<"), kernel
        results[fused] = dict(out=out.detach(), q=q.grad, s=s_.grad, **{k: p.grad.clone() for k, p in mod.named_parameters()})
    for k in results[True]:
        err = max_norm_err(results[True][k].double().cpu().numpy(), results[False][k].double().cpu().numpy())
        assert err < 2e-4, (k, err)
