"""CPU: property tests of the oracle itself (hypothesis) -- the C restatement against the independent numpy one over
random geometries, plus the algebraic identities the GPU tests rely on at sizes the oracle cannot reach."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from . import helpers
from .conftest import max_norm_err

geometry = st.tuples(
    st.integers(1, 3),                                   # N
    st.integers(0, 9),                                   # Lq (0 = empty query set)
    st.integers(1, 4),                                   # M
    st.integers(1, 9),                                   # D
    st.lists(st.tuples(st.integers(1, 7), st.integers(1, 7)), min_size=1, max_size=4),  # level shapes
    st.integers(1, 3),                                   # P
    st.integers(0, 10_000),                              # seed
)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(geometry)
def test_c_and_numpy_restatements_agree(oracle, geo):
    N, Lq, M, D, shapes, P, seed = geo
    case = helpers.make_inputs(N, Lq, M, D, shapes, P, seed=seed, lo=-0.4, hi=1.4)
    out_c = oracle.forward(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"])
    out_np = oracle.forward_numpy(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"])
    assert out_c.shape == (N, Lq, M * D)
    if out_c.size:
        assert np.abs(out_c - out_np).max() <= 1e-12 * max(1.0, np.abs(out_np).max())


@settings(max_examples=40, deadline=None, derandomize=True)
@given(geometry)
def test_adjoint_identities_hold_for_the_oracle(oracle, geo):
    """<out, g> == <value, grad_value> == <attn, grad_attn>: the forward is linear in value and in attn."""
    N, Lq, M, D, shapes, P, seed = geo
    case = helpers.make_inputs(N, Lq, M, D, shapes, P, seed=seed, lo=-0.2, hi=1.2)
    out = oracle.forward(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"])
    gv, gl, ga = oracle.backward(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"],
                                 case["grad_out"])
    lhs = float((out * case["grad_out"]).sum())
    scale = float(np.abs(out * case["grad_out"]).sum()) + 1e-30
    assert abs(lhs - float((case["value"] * gv).sum())) <= 1e-10 * scale + 1e-12
    assert abs(lhs - float((case["attn"] * ga).sum())) <= 1e-10 * scale + 1e-12


@settings(max_examples=25, deadline=None, derandomize=True)
@given(geometry)
def test_location_gradient_matches_finite_differences(oracle, geo):
    N, Lq, M, D, shapes, P, seed = geo
    if Lq == 0:
        return
    case = helpers.make_inputs(N, Lq, M, D, shapes, P, seed=seed, lo=0.05, hi=0.95)
    ties = helpers.tie_mask(case["loc"], case["shapes"], eps=1e-3)
    _, gl, _ = oracle.backward(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"],
                               case["grad_out"])
    rng = np.random.default_rng(seed)
    idx = tuple(int(rng.integers(0, n)) for n in case["loc"].shape)
    if ties[idx[:-1]]:
        return
    eps = 1e-6
    f = lambda loc: float((oracle.forward(case["value"], case["shapes"], case["level_start"], loc, case["attn"]) *
                           case["grad_out"]).sum())
    lp, lm = case["loc"].copy(), case["loc"].copy()
    lp[idx] += eps
    lm[idx] -= eps
    fd = (f(lp) - f(lm)) / (2 * eps)
    assert abs(fd - gl[idx]) <= 1e-5 * max(1.0, abs(gl[idx]))


def test_batch_sharding_is_exact(oracle):
    """Images are independent (what bench.py --gpus N relies on): per-shard results equal the full-batch slices."""
    case = helpers.make_inputs(5, 11, 3, 6, [(4, 5), (2, 3)], 2, seed=4)
    full = oracle.forward(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"])
    gfull = oracle.backward(case["value"], case["shapes"], case["level_start"], case["loc"], case["attn"],
                            case["grad_out"])
    for sl in (slice(0, 2), slice(2, 5)):
        part = oracle.forward(case["value"][sl], case["shapes"], case["level_start"], case["loc"][sl], case["attn"][sl])
        assert np.array_equal(part, full[sl])
        gpart = oracle.backward(case["value"][sl], case["shapes"], case["level_start"], case["loc"][sl],
                                case["attn"][sl], case["grad_out"][sl])
        for a, b in zip(gpart, gfull):
            assert np.array_equal(a, b[sl])
