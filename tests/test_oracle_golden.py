"""CPU: pin the oracle (C, numpy and torch restatements) to the reference-generated golden vectors.

The golden files were written by tests/golden/gen_golden.py from the reference's own
ms_deform_attn_core_pytorch (models/ops/functions/ms_deform_attn_func.py:41-61) + autograd.
"""
import numpy as np
import pytest
import torch

from .conftest import load_golden, max_norm_err

CORE_CASES = ["testpy_grad_D30", "testpy_grad_D32", "testpy_grad_D64", "testpy_grad_D71", "oob_small", "det_small"]


@pytest.mark.parametrize("name", ["testpy_fwd_double", "testpy_fwd_float"] + CORE_CASES)
def test_c_oracle_forward_matches_reference(oracle, name):
    g = load_golden(name)
    out = oracle.forward(g["value"], g["shapes"], g["level_start"], g["loc"], g["attn"])
    tol = 1e-5 if name == "testpy_fwd_float" else 1e-12  # the float golden was computed in fp32
    assert max_norm_err(out, g["out"]) < tol
    # the reference test's own acceptance rule (models/ops/test.py:40,56)
    if name == "testpy_fwd_double":
        assert np.allclose(out, g["out"])
    if name == "testpy_fwd_float":
        assert np.allclose(out, g["out"], rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("name", CORE_CASES)
def test_c_oracle_backward_matches_reference_autograd(oracle, name):
    g = load_golden(name)
    gv, gl, ga = oracle.backward(g["value"], g["shapes"], g["level_start"], g["loc"], g["attn"], g["grad_out"])
    assert max_norm_err(gv, g["grad_value"]) < 1e-12
    assert max_norm_err(gl, g["grad_loc"]) < 1e-12
    assert max_norm_err(ga, g["grad_attn"]) < 1e-12


@pytest.mark.parametrize("name", CORE_CASES)
def test_c_oracle_fp32_build_close_to_fp64(oracle, name):
    g = load_golden(name)
    out = oracle.forward(g["value"], g["shapes"], g["level_start"], g["loc"], g["attn"], dtype=np.float32)
    gv, gl, ga = oracle.backward(g["value"], g["shapes"], g["level_start"], g["loc"], g["attn"], g["grad_out"],
                                 dtype=np.float32)
    assert max_norm_err(out, g["out"]) < 1e-5
    for got, key in ((gv, "grad_value"), (gl, "grad_loc"), (ga, "grad_attn")):
        assert max_norm_err(got, g[key]) < 1e-4


@pytest.mark.parametrize("name", CORE_CASES)
def test_numpy_restatement_matches_reference(oracle, name):
    g = load_golden(name)
    out = oracle.forward_numpy(g["value"], g["shapes"], g["level_start"], g["loc"], g["attn"])
    assert max_norm_err(out, g["out"]) < 1e-12


@pytest.mark.parametrize("name", CORE_CASES)
def test_torch_cpu_restatement_matches_reference(name):
    from oracle import msda_ref_torch
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k]).double()
    out, gv, gl, ga = msda_ref_torch.forward_backward(t("value"), g["shapes"].tolist(), t("loc"), t("attn"),
                                                      t("grad_out"))
    assert max_norm_err(out.numpy(), g["out"]) < 1e-12
    assert max_norm_err(gv.numpy(), g["grad_value"]) < 1e-12
    assert max_norm_err(gl.numpy(), g["grad_loc"]) < 1e-12
    assert max_norm_err(ga.numpy(), g["grad_attn"]) < 1e-12


@pytest.mark.parametrize("name", ["module_ref2", "module_ref4"])
def test_module_oracle_matches_reference_module(oracle, name):
    g = load_golden(name)
    params = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    out = oracle.module_forward_numpy(params, g["query"], g["reference_points"], g["input_flatten"], g["shapes"],
                                      g["level_start"], g["padding_mask"], n_heads=int(g["n_heads"]),
                                      n_levels=int(g["n_levels"]), n_points=int(g["n_points"]))
    assert max_norm_err(out, g["out"]) < 1e-12


def test_oracle_edge_cases(oracle):
    """NaN / far-outside coordinates are skipped; empty query set; single-pixel level."""
    shapes = np.array([[1, 1], [2, 3]], dtype=np.int64)
    lsi = np.array([0, 1], dtype=np.int64)
    rng = np.random.default_rng(0)
    value = rng.standard_normal((1, 7, 2, 3))
    loc = rng.random((1, 4, 2, 2, 2, 2))
    loc[0, 0] = np.nan
    loc[0, 1] = 5.0
    loc[0, 2] = -5.0
    attn = rng.random((1, 4, 2, 2, 2))
    out = oracle.forward(value, shapes, lsi, loc, attn).reshape(1, 4, 2, 3)
    assert np.all(out[0, :3] == 0) and np.all(np.isfinite(out)) and np.any(out[0, 3] != 0)
    gv, gl, ga = oracle.backward(value, shapes, lsi, loc, attn, np.ones((1, 4, 6)))
    assert np.all(gl[0, :3] == 0) and np.all(ga[0, :3] == 0) and np.all(np.isfinite(gv))
    empty = oracle.forward(value, shapes, lsi, loc[:, :0], attn[:, :0])
    assert empty.shape == (1, 0, 6)
