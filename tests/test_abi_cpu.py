"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/msda.h declares;
host-side argument checks mirror the reference's error behaviour.  No compute call is made here."""
import ctypes
import os
import re

import pytest
import torch

import grit_b200
from grit_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "msda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert {"msda_forward", "msda_backward", "msda_last_error", "msda_backward_workspace_bytes",
            "msda_host_forward_backward"} <= set(names)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/msda.h but not exported"
    assert lib.msda_abi_version() == 1


def test_library_contains_sm100a_code(lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_invalid_arguments_are_reported_not_crashed(lib):
    dims = _lib.MsdaDims(1, 4, 1, 4, 1, 1, 1)
    rc = lib.msda_forward(None, None, None, None, None, None, ctypes.byref(dims), 0, 0, None)
    assert rc == 1 and b"null" in lib.msda_last_error()
    rc = lib.msda_forward(None, None, None, None, None, None, ctypes.byref(dims), 7, 0, None)
    assert rc == 1 and b"dtype" in lib.msda_last_error()
    bad = _lib.MsdaDims(1, 4, 0, 4, 1, 1, 1)
    assert lib.msda_forward(None, None, None, None, None, None, ctypes.byref(bad), 0, 0, None) == 1
    assert lib.msda_backward_workspace_bytes(ctypes.byref(dims), _lib.MSDA_BF16, 0) == 1 * 4 * 1 * 4 * 4
    assert lib.msda_backward_workspace_bytes(ctypes.byref(dims), _lib.MSDA_F32, 0) == 0


def test_cpu_tensors_are_rejected_like_the_reference():
    # reference: AT_ERROR("Not implemented on the CPU")  (models/ops/src/ms_deform_attn.h:38,60)
    shapes = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    value = torch.zeros(1, 4, 1, 4)
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    attn = torch.ones(1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        grit_b200.MSDeformAttnFunction.apply(value, shapes, lsi, loc, attn, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        grit_b200.MSDeformAttn(8, 1, 2, 1)(torch.zeros(1, 1, 8), torch.zeros(1, 1, 1, 2), torch.zeros(1, 4, 8),
                                            shapes, lsi)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_module_api_surface_matches_reference():
    with pytest.raises(ValueError, match="d_model must be divisible by n_heads"):
        grit_b200.MSDeformAttn(d_model=30, n_heads=8)
    with pytest.warns(UserWarning, match="power of 2"):
        m = grit_b200.MSDeformAttn(d_model=24, n_heads=4, n_levels=2, n_points=3)
    assert m.im2col_step == 64
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {
        "sampling_offsets.weight": (4 * 2 * 3 * 2, 24), "sampling_offsets.bias": (48,),
        "attention_weights.weight": (4 * 2 * 3, 24), "attention_weights.bias": (24,),
        "value_proj.weight": (24, 24), "value_proj.bias": (24,),
        "output_proj.weight": (24, 24), "output_proj.bias": (24,)}
    assert float(m.sampling_offsets.weight.detach().abs().max()) == 0
    assert float(m.attention_weights.bias.detach().abs().max()) == 0


def test_module_init_matches_reference_golden():
    from .conftest import load_golden
    g = load_golden("module_ref4")
    torch.manual_seed(104)  # same seed as tests/golden/gen_golden.py -> identical xavier draws
    m = grit_b200.MSDeformAttn(int(g["d_model"]), int(g["n_levels"]), int(g["n_heads"]), int(g["n_points"]))
    for k, v in m.state_dict().items():
        assert torch.allclose(v.double(), torch.from_numpy(g["init." + k]).double(), atol=1e-7), k


def test_install_as_reference_ops_registers_names():
    import sys
    mod = grit_b200.install_as_reference_ops()
    assert sys.modules["MultiScaleDeformableAttention"] is mod
    assert hasattr(mod, "ms_deform_attn_forward") and hasattr(mod, "ms_deform_attn_backward")
    from models.ops.modules import MSDeformAttn  # noqa: F401  (what GRIT's det_module.py imports)
    from models.ops.functions import MSDeformAttnFunction  # noqa: F401


def test_debug_core_pytorch_matches_golden():
    """The exported ms_deform_attn_core_pytorch (debug helper, not on the product path) agrees with the reference's."""
    from .conftest import load_golden, max_norm_err
    for name in ("oob_small", "testpy_grad_D30"):
        g = load_golden(name)
        out = grit_b200.ms_deform_attn_core_pytorch(torch.from_numpy(g["value"]).double(), torch.from_numpy(g["shapes"]),
                                                    torch.from_numpy(g["loc"]).double(),
                                                    torch.from_numpy(g["attn"]).double())
        assert max_norm_err(out.numpy(), g["out"]) < 1e-12


def test_every_tuning_knob_is_documented_in_the_header():
    """msda_set_tuning's keys (grit_b200/csrc/msda_capi.cu) are part of the public surface: include/msda.h documents each."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "grit_b200", "csrc", "msda_capi.cu")).read()
    hdr = open(os.path.join(root, "include", "msda.h")).read()
    keys = re.findall(r'!strcmp\(key, "([a-z0-9_]+)"\)', src)
    assert len(keys) >= 10
    missing = [k for k in keys if f'"{k}"' not in hdr]
    assert not missing, f"undocumented msda_set_tuning keys: {missing}"
