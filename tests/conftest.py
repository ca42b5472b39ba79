"""Shared pytest plumbing: the ``gpu`` marker, repo-root imports, golden-vector loading."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import msda_oracle
    msda_oracle.build()
    return msda_oracle


def max_norm_err(got, ref):
    """max |got-ref| / max |ref|  -- the parity metric of BASELINE.md section 4."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = max(float(np.abs(ref).max()), 1e-300)
    return float(np.abs(got - ref).max()) / denom


def l2_rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(got - ref)) / max(float(np.linalg.norm(ref)), 1e-300)
