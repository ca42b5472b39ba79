"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

Replaces the reference's models/ops/setup.py:23-60 (a torch CUDAExtension that refuses to build
without a visible GPU): here the product is a plain C-ABI ``libmsda_b200.so`` with no torch or
pybind dependency, so one nvcc invocation is the whole build.
"""
import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
LIB_NAME = "libmsda_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
SOURCES = [os.path.join(PKG_DIR, "csrc", "msda_capi.cu")]
HEADERS = sorted(glob.glob(os.path.join(PKG_DIR, "csrc", "*.cuh"))) + [os.path.join(REPO_DIR, "include", "msda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "177",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build " + LIB_NAME)


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile grit_b200/libmsda_b200.so if missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-I", os.path.join(REPO_DIR, "include"), "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
