"""GPU: a pure C++ program (no Python, no torch in the process) links libmsda_b200.so through include/msda.h, runs
forward/backward on cudaMalloc'ed buffers and checks them against the C oracle -- the boundary as a non-Python host
of the reference would use it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_client_of_the_c_abi(tmp_path):
    from grit_b200 import build
    from oracle import msda_oracle
    build.build_library()
    msda_oracle.build()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "c_abi_parity")
    lib_dir, oracle_dir = os.path.join(ROOT, "grit_b200"), os.path.join(ROOT, "oracle")
    cmd = [nvcc, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "c_abi", "c_abi_parity.cu"), "-L", lib_dir, "-lmsda_b200", "-L", oracle_dir,
           "-lmsda_oracle", "-Xlinker", f"-rpath={lib_dir}", "-Xlinker", f"-rpath={oracle_dir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(proc.stdout)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "C ABI PARITY OK" in proc.stdout and "fwd_v5" in proc.stdout
