"""Build the REFERENCE's own CUDA extension for sm_100a into baseline/_ref/ (git-ignored; travels with gpurun).

Used only as the on-GPU comparison kernel ("the 2020 scalar-gather kernels recompiled for B200") by
`scripts/ref_cuda_bench.py`; never imported by grit_b200/ or the tests.  Reference sources are NOT copied into the repo:
they are copied to a temp dir, the two lines that no longer compile against torch >= 2.x
(models/ops/src/cuda/ms_deform_attn_cuda.cu:64,134  value.type() -> value.scalar_type()) are patched there, and only the
built .so lands in baseline/_ref/.  Needs /root/reference, i.e. runs in the build container only.
"""
import glob
import os
import shutil
import sys
import tempfile

REF_SRC = "/root/reference/models/ops/src"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main():
    if not os.path.isdir(REF_SRC):
        print("reference sources not present; nothing to build")
        return 0
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    tmp = tempfile.mkdtemp(prefix="msda_ref_")
    src = os.path.join(tmp, "src")
    shutil.copytree(REF_SRC, src)
    cu = os.path.join(src, "cuda", "ms_deform_attn_cuda.cu")
    text = open(cu).read().replace("AT_DISPATCH_FLOATING_TYPES(value.type(),", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type(),")
    open(cu, "w").write(text)
    os.makedirs(OUT_DIR, exist_ok=True)
    sources = glob.glob(os.path.join(src, "*.cpp")) + glob.glob(os.path.join(src, "cpu", "*.cpp")) + \
        glob.glob(os.path.join(src, "cuda", "*.cu"))
    load(name="MultiScaleDeformableAttentionRef", sources=sources, extra_include_paths=[src],
         extra_cflags=["-DWITH_CUDA"], extra_cuda_cflags=["-DWITH_CUDA", "-O3"], build_directory=OUT_DIR,
         is_python_module=False, verbose=False)
    for f in glob.glob(os.path.join(OUT_DIR, "*")):
        if not f.endswith(".so"):
            (shutil.rmtree if os.path.isdir(f) else os.remove)(f)
    shutil.rmtree(tmp, ignore_errors=True)
    print("built", glob.glob(os.path.join(OUT_DIR, "*.so")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
