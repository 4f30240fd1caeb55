"""Compile the reference's own 3D deformable-conv extension (D3D) for sm_100a -- TEST INFRASTRUCTURE ONLY.

Recipe (no reference source enters the repo):
  1. copy /root/reference/3D/dcn/src/** to a temporary directory,
  2. apply the two-token patch torch>=2 needs (``input.type()`` -> ``input.scalar_type()`` inside
     AT_DISPATCH_FLOATING_TYPES, deform_conv_cuda.cu:96 and :233),
  3. build with torch.utils.cpp_extension into oracle/_ref/ (git-ignored, travels to the GPU box).

The result ``oracle/_ref/D3D.so`` exposes the reference's pybind11 module ``D3D``
(3D/dcn/src/vision.cpp:4-7).  It is CUDA-only, so it can only RUN on the GPU box, where
tests/test_ref_d3d_gpu.py uses it to pin the oracle and the product against the real reference.
"""
import glob
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/3D/dcn/src"


def built_path():
    c = glob.glob(os.path.join(OUT, "D3D*.so"))
    return c[0] if c else None


def build(quiet: bool = False, force: bool = False):
    if built_path() and not force:
        return built_path()
    if not os.path.isdir(SRC):
        raise RuntimeError(f"{SRC} not present (the GPU box uses the prebuilt oracle/_ref)")
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="d3d_ref_")
    try:
        dst = os.path.join(tmp, "src")
        shutil.copytree(SRC, dst)
        cu = os.path.join(dst, "cuda", "deform_conv_cuda.cu")
        text = open(cu).read()
        text, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(input\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(input.scalar_type()", text)
        assert n == 2, f"expected 2 patch sites, found {n}"
        open(cu, "w").write(text)
        sources = [os.path.join(dst, "vision.cpp"), os.path.join(dst, "cpu", "deform_cpu.cpp"), cu]
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        load(name="D3D", sources=sources, extra_include_paths=[dst],
             extra_cflags=["-DWITH_CUDA", "-O2"],
             extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                                "-gencode", "arch=compute_100a,code=sm_100a"],
             build_directory=OUT, verbose=not quiet, is_python_module=False)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    # keep only the shared object
    for f in os.listdir(OUT):
        if not f.endswith(".so"):
            p = os.path.join(OUT, f)
            shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)
    return built_path()


def load_d3d():
    """Import the compiled reference module (needs CUDA at call time, not at import time)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    p = built_path()
    if p is None:
        return None
    spec = importlib.util.spec_from_file_location("D3D", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(quiet="-q" in sys.argv, force="--force" in sys.argv))
