import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    """The C-ABI library must exist (built by __graft_entry__.build()); build it if a toolchain is here."""
    from streetunveiler_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


@pytest.fixture(scope="session")
def emul():
    """tests/emul/*.cu: the __host__ __device__ bodies of the CUDA kernels compiled for the CPU (test infrastructure,
    never loaded by the product).  Built on demand with nvcc's host compiler."""
    import ctypes
    import glob
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    d = os.path.join(ROOT, "tests", "emul")
    so = os.path.join(d, "libkernel_emul.so")
    srcs = sorted(glob.glob(os.path.join(d, "*.cu")))
    deps = srcs + glob.glob(os.path.join(ROOT, "streetunveiler_b200", "csrc", "*_tile.cuh")) + \
        [os.path.join(ROOT, "include", "surfel_rasterizer.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in deps):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-o", so] + srcs)
    return ctypes.CDLL(so)
