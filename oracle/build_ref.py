"""Build the UNMODIFIED reference rasterizer extension for sm_100a into oracle/_ref/.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under streetunveiler_b200/ may import this.

The reference sources are compiled *where they lie* under
/root/reference/submodules/diff-surfel-rasterization (setup.py:22-30 lists the five
translation units); nothing is copied into the tracked tree.  Outputs go only to
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun):

    oracle/_ref/diff_surfel_rasterization/_C.so      the pybind module (ext.cpp:15-19)
    oracle/_ref/diff_surfel_rasterization/__init__.py the reference's own Python surface (installed copy)
    oracle/_ref/caller/{gaussian_renderer,utils,scene}/... the reference's CALLER of the operator, installed verbatim
        (gaussian_renderer/__init__.py, scene/cameras.py and the utils/ modules they import), so that the GPU box can
        execute the unchanged ``gaussian_renderer.render`` on top of either extension
        (tests/test_reference_caller_gpu.py)

Two deviations from the reference's setup.py, both build-only:
  * ``-include cstdint``: gcc 13 no longer leaks <cstdint> into rasterizer_impl.h
  * arch: ``-gencode arch=compute_100a,code=sm_100a`` instead of torch's default list
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_RAST = "/root/reference/submodules/diff-surfel-rasterization"
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "diff_surfel_rasterization")
REF_ROOT = "/root/reference"
CALLER = os.path.join(OUT, "caller")
CALLER_FILES = ["gaussian_renderer/__init__.py", "scene/cameras.py", "utils/point_utils.py", "utils/sh_utils.py",
                "utils/semantic_utils.py", "utils/graphics_utils.py", "utils/general_utils.py", "utils/loss_utils.py"]


def built() -> bool:
    return os.path.exists(os.path.join(PKG, "_C.so")) and os.path.exists(os.path.join(PKG, "__init__.py"))


def caller_installed() -> bool:
    return all(os.path.exists(os.path.join(CALLER, f)) for f in CALLER_FILES)


def install_caller(force: bool = False) -> bool:
    """Install the reference's caller modules (verbatim copies, like ``pip --target``) into the git-ignored
    oracle/_ref/caller/.  scene/__init__.py is NOT installed (it imports the dataset readers and native deps the
    operator's path never touches); Python treats oracle/_ref/caller/scene as a namespace package instead."""
    if caller_installed() and not force:
        return True
    if not os.path.isdir(REF_ROOT):
        return False
    for f in CALLER_FILES:
        dst = os.path.join(CALLER, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(os.path.join(REF_ROOT, f), dst)
    return caller_installed()


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds a usable build."""
    install_caller(force)
    if built() and not force:
        return True
    if not os.path.isdir(REF_RAST):
        return False  # GPU box: only prebuilt files travel
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    build_dir = os.path.join(OUT, "build")
    os.makedirs(build_dir, exist_ok=True)
    os.makedirs(PKG, exist_ok=True)
    srcs = [
        "cuda_rasterizer/rasterizer_impl.cu",
        "cuda_rasterizer/forward.cu",
        "cuda_rasterizer/backward.cu",
        "rasterize_points.cu",
        "ext.cpp",
    ]
    load(
        name="_C",
        sources=[os.path.join(REF_RAST, s) for s in srcs],
        extra_include_paths=[os.path.join(REF_RAST, "third_party/glm")],
        extra_cuda_cflags=[
            "-include", "cstdint",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-lineinfo",
        ],
        build_directory=build_dir,
        verbose=verbose,
        is_python_module=False,
    )
    shutil.copy2(os.path.join(build_dir, "_C.so"), os.path.join(PKG, "_C.so"))
    # "install" the reference's Python surface next to its extension (same as pip --target would)
    shutil.copy2(os.path.join(REF_RAST, "diff_surfel_rasterization", "__init__.py"),
                 os.path.join(PKG, "__init__.py"))
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference extension:", "OK" if ok else "UNAVAILABLE", PKG)
    sys.exit(0 if ok else 1)
