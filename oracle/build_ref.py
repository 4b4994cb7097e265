"""Build the UNMODIFIED reference rasterizer extension for sm_100a into oracle/_ref/.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under streetunveiler_b200/ may import this.

The reference sources are compiled *where they lie* under
/root/reference/submodules/diff-surfel-rasterization (setup.py:22-30 lists the five
translation units); nothing is copied into the tracked tree.  Outputs go only to
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun):

    oracle/_ref/diff_surfel_rasterization/_C.so      the pybind module (ext.cpp:15-19)
    oracle/_ref/diff_surfel_rasterization/__init__.py the reference's own Python surface (installed copy)

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


def built() -> bool:
    return os.path.exists(os.path.join(PKG, "_C.so")) and os.path.exists(os.path.join(PKG, "__init__.py"))


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds a usable build."""
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
