"""ctypes loader of libsurfel_b200.so (C ABI: include/surfel_rasterizer.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, a RuntimeError
is raised.  Nothing here imports or executes oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurfel_b200.so")
ABI_VERSION = 1

_lib = None

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes): every symbol declared in include/surfel_rasterizer.h
SIGNATURES = {
    "surfel_abi_version": (_i, []),
    "surfel_last_error": (C.c_char_p, []),
    "surfel_geometry_bytes": (C.c_size_t, [_i]),
    "surfel_image_bytes": (C.c_size_t, [_i, _i]),
    "surfel_binning_bytes": (C.c_size_t, [_i64]),
    "surfel_grad_scratch_bytes": (C.c_size_t, [_i]),
    "surfel_forward_prepare": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f,
                                    _i, _vp, _vp, C.POINTER(_i64), _vp, _i]),
    "surfel_forward_render": (_i, [_i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_backward": (_i, [_i, _i, _i, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f,
                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "surfel_debug_copy_binning": (_i, [_i, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "surfel_shard_preprocess": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f, _i,
                                     _vp, _vp, _vp, _vp, _vp]),
    "surfel_shard_compact_bytes": (C.c_size_t, [_i]),
    "surfel_shard_compact": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_shard_tile_hist": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "surfel_shard_partition_bytes": (C.c_size_t, [_i, _i]),
    "surfel_shard_partition": (_i, [_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "surfel_shard_route_bytes": (C.c_size_t, [_i, _i]),
    "surfel_shard_route_count": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "surfel_shard_route_scatter": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_shard_route_scatter_peers": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_window_push_grad_rows": (_i, [_i64, _vp, _i, _vp, _vp, _vp, _vp]),
    "surfel_window_unpack": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "surfel_shard_grad_accumulate": (_i, [_i, _i64, _vp, _vp, _vp, _vp]),
    "surfel_window_bytes": (C.c_size_t, [_i]),
    "surfel_window_prepare": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp, _i]),
    "surfel_window_render": (_i, [_i, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_window_render_peers": (_i, [_i, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i]),
    "surfel_window_backward": (_i, [_i, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_shard_backward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_epilogue_forward": (_i, [_i, _i, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_epilogue_backward": (_i, [_i, _i, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_pass_set_colors": (_i, [_i, _vp, _vp, _vp]),
    "surfel_pass_render": (_i, [_i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_pass_backward_blend": (_i, [_i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i]),
    "surfel_pass_take_color_grad": (_i, [_i, _vp, _vp, _vp]),
    "surfel_pass_backward_geometry": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_forward_bin": (_i, [_i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_classes_set_labels": (_i, [_i, _vp, _vp, _vp]),
    "surfel_classes_render": (_i, [_i, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "surfel_classes_backward_blend": (_i, [_i, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i]),
    "surfel_loss_scratch_bytes": (C.c_size_t, [_i, _i]),
    "surfel_loss_photometric_forward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_loss_photometric_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_loss_regulariser_forward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_loss_regulariser_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_loss_training_forward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "surfel_loss_training_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp]),
    "surfel_activate_forward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_activate_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_adam_step": (_i, [_i, _vp, C.c_double, C.c_double, C.c_double, _vp]),
    "surfel_densification_stats": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_debug_sort_pairs": (_i, [_i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "surfel_debug_copy_geometry": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "surfel_debug_aux_flag": (_i, [_i, _vp, C.POINTER(_i), _vp]),
    "surfel_set_option": (_i, [C.c_char_p, _i]),
    "surfel_stage_count": (_i, []),
    "surfel_stage_name": (C.c_char_p, [_i]),
    "surfel_stage_time": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_i)]),
}


class AdamGroup(C.Structure):
    """struct surfel_adam_group (include/surfel_rasterizer.h)."""
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("n", _i64), ("lr", C.c_double),
                ("step", _i)]


def build(verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess

    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libsurfel_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C streetunveiler_b200/csrc`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if L.surfel_abi_version() != ABI_VERSION:
            raise RuntimeError("libsurfel_b200.so ABI version mismatch; rebuild it")
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().surfel_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")


def size(n: int, what: str) -> int:
    if n == 0:
        msg = lib().surfel_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")
    return int(n)


def set_option(name: str, value: int) -> None:
    check(lib().surfel_set_option(name.encode(), int(value)), f"surfel_set_option({name})")


def stage_times() -> dict:
    """{stage name: (total ms, timed launches)} accumulated since time_stages was last set."""
    L = lib()
    out = {}
    for i in range(L.surfel_stage_count()):
        ms, n = C.c_double(0), C.c_int(0)
        check(L.surfel_stage_time(i, C.byref(ms), C.byref(n)), "surfel_stage_time")
        out[L.surfel_stage_name(i).decode()] = (ms.value, n.value)
    return out
