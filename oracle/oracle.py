"""ctypes front-end of the CPU oracle (oracle/surfel_oracle.c).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from streetunveiler_b200/.

The call signatures follow the reference's pybind entry points
(RAST/rasterize_points.cu:39-134 forward, :136-233 backward, :235-254 markVisible) with numpy
arrays instead of CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsurfel_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "surfel_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsurfel_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.oracle_forward.restype = vp
        L.oracle_forward.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp, fp, fp, fp, fp, C.c_float,
                                     fp, fp, fp, fp, fp, C.c_float, C.c_float, fp, fp, ip, C.POINTER(C.c_int64)]
        L.oracle_backward.restype = None
        L.oracle_backward.argtypes = [vp, fp, fp, fp, fp, fp, C.c_float, fp, fp, fp, fp, fp, C.c_float, C.c_float,
                                      fp, fp] + [fp] * 9
        L.oracle_free.argtypes = [vp]
        L.oracle_free.restype = None
        for name, rt in [("oracle_point_list", C.POINTER(C.c_uint32)), ("oracle_ranges", C.POINTER(C.c_uint32)),
                         ("oracle_transmat", fp), ("oracle_means2d", fp), ("oracle_depths", fp), ("oracle_rgb", fp),
                         ("oracle_normal_opacity", fp), ("oracle_tiles_touched", C.POINTER(C.c_uint32)),
                         ("oracle_final_T", fp), ("oracle_n_contrib", C.POINTER(C.c_uint32))]:
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp]
        L.oracle_num_rendered.restype = C.c_int64
        L.oracle_num_rendered.argtypes = [vp]
        L.oracle_mark_visible.argtypes = [C.c_int, fp, fp, fp, C.POINTER(C.c_uint8)]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if a.size else None


def _p(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class OracleForward:
    """Result of one oracle forward; keeps the saved state the backward needs."""

    def __init__(self):
        self.handle = None

    def __del__(self):
        if self.handle:
            lib().oracle_free(self.handle)
            self.handle = None

    def _arr(self, fn, n, dtype):
        ptr = getattr(lib(), fn)(self.handle)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    # debug views of the saved state (GeometryState / BinningState / ImageState)
    def point_list(self):
        return self._arr("oracle_point_list", max(self.num_rendered, 1), np.uint32)[: self.num_rendered]

    def ranges(self):
        return self._arr("oracle_ranges", 2 * self.tiles, np.uint32).reshape(self.tiles, 2)

    def transmat(self):
        return self._arr("oracle_transmat", 9 * self.P, np.float32).reshape(self.P, 9)

    def means2d(self):
        return self._arr("oracle_means2d", 2 * self.P, np.float32).reshape(self.P, 2)

    def depths(self):
        return self._arr("oracle_depths", self.P, np.float32)

    def rgb(self):
        return self._arr("oracle_rgb", 3 * self.P, np.float32).reshape(self.P, 3)

    def normal_opacity(self):
        return self._arr("oracle_normal_opacity", 4 * self.P, np.float32).reshape(self.P, 4)

    def tiles_touched(self):
        return self._arr("oracle_tiles_touched", self.P, np.uint32)

    def final_T(self):
        return self._arr("oracle_final_T", 3 * self.W * self.H, np.float32).reshape(3, self.H, self.W)

    def n_contrib(self):
        return self._arr("oracle_n_contrib", 2 * self.W * self.H, np.uint32).reshape(2, self.H, self.W)


def rasterize_forward(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, transMat_precomp,
                      viewmatrix, projmatrix, tanfovx, tanfovy, image_height, image_width, shs, sh_degree, campos):
    """Argument order of _C.rasterize_gaussians (RAST/rasterize_points.cu:39-59) minus prefiltered/debug."""
    means3D = _f32(means3D)
    P = 0 if means3D is None else means3D.shape[0]
    H, W = int(image_height), int(image_width)
    out = OracleForward()
    out.P, out.W, out.H = P, W, H
    out.tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out.color = np.zeros((3, H, W), np.float32)
    out.allmap = np.zeros((7, H, W), np.float32)
    out.radii = np.zeros((P,), np.int32)
    out.num_rendered = 0
    if P == 0:
        return out
    shs_a, col_a = _f32(shs), _f32(colors_precomp)
    M = shs_a.shape[1] if shs_a is not None else 0
    keep = [_f32(bg), means3D, shs_a, col_a, _f32(opacities), _f32(scales), _f32(rotations), _f32(transMat_precomp),
            _f32(viewmatrix), _f32(projmatrix), _f32(campos)]
    out._inputs = keep
    nr = C.c_int64(0)
    out.handle = lib().oracle_forward(
        P, int(sh_degree), M, _p(keep[0]), W, H, _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(keep[5]),
        float(scale_modifier), _p(keep[6]), _p(keep[7]), _p(keep[8]), _p(keep[9]), _p(keep[10]), float(tanfovx),
        float(tanfovy), _p(out.color), _p(out.allmap), out.radii.ctypes.data_as(C.POINTER(C.c_int)), C.byref(nr))
    out.num_rendered = int(nr.value)
    out.M = M
    out.meta = dict(scale_modifier=float(scale_modifier), tanfovx=float(tanfovx), tanfovy=float(tanfovy))
    return out


def rasterize_backward(fwd: OracleForward, dL_dcolor, dL_dallmap):
    """Returns the dict of gradients the reference's backward returns (RAST/rasterize_points.cu:232),
    plus the internal dL_dnormal."""
    P, M = fwd.P, fwd.M
    g = {
        "means2D": np.zeros((P, 3), np.float32), "colors": np.zeros((P, 3), np.float32),
        "opacities": np.zeros((P, 1), np.float32), "means3D": np.zeros((P, 3), np.float32),
        "transMat": np.zeros((P, 9), np.float32), "sh": np.zeros((P, M, 3), np.float32),
        "scales": np.zeros((P, 2), np.float32), "rotations": np.zeros((P, 4), np.float32),
        "normal": np.zeros((P, 3), np.float32),
    }
    if P == 0:
        return g
    bg, means3D, shs, col, _opa, scales, rots, tmp, view, proj, campos = fwd._inputs
    dc, da = _f32(dL_dcolor), _f32(dL_dallmap)
    lib().oracle_backward(
        fwd.handle, _p(bg), _p(means3D), _p(shs), _p(col), _p(scales), fwd.meta["scale_modifier"], _p(rots), _p(tmp),
        _p(view), _p(proj), _p(campos), fwd.meta["tanfovx"], fwd.meta["tanfovy"], _p(dc), _p(da),
        _p(g["means2D"]), _p(g["colors"]), _p(g["opacities"]), _p(g["means3D"]), _p(g["transMat"]),
        _p(g["sh"]) if M > 0 else None, _p(g["scales"]), _p(g["rotations"]), _p(g["normal"]))
    return g


def mark_visible(means3D, viewmatrix, projmatrix) -> np.ndarray:
    m = _f32(means3D)
    P = 0 if m is None else m.shape[0]
    out = np.zeros((P,), np.uint8)
    if P:
        v, pr = _f32(viewmatrix), _f32(projmatrix)
        lib().oracle_mark_visible(P, _p(m), _p(v), _p(pr), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.astype(bool)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))
