"""Deterministic synthetic scenes and cameras for parity tests and bench.py.

Everything is generated on the CPU with a seeded ``torch.Generator`` so that the oracle, the
reference extension and the B200 kernels all see bit-identical inputs.  Camera matrices use the
reference's memory layout: ``viewmatrix`` / ``projmatrix`` are the *transposes* of the usual
column-vector matrices (scene/cameras.py:61-71, utils/graphics_utils.py:38-79).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch


@dataclass
class Camera:
    width: int
    height: int
    fx: float
    fy: float
    tanfovx: float
    tanfovy: float
    fovx: float
    fovy: float
    viewmatrix: torch.Tensor   # [4,4] world->view, transposed (row-vector convention)
    projmatrix: torch.Tensor   # [4,4] full view*proj, transposed
    campos: torch.Tensor       # [3]
    znear: float = 0.01
    zfar: float = 100.0


def _projection(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """Same frustum matrix as utils/graphics_utils.py:51-79 (K=None branch), z in [0, 1]."""
    t = math.tan(fovy / 2) * znear
    r = math.tan(fovx / 2) * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(width: int, height: int, fx: float, fy: float, R: torch.Tensor | None = None,
                T: torch.Tensor | None = None, znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """R is the camera-to-world rotation and T the world-to-view translation, as in scene/cameras.py."""
    fovx = 2 * math.atan(width / (2 * fx))
    fovy = 2 * math.atan(height / (2 * fy))
    R = torch.eye(3, dtype=torch.float64) if R is None else R.to(torch.float64)
    T = torch.zeros(3, dtype=torch.float64) if T is None else T.to(torch.float64)
    Rt = torch.zeros(4, 4, dtype=torch.float64)
    Rt[:3, :3] = R.t()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    w2v = Rt.to(torch.float32)
    view_t = w2v.t().contiguous()
    proj_t = _projection(znear, zfar, fovx, fovy).t().contiguous()
    full = (view_t.unsqueeze(0).bmm(proj_t.unsqueeze(0))).squeeze(0).contiguous()
    campos = view_t.inverse()[3, :3].contiguous()
    return Camera(width, height, fx, fy, math.tan(fovx * 0.5), math.tan(fovy * 0.5), fovx, fovy,
                  view_t, full, campos, znear, zfar)


def cam_a() -> Camera:
    """CAM-A: 1920x1280, fx=fy=2055 (Waymo FRONT-like), camera at origin looking +z."""
    return make_camera(1920, 1280, 2055.0, 2055.0)


def cam_s(width: int = 256, height: int = 256, f: float = 221.7) -> Camera:
    """CAM-S: 256x256, ~60 deg FoV."""
    return make_camera(width, height, f, f)


def cam_tilted(width: int, height: int, f: float, yaw: float = 0.2, pitch: float = -0.1,
               t=(0.3, -0.2, 0.5)) -> Camera:
    """A non-trivial pose so that view/proj matrices are dense (exercises every matrix term)."""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    Rx = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=torch.float64)
    return make_camera(width, height, f, f, R=Ry @ Rx, T=torch.tensor(t, dtype=torch.float64))


def _common(P: int, g: torch.Generator, scale_med: float, scale_sig: float, smin: float, smax: float,
            sh_coeffs: int) -> Dict[str, torch.Tensor]:
    # Everything is drawn and transformed in float64 and rounded to float32 once at the end, so that
    # 1-ulp differences of vectorised float32 exp/sigmoid between hosts or thread counts cannot leak
    # into the scene (both benchmark arms must see bit-identical inputs; bench.py prints their CRC).
    f64 = dict(generator=g, dtype=torch.float64)
    scales = torch.exp(torch.randn(P, 2, **f64) * scale_sig + math.log(scale_med)).clamp(smin, smax)
    rot = torch.randn(P, 4, **f64)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opac = torch.sigmoid(torch.randn(P, 1, **f64) * 1.5)
    shs = torch.randn(P, sh_coeffs, 3, **f64)
    shs[:, 0, :] *= 0.6
    if sh_coeffs > 1:
        shs[:, 1:, :] *= 0.08
    return {"scales": scales.float().contiguous(), "rotations": rot.float().contiguous(),
            "opacities": opac.float().contiguous(), "shs": shs.float().contiguous()}


def street_scene(P: int, seed: int, sh_degree: int = 3) -> Dict[str, torch.Tensor]:
    """STREET(P, seed): 40% ground, 40% facades, 20% clutter, then one fixed randperm."""
    g = torch.Generator("cpu").manual_seed(seed)
    n_g = int(0.4 * P)
    n_f = int(0.4 * P)
    n_c = P - n_g - n_f
    f64 = dict(generator=g, dtype=torch.float64)
    u = lambda n, a, b: torch.rand(n, **f64) * (b - a) + a  # noqa: E731
    ground = torch.stack([u(n_g, -15, 15), 1.6 + 0.02 * torch.randn(n_g, **f64), u(n_g, 1, 90)], 1)
    side = (torch.rand(n_f, **f64) < 0.5).double() * 2 - 1
    fac = torch.stack([side * (10 + 0.05 * torch.randn(n_f, **f64)), u(n_f, -10, 1.6), u(n_f, 1, 90)], 1)
    clu = torch.stack([u(n_c, -9, 9), u(n_c, -3, 1.6), u(n_c, 3, 60)], 1)
    xyz = torch.cat([ground, fac, clu], 0)
    out = _common(P, g, 0.03, 0.5, 0.004, 0.4, (sh_degree + 1) ** 2)
    perm = torch.randperm(P, generator=g)
    out["means3D"] = xyz[perm].float().contiguous()
    out["sh_degree"] = sh_degree
    return out


def box_scene(P: int = 10_000, seed: int = 3, sh_degree: int = 0) -> Dict[str, torch.Tensor]:
    """BOX(P, seed): x,y~U(-3,3), z~U(2,10)."""
    g = torch.Generator("cpu").manual_seed(seed)
    f64 = dict(generator=g, dtype=torch.float64)
    xyz = torch.stack([torch.rand(P, **f64) * 6 - 3, torch.rand(P, **f64) * 6 - 3, torch.rand(P, **f64) * 8 + 2], 1)
    out = _common(P, g, 0.08, 0.4, 0.01, 0.5, (sh_degree + 1) ** 2)
    out["means3D"] = xyz.float().contiguous()
    out["sh_degree"] = sh_degree
    return out


def scene_crc(scene: Dict[str, torch.Tensor]) -> int:
    """CRC32 over the raw bytes of every input tensor (identical inputs <=> identical CRC)."""
    import zlib
    crc = 0
    for k in sorted(scene):
        v = scene[k]
        if isinstance(v, torch.Tensor):
            crc = zlib.crc32(v.contiguous().numpy().tobytes(), crc)
    return crc


def upstream_grads(W: int, H: int, mode: str = "color_alpha", seed: int = 100):
    """dL/dcolor [3,H,W] and dL/dallmap [7,H,W] (SURVEY.md 8d 'Upstream gradients')."""
    g = torch.Generator("cpu").manual_seed(seed)
    f64 = dict(generator=g, dtype=torch.float64)
    HW = W * H
    d_color = (torch.randn(3, H, W, **f64) / (3 * HW)).float()
    d_all = torch.zeros(7, H, W)
    if mode == "color_alpha":
        d_all[1] = (torch.randn(H, W, **f64) / HW).float()
    elif mode == "all":
        d_all = (torch.randn(7, H, W, **f64) / (7 * HW)).float()
    elif mode == "color":
        pass
    else:
        raise ValueError(mode)
    return d_color.contiguous(), d_all.contiguous()


def depth_separable_order(means3D: torch.Tensor, cam: Camera) -> torch.Tensor:
    """Permutation that sorts Gaussians by view-space depth (front first) for the sharded path."""
    z = means3D @ cam.viewmatrix[:3, 2] + cam.viewmatrix[3, 2]
    return torch.argsort(z, stable=True)
