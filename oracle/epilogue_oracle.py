"""CPU (numpy, float64 internally) restatement of the render() epilogue -- SURVEY.md section 8(f) row 1.

TEST INFRASTRUCTURE ONLY (same rules as surfel_oracle.c): never imported by streetunveiler_b200/.

Follows the reference line by line:
  gaussian_renderer/__init__.py:148-186   alpha, view->world normal rotation, median / expected depth with
                                          nan_to_num, surf_depth mix by depth_ratio, alpha-weighting (detached)
  utils/point_utils.py:8-23               depths_to_points (pinhole unprojection, intrinsics from the FoV)
  utils/point_utils.py:26-37              depth_to_normal (central differences, cross product, F.normalize)
The backward is written by hand (what autograd does for those lines); it is pinned, together with the
forward, against tests/golden/epilogue_*.npz, which were produced by the reference's unchanged render()
running on CPU with only the rasterizer stubbed (tests/golden/make_golden_epilogue.py).
"""
from __future__ import annotations

import math

import numpy as np

FLT_MAX = float(np.finfo(np.float32).max)


def camera_terms(viewmatrix, fovx, fovy, W, H):
    """viewmatrix = world_view_transform (transposed view, scene/cameras.py:61).  Returns (W3, c2w_rot, origin, K)."""
    v = np.asarray(viewmatrix, np.float64)
    c2w = np.linalg.inv(v.T)                                  # point_utils.py:9
    fx = W / (2 * math.tan(fovx / 2.0))                        # point_utils.py:11-12
    fy = H / (2 * math.tan(fovy / 2.0))
    return v[:3, :3], c2w[:3, :3], c2w[:3, 3], (fx, fy, W / 2.0, H / 2.0)


def _nan_to_num(x):
    """torch.nan_to_num(x, 0, 0): nan -> 0, +inf -> 0, -inf -> most negative float32 (gaussian_renderer:156,161)."""
    y = np.where(np.isnan(x), 0.0, x)
    y = np.where(np.isposinf(x), 0.0, y)
    return np.where(np.isneginf(x), -FLT_MAX, y)


def _rays(c2w_rot, K, W, H):
    fx, fy, cx, cy = K
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    d_cam = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], -1)     # [H,W,3] = [x,y,1] @ inv(K).T
    return d_cam @ c2w_rot.T                                                        # point_utils.py:20


def forward(allmap, viewmatrix, fovx, fovy, depth_ratio):
    a = np.asarray(allmap, np.float64)
    _, H, W = a.shape
    W3, c2w_rot, origin, K = camera_terms(viewmatrix, fovx, fovy, W, H)
    alpha = a[1]
    rend_normal = np.einsum("ji,ihw->jhw", W3, a[2:5])          # (n_row @ W3.T): out = W3 n   (:153)
    med = _nan_to_num(a[5])                                      # :156
    with np.errstate(all="ignore"):
        expected = _nan_to_num(a[0] / alpha)                     # :160-161
    surf_depth = expected * (1 - depth_ratio) + depth_ratio * med   # :168
    rays = _rays(c2w_rot, K, W, H)
    points = surf_depth[..., None] * rays + origin               # point_utils.py:22
    normal = np.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]                    # point_utils.py:33 (difference along rows)
    dy = points[1:-1, 2:] - points[1:-1, :-2]                    # :34
    n = np.cross(dx, dy)
    ln = np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-12)   # F.normalize eps
    normal[1:-1, 1:-1] = n / ln
    surf_normal = normal.transpose(2, 0, 1) * alpha              # :176 (alpha detached)
    out = {"rend_alpha": a[1:2], "rend_normal": rend_normal, "rend_dist": a[6:7], "surf_depth": surf_depth[None],
           "surf_normal": surf_normal, "surf_point": points.transpose(2, 0, 1)}
    return {k: v.astype(np.float32) for k, v in out.items()}


def backward(allmap, viewmatrix, fovx, fovy, depth_ratio, upstream):
    """upstream: dict of dL/d(output) for the six outputs (missing -> zero).  Returns dL/dallmap [7,H,W]."""
    a = np.asarray(allmap, np.float64)
    _, H, W = a.shape
    z = lambda c: np.zeros((c, H, W))  # noqa: E731
    g = {k: np.asarray(upstream.get(k, z(c)), np.float64)
         for k, c in [("rend_alpha", 1), ("rend_normal", 3), ("rend_dist", 1), ("surf_depth", 1), ("surf_normal", 3),
                      ("surf_point", 3)]}
    W3, c2w_rot, origin, K = camera_terms(viewmatrix, fovx, fovy, W, H)
    alpha = a[1]
    out = np.zeros_like(a)
    out[1] += g["rend_alpha"][0]
    out[6] += g["rend_dist"][0]
    out[2:5] += np.einsum("ji,jhw->ihw", W3, g["rend_normal"])   # transpose of the rotation
    # recompute forward intermediates
    med_raw = a[5]
    with np.errstate(all="ignore"):
        exp_raw = a[0] / alpha
    surf_depth = _nan_to_num(exp_raw) * (1 - depth_ratio) + depth_ratio * _nan_to_num(med_raw)
    rays = _rays(c2w_rot, K, W, H)
    points = surf_depth[..., None] * rays + origin
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    n = np.cross(dx, dy)
    ln_raw = np.linalg.norm(n, axis=-1, keepdims=True)
    ln = np.maximum(ln_raw, 1e-12)
    nh = n / ln
    # surf_normal = normalize(n) * alpha.detach()
    g_nh = (g["surf_normal"].transpose(1, 2, 0) * alpha[..., None])[1:-1, 1:-1]
    g_n = np.where(ln_raw > 1e-12, (g_nh - nh * np.sum(nh * g_nh, -1, keepdims=True)) / ln, g_nh / 1e-12)
    g_dx = np.cross(dy, g_n)                                      # n = dx x dy
    g_dy = np.cross(g_n, dx)
    g_pts = g["surf_point"].transpose(1, 2, 0).copy()
    g_pts[2:, 1:-1] += g_dx
    g_pts[:-2, 1:-1] -= g_dx
    g_pts[1:-1, 2:] += g_dy
    g_pts[1:-1, :-2] -= g_dy
    g_sd = g["surf_depth"][0] + np.sum(g_pts * rays, -1)
    # nan_to_num passes gradients only where its input is finite; the division's own backward then
    # divides that (zeroed) gradient by alpha again, so pixels with alpha == 0 receive 0/0 = NaN in
    # channels 0 and 1 exactly as they do from autograd in the reference (they have no contributors, so
    # the rasterizer's backward never reads them).
    fin_e, fin_m = np.isfinite(exp_raw), np.isfinite(med_raw)
    g_e = np.where(fin_e, g_sd * (1 - depth_ratio), 0.0)
    with np.errstate(all="ignore"):
        out[0] += g_e / alpha
        out[1] += -g_e * a[0] / (alpha * alpha)
    out[5] += np.where(fin_m, g_sd * depth_ratio, 0.0)
    return out.astype(np.float32)
