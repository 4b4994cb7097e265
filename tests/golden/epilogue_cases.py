"""Seeded inputs of the render()-epilogue golden cases (see make_golden_epilogue.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from streetunveiler_b200 import synthetic as syn  # noqa: E402

EPILOGUE_CASES = ["expected_depth", "median_depth", "mixed_with_holes"]
KEYS = ["rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal", "surf_point"]
CH = {"rend_alpha": 1, "rend_normal": 3, "rend_dist": 1, "surf_depth": 1, "surf_normal": 3, "surf_point": 3}


def synthetic_allmap(H, W, seed, holes):
    """A plausible allmap: smooth depth field, alpha in (0,1], unit-ish normals, distortion >= 0;
    `holes` punches alpha == 0 regions (-> 0/0 in the expected depth, exercised through nan_to_num)."""
    g = torch.Generator().manual_seed(seed)
    f64 = dict(generator=g, dtype=torch.float64)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    depth = 4.0 + 1.5 * torch.sin(xx / 9.0) * torch.cos(yy / 7.0) + 0.02 * (xx + yy) + 0.05 * torch.randn(H, W, **f64)
    alpha = (0.35 + 0.6 * torch.rand(H, W, **f64)).clamp(0.05, 1.0)
    if holes:
        alpha[torch.rand(H, W, **f64) < 0.08] = 0.0
        alpha[H // 3: H // 3 + 5, W // 4: W // 4 + 9] = 0.0
    n = torch.randn(3, H, W, **f64)
    n = n / n.norm(dim=0, keepdim=True)
    allmap = torch.zeros(7, H, W, dtype=torch.float64)
    allmap[0] = depth * alpha
    allmap[1] = alpha
    allmap[2:5] = n * alpha
    allmap[5] = torch.where(alpha > 0, depth + 0.03 * torch.randn(H, W, **f64), torch.zeros_like(depth))
    allmap[6] = 0.01 * torch.rand(H, W, **f64) * alpha
    return allmap.float().contiguous()


def build_epilogue_case(name):
    if name == "expected_depth":
        cam, ratio, holes, seed = syn.cam_tilted(96, 64, 90.0), 0.0, False, 1
    elif name == "median_depth":
        cam, ratio, holes, seed = syn.cam_tilted(80, 72, 70.0, yaw=-0.4, pitch=0.25, t=(1.0, 0.5, -0.3)), 1.0, False, 2
    elif name == "mixed_with_holes":
        cam, ratio, holes, seed = syn.cam_tilted(101, 67, 95.0), 0.3, True, 3
    else:
        raise KeyError(name)
    H, W = cam.height, cam.width
    g = torch.Generator().manual_seed(100 + seed)
    up = {k: (torch.randn(CH[k], H, W, generator=g, dtype=torch.float64) / (H * W)).float() for k in KEYS}
    return dict(cam=cam, depth_ratio=ratio, allmap=synthetic_allmap(H, W, seed, holes), upstream=up)
