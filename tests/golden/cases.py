"""Definition of the golden-vector cases (inputs are regenerated from seeds; see make_golden.py)."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from streetunveiler_b200 import synthetic as syn  # noqa: E402


def _transmat_like(scene, cam):
    """A deterministic [P,9] splat->pixel matrix for the precomputed-transMat path: the textbook
    (double precision) formula of SURVEY.md 8(a9), independent of any implementation under test."""
    P = scene["means3D"].shape[0]
    q = scene["rotations"].double()
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
    L0 = R[:, :, 0] * scene["scales"][:, 0:1].double()
    L1 = R[:, :, 1] * scene["scales"][:, 1:2].double()
    pm = cam.projmatrix.double()  # row-vector convention: hom = [v, w] @ pm
    W, H = cam.width, cam.height

    def hom(v, w):
        return torch.cat([v, torch.full((P, 1), float(w), dtype=torch.float64)], 1) @ pm

    rows = [hom(L0, 0), hom(L1, 0), hom(scene["means3D"].double(), 1)]
    T = torch.zeros(P, 3, 3, dtype=torch.float64)  # T[:, j, r]: j = u/v/w row, r = tangent-u/tangent-v/centre
    for r, h in enumerate(rows):
        T[:, 0, r] = h[:, 0] * W / 2 + h[:, 3] * (W - 1) / 2
        T[:, 1, r] = h[:, 1] * H / 2 + h[:, 3] * (H - 1) / 2
        T[:, 2, r] = h[:, 3]
    return T.reshape(P, 9).float().contiguous()


def build_case(name: str):
    """-> dict(scene, cam, bg, grads, kwargs) for the named case."""
    if name == "box_sh0":  # BASELINE config 1 in miniature: SH degree 0, black background
        cam = syn.cam_s(128, 128, 110.85)
        scene = syn.box_scene(2500, 3, 0)
        return dict(scene=scene, cam=cam, bg=torch.zeros(3), grads=syn.upstream_grads(128, 128, "color_alpha"), kw={})
    if name == "box_sh3_tilt":  # dense view/proj matrices, all 7 allmap gradient channels, coloured bg
        cam = syn.cam_tilted(112, 80, 100.0)
        scene = syn.box_scene(1200, 7, 3)
        return dict(scene=scene, cam=cam, bg=torch.tensor([0.3, 0.1, 0.7]),
                    grads=syn.upstream_grads(112, 80, "all", seed=101), kw={})
    if name == "odd_size_sh2":  # image size not a multiple of the 16x16 tile, SH degree 2 of 3 active
        cam = syn.cam_tilted(101, 67, 95.0, yaw=-0.15, pitch=0.05)
        scene = syn.box_scene(1000, 11, 3)
        scene["sh_degree"] = 2
        return dict(scene=scene, cam=cam, bg=torch.tensor([1.0, 1.0, 1.0]),
                    grads=syn.upstream_grads(101, 67, "all", seed=102), kw={})
    if name == "precomp_color":  # colors_precomp instead of SH
        cam = syn.cam_tilted(96, 64, 90.0)
        scene = syn.box_scene(900, 13, 0)
        return dict(scene=scene, cam=cam, bg=torch.tensor([0.0, 0.5, 0.0]),
                    grads=syn.upstream_grads(96, 64, "all", seed=103), kw=dict(use_precomp_color=True))
    if name == "precomp_transmat":  # cov3D_precomp ([P,9] transMat) instead of scale+rotation
        cam = syn.cam_tilted(96, 64, 90.0)
        scene = syn.box_scene(900, 17, 1)
        return dict(scene=scene, cam=cam, bg=torch.zeros(3), grads=syn.upstream_grads(96, 64, "all", seed=104),
                    kw=dict(transmat_precomp=_transmat_like(scene, cam)))
    if name == "scale_modifier":  # scale_modifier != 1 (backward ignores it: quirk 2)
        cam = syn.cam_s(96, 96, 83.0)
        scene = syn.box_scene(900, 19, 3)
        return dict(scene=scene, cam=cam, bg=torch.zeros(3), grads=syn.upstream_grads(96, 96, "all", seed=105),
                    kw=dict(scale_modifier=1.3))
    if name == "street_small":  # the benchmark's scene family and camera aspect at small scale
        cam = syn.make_camera(240, 160, 256.875, 256.875)
        scene = syn.street_scene(6000, 5, 3)
        scene["scales"] = (scene["scales"] * 3).contiguous()
        return dict(scene=scene, cam=cam, bg=torch.zeros(3), grads=syn.upstream_grads(240, 160, "color_alpha", seed=106),
                    kw={})
    if name == "config1_box10k":  # BASELINE.json configs[0] at full size: BOX(10k, seed 3), CAM-S 256x256, SH degree 0
        cam = syn.cam_s()
        scene = syn.box_scene(10_000, 3, 0)
        return dict(scene=scene, cam=cam, bg=torch.zeros(3), grads=syn.upstream_grads(256, 256, "color_alpha"), kw={})
    raise KeyError(name)


CASES = ["box_sh0", "box_sh3_tilt", "odd_size_sh2", "precomp_color", "precomp_transmat", "scale_modifier",
         "street_small", "config1_box10k"]
