"""render_semantic in one rasterizer traversal ("next" row 2).

``render_semantic`` has the signature and return dict of the reference's ``gaussian_renderer.render_semantic``
(gaussian_renderer/__init__.py:327-460): per-class probability images rendered from one-hot "colours", three classes
per rasterizer pass, the background one-hot at the sky class.  The reference issues ceil(6/3) = 2 complete rasterizer
calls (each with its own projection, sorts, binning and backward); here all classes come out of ONE blend pass over the
per-Gaussian labels (``rasterize_class_probabilities``; up to 8 classes), or -- ``single_pass=False``, any number of
classes -- from colour passes that share one geometry / binning state (``rasterize_color_passes``).  The reference's own
file keeps running unchanged on the drop-in rasterizer -- this module is the optional faster caller, like
``surface_epilogue.render``.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from .diff_surfel_rasterization import GaussianRasterizationSettings
from .diff_surfel_rasterization.class_pass import MAX_CLASSES, rasterize_class_probabilities
from .diff_surfel_rasterization.color_passes import rasterize_color_passes

# utils/semantic_utils.py:100-102 (the six classes StreetUnveiler trains on) and :14-21 (their display colours)
CONCERNED_CLASSES = ("road", "sidewalk", "building", "vegetation", "sky", "vehicle")
CLASS_COLORS = ((255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (255, 0, 255), (0, 255, 255))


def one_hot_colors(semantics_tag: torch.Tensor, first_class: int, num_classes: int) -> torch.Tensor:
    """[P,3] with column j = 1 where tag == first_class + j (reference :420-430; columns past the last class stay 0)."""
    tag = semantics_tag.reshape(-1, 1)
    cls = torch.arange(first_class, first_class + 3, device=tag.device, dtype=tag.dtype).reshape(1, 3)
    valid = (cls < num_classes)
    return ((tag == cls) & valid).to(torch.float32)


def render_semantic(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0,
                    semantic_filter_bit: Optional[int] = None, reverse_semantic: Optional[bool] = None,
                    classes: Sequence[str] = CONCERNED_CLASSES, single_pass: bool = True):
    n_cls = len(classes)
    sky = list(classes).index("sky")
    dev = pc.get_xyz.device
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    bg_prob = [0.0] * n_cls
    bg_prob[sky] = 1.0                                                   # reference :348-349
    bgs = []
    for i in range(0, n_cls, 3):
        b = bg_prob[i:i + 3]
        bgs.append(torch.tensor(b + [0.0] * (3 - len(b)), dtype=torch.float32, device=dev))
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bgs[0],
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=pipe.debug)

    if semantic_filter_bit is None:
        sel = lambda t: t
    else:
        assert reverse_semantic is not None
        semantic_mask = (pc.get_semantics_32bit & np.int32(semantic_filter_bit) != 0)
        if reverse_semantic is False:
            semantic_mask = ~semantic_mask
        semantic_mask = semantic_mask.reshape(-1)
        sel = lambda t: t[semantic_mask]
    means3D, means2D, opacity = sel(pc.get_xyz), sel(screenspace_points), sel(pc.get_opacity)
    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = sel(pc.get_covariance(scaling_modifier))
    else:
        scales, rotations = sel(pc.get_scaling), sel(pc.get_rotation)
    semantics_tag = sel(pc.get_semantics)

    if single_pass and n_cls <= MAX_CLASSES:
        labels = semantics_tag.reshape(-1).to(torch.int32)
        bg_vec = torch.tensor(bg_prob, dtype=torch.float32, device=dev)
        output_semantic, _radii = rasterize_class_probabilities(settings, means3D, means2D, opacity, labels, bg_vec,
                                                                scales=scales, rotations=rotations,
                                                                cov3D_precomp=cov3D_precomp)
    else:
        colors = [one_hot_colors(semantics_tag, i, n_cls) for i in range(0, n_cls, 3)]
        images, _radii, _allmap = rasterize_color_passes(settings, means3D, means2D, opacity, colors, bgs, scales=scales,
                                                         rotations=rotations, cov3D_precomp=cov3D_precomp)
        output_semantic = torch.cat([img[:min(3, n_cls - 3 * k)] for k, img in enumerate(images)], dim=0)

    topk_values, _ = torch.topk(output_semantic, k=2, dim=0)
    uncertainty = 1.0 - (topk_values[0, ...] - topk_values[1, ...])
    palette = torch.tensor(CLASS_COLORS[:n_cls], dtype=torch.uint8, device=dev)
    semantic_rgb = palette[torch.argmax(output_semantic, dim=0)].permute(2, 0, 1) / 255.0
    return {"render_semantics": output_semantic, "semantic_rgb": semantic_rgb, "semantic_uncertainty": uncertainty}
