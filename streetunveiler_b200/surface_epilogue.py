"""Fused render() epilogue and a render() front-end built on it ("next" row 1).

``render_epilogue`` replaces lines 148-186 of the reference's gaussian_renderer/__init__.py (plus
utils/point_utils.py:8-37): same six outputs with the same names, shapes and gradients, computed by one
CUDA kernel forward and one backward (csrc/epilogue.cu) instead of ~15 PyTorch kernels each way.
``render`` has the signature and return dict of the reference's ``gaussian_renderer.render`` (:18-188);
the reference's own file keeps working unchanged on top of the drop-in rasterizer -- this module is the
optional faster caller.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def _p(t):
    return C.c_void_p(t.data_ptr())


class _SurfaceEpilogue(torch.autograd.Function):
    """All six outputs come from one kernel and all gradients go back through one kernel: returning
    rend_alpha / rend_dist as copies (instead of autograd views of allmap, as the reference does) avoids two
    zero-filled [7,H,W] slice-backward tensors and their additions per step."""

    @staticmethod
    def forward(ctx, allmap, viewmatrix, fx, fy, depth_ratio):
        if not allmap.is_cuda:
            raise RuntimeError("allmap must be a CUDA tensor")
        allmap = allmap.contiguous().float()
        viewmatrix = viewmatrix.contiguous().float()
        _, H, W = allmap.shape
        f32 = dict(dtype=torch.float32, device=allmap.device)
        rend_normal = torch.empty((3, H, W), **f32)
        surf_depth = torch.empty((1, H, W), **f32)
        surf_normal = torch.empty((3, H, W), **f32)
        surf_point = torch.empty((3, H, W), **f32)
        rend_alpha = torch.empty((1, H, W), **f32)
        rend_dist = torch.empty((1, H, W), **f32)
        with torch.cuda.device(allmap.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_epilogue_forward(W, H, _p(allmap), _p(viewmatrix), fx, fy, depth_ratio,
                                                          _p(rend_normal), _p(surf_depth), _p(surf_normal),
                                                          _p(surf_point), _p(rend_alpha), _p(rend_dist), st),
                       "surfel_epilogue_forward")
        ctx.save_for_backward(allmap, surf_point, viewmatrix)
        ctx.consts = (fx, fy, depth_ratio)
        return rend_alpha, rend_normal, rend_dist, surf_depth, surf_normal, surf_point

    @staticmethod
    def backward(ctx, g_ra, g_rn, g_rd, g_sd, g_sn, g_sp):
        allmap, surf_point, viewmatrix = ctx.saved_tensors
        fx, fy, depth_ratio = ctx.consts
        _, H, W = allmap.shape
        g = [x.contiguous().float() for x in (g_rn, g_sd, g_sn, g_sp, g_ra, g_rd)]
        out = torch.empty_like(allmap)
        with torch.cuda.device(allmap.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_epilogue_backward(W, H, _p(allmap), _p(surf_point), _p(viewmatrix), fx, fy,
                                                           depth_ratio, _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]),
                                                           _p(g[5]), _p(out), st), "surfel_epilogue_backward")
        return out, None, None, None, None


def render_epilogue(allmap: torch.Tensor, viewpoint_camera, depth_ratio: float) -> dict:
    """allmap [7,H,W] from the rasterizer -> the dict entries the reference adds at :178-186."""
    W, H = int(viewpoint_camera.image_width), int(viewpoint_camera.image_height)
    fx = W / (2 * math.tan(viewpoint_camera.FoVx / 2.0))          # utils/point_utils.py:11-12
    fy = H / (2 * math.tan(viewpoint_camera.FoVy / 2.0))
    rend_alpha, rend_normal, rend_dist, surf_depth, surf_normal, surf_point = _SurfaceEpilogue.apply(
        allmap, viewpoint_camera.world_view_transform, float(fx), float(fy), float(depth_ratio))
    return {"rend_alpha": rend_alpha, "rend_normal": rend_normal, "rend_dist": rend_dist, "surf_depth": surf_depth,
            "surf_normal": surf_normal, "surf_point": surf_point}


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """Same contract as the reference's gaussian_renderer.render for the un-filtered case
    (semantic_filter_bit=None, SH colours or override_color, scale+rotation input)."""
    means3D = pc.get_xyz
    means2D = torch.zeros_like(means3D, requires_grad=True) + 0
    try:
        means2D.retain_grad()
    except Exception:
        pass
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=pipe.debug)
    kw = dict(shs=pc.get_features) if override_color is None else dict(colors_precomp=override_color)
    rendered_image, radii, allmap = GaussianRasterizer(settings)(
        means3D=means3D, means2D=means2D, opacities=pc.get_opacity, scales=pc.get_scaling, rotations=pc.get_rotation,
        **kw)
    rets = {"render": rendered_image, "viewspace_points": means2D, "visibility_filter": radii > 0, "radii": radii}
    rets.update(render_epilogue(allmap, viewpoint_camera, pipe.depth_ratio))
    return rets
