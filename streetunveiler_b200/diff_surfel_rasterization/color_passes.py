"""Several colour passes over ONE projection / sort / binning state ("next" row 2).

The reference's ``render_semantic`` (gaussian_renderer/__init__.py:417-446) calls ``GaussianRasterizer`` once per three
semantic classes with identical geometry and different ``colors_precomp`` / ``bg``; every call repeats the projection,
both sorts and the binning, and autograd runs one complete backward per call.  ``rasterize_color_passes`` gives the same
images, ``radii``, ``allmap`` and gradients from

    forward    K1-K5 once, then one blend (K6) per colour set
    backward   one backward blend (K7) per colour set into a shared per-Gaussian accumulator, then K8 once

through the ``surfel_pass_*`` entry points of the C ABI.  There is no fallback: CPU tensors or a failing call raise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .. import _lib
from . import _C
from ._C import _dev_f32, _ptr, _stream


class _RasterizeColorPasses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, scales, rotations, cov3Ds_precomp, raster_settings, bgs, *colors):
        s = raster_settings
        n = len(colors)
        if n == 0 or len(bgs) != n:
            raise RuntimeError("one background per colour set")
        P = int(means3D.size(0))
        colors = tuple(_dev_f32(c, "colors_precomp") for c in colors)
        for c in colors:
            if c.dim() != 2 or tuple(c.shape) != (P, 3):
                raise RuntimeError("every colour set must have shape (num_points, 3)")
        bgs = tuple(_dev_f32(b, "bg") for b in bgs)
        empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
        # pass 0 = the ordinary operator: projection, sorts, binning, blend
        num_rendered, color0, others, radii, geom_buf, bin_buf, img_buf = _C.rasterize_gaussians(
            bgs[0], means3D, colors[0], opacities, scales, rotations, s.scale_modifier, cov3Ds_precomp, s.viewmatrix,
            s.projmatrix, s.tanfovx, s.tanfovy, s.image_height, s.image_width, empty, s.sh_degree, s.campos,
            s.prefiltered, s.debug)
        outs = [color0]
        H, W = int(s.image_height), int(s.image_width)
        L = _lib.lib()
        with torch.cuda.device(means3D.device):
            st = _stream()
            scratch_others = torch.empty_like(others) if n > 1 else None
            for i in range(1, n):
                out = torch.empty_like(color0)
                _lib.check(L.surfel_pass_set_colors(P, _ptr(colors[i]), _ptr(geom_buf), st), "surfel_pass_set_colors")
                _lib.check(L.surfel_pass_render(P, W, H, num_rendered, _ptr(bgs[i]), _ptr(geom_buf), _ptr(bin_buf),
                                                _ptr(img_buf), _ptr(out), _ptr(scratch_others), st, int(bool(s.debug))),
                           "surfel_pass_render")
                outs.append(out)
        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.n = n
        ctx.save_for_backward(means3D, scales, rotations, cov3Ds_precomp, radii, geom_buf, bin_buf, img_buf, *bgs, *colors)
        ctx.mark_non_differentiable(radii)
        return (radii, others, *outs)

    @staticmethod
    def backward(ctx, g_radii, g_others, *g_colors):
        s, n = ctx.raster_settings, ctx.n
        saved = ctx.saved_tensors
        means3D, scales, rotations, cov3Ds_precomp, radii, geom_buf, bin_buf, img_buf = saved[:8]
        bgs, colors = saved[8:8 + n], saved[8 + n:8 + 2 * n]
        P = int(means3D.size(0))
        H, W = int(s.image_height), int(s.image_width)
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        need_color = ctx.needs_input_grad[8:8 + n]
        L = _lib.lib()
        with torch.cuda.device(dev):
            st = _stream()
            zeros_others = None
            d_colors: List[Optional[torch.Tensor]] = [None] * n
            d_means3D, d_means2D = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32)
            d_opacity, d_transMat = torch.empty((P, 1), **f32), torch.empty((P, 9), **f32)
            d_scales, d_rot = torch.empty((P, 2), **f32), torch.empty((P, 4), **f32)
            d_color_last = torch.empty((P, 3), **f32)
            if P:
                scratch = torch.empty((_lib.size(L.surfel_grad_scratch_bytes(P), "surfel_grad_scratch_bytes"),),
                                      dtype=torch.uint8, device=dev)
                for i in range(n):
                    # autograd materialises missing output gradients as zeros; None only with materialize_grads(False)
                    g_pix = g_colors[i] if g_colors[i] is not None else torch.zeros((3, H, W), **f32)
                    if i == 0 and g_others is not None:
                        g_oth = g_others                            # allmap was returned once, by pass 0
                    else:
                        if zeros_others is None:
                            zeros_others = torch.zeros((7, H, W), **f32)
                        g_oth = zeros_others
                    g_pix, g_oth = _dev_f32(g_pix, "dL_dout_color"), _dev_f32(g_oth, "dL_dout_others")
                    _lib.check(L.surfel_pass_set_colors(P, _ptr(colors[i]), _ptr(geom_buf), st), "surfel_pass_set_colors")
                    _lib.check(L.surfel_pass_backward_blend(P, W, H, ctx.num_rendered, _ptr(bgs[i]), _ptr(geom_buf),
                                                            _ptr(bin_buf), _ptr(img_buf), _ptr(g_pix), _ptr(g_oth),
                                                            _ptr(scratch), int(i == 0), st, int(bool(s.debug))),
                               "surfel_pass_backward_blend")
                    if need_color[i]:
                        d_colors[i] = torch.empty((P, 3), **f32)
                    _lib.check(L.surfel_pass_take_color_grad(P, _ptr(scratch), _ptr(d_colors[i]), st),
                               "surfel_pass_take_color_grad")
                have_sr = scales.numel() != 0
                _lib.check(L.surfel_pass_backward_geometry(
                    P, W, H, _ptr(means3D), _ptr(scales) if have_sr else None, _ptr(rotations) if have_sr else None,
                    _ptr(cov3Ds_precomp), _ptr(s.viewmatrix.contiguous()), _ptr(s.projmatrix.contiguous()),
                    _ptr(s.campos.contiguous()), float(s.tanfovx), float(s.tanfovy), _ptr(radii), _ptr(geom_buf),
                    _ptr(scratch), _ptr(d_means2D), None, _ptr(d_opacity), _ptr(d_color_last), _ptr(d_means3D),
                    _ptr(d_transMat), _ptr(d_scales), _ptr(d_rot), st, int(bool(s.debug))), "surfel_pass_backward_geometry")
            for i in range(n):
                if need_color[i] and d_colors[i] is None:   # P == 0
                    d_colors[i] = torch.zeros((P, 3), **f32)
        return (d_means3D, d_means2D, d_opacity, d_scales, d_rot, d_transMat, None, None, *d_colors)


def rasterize_color_passes(raster_settings, means3D, means2D, opacities, colors_precomp: Sequence[torch.Tensor],
                           bgs: Optional[Sequence[torch.Tensor]] = None, scales=None, rotations=None, cov3D_precomp=None):
    """``[GaussianRasterizer(settings with bg=bgs[i])(colors_precomp=colors_precomp[i], ...) for i]`` in one call.

    Returns ``(colors, radii, allmap)`` where ``colors`` is the list of ``[3,H,W]`` images; ``radii`` and ``allmap`` are
    those of any single call (they do not depend on the colours).  ``bgs`` defaults to ``raster_settings.bg`` for every pass.
    """
    have_sr = scales is not None or rotations is not None
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (have_sr and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    colors_precomp = list(colors_precomp)
    if len(colors_precomp) == 0:
        raise Exception('Please provide at least one set of precomputed colors!')
    bgs = [raster_settings.bg] * len(colors_precomp) if bgs is None else list(bgs)
    empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
    scales = empty if scales is None else scales
    rotations = empty if rotations is None else rotations
    cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
    out = _RasterizeColorPasses.apply(means3D, means2D, opacities, scales, rotations, cov3D_precomp, raster_settings,
                                      tuple(bgs), *colors_precomp)
    return list(out[2:]), out[0], out[1]
