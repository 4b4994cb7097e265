"""Class-probability pass: every semantic class of ``render_semantic`` in ONE rasterizer traversal.

The reference renders class probabilities as one-hot "colours", three classes per ``GaussianRasterizer`` call
(gaussian_renderer/__init__.py:417-446): two complete forward+backward rasterizations for the six classes it trains on.
``rasterize_class_probabilities`` takes the per-Gaussian label instead and accumulates up to 8 class channels in a single
blend pass (``surfel_classes_*`` in the C ABI); images are bit-identical to the one-hot formulation (``1 * w`` and
``0 * w`` are exact) and the gradients with respect to positions, opacities, scales and rotations equal the sum over its
passes.  Labels are not trainable.  There is no fallback: CPU tensors or a failing call raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ._C import _dev_f32, _ptr, _stream

MAX_CLASSES = 8


class _RasterizeClasses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, scales, rotations, cov3Ds_precomp, labels, bg_probs, raster_settings):
        s = raster_settings
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        P = int(means3D.size(0))
        n = int(bg_probs.numel())
        if not 1 <= n <= MAX_CLASSES:
            raise RuntimeError(f"between 1 and {MAX_CLASSES} classes per pass")
        if not labels.is_cuda or labels.dtype != torch.int32 or labels.numel() != P:
            raise RuntimeError("labels must be a CUDA int32 tensor with one entry per Gaussian")
        labels = labels.contiguous()
        means3D, opacities = _dev_f32(means3D, "means3D"), _dev_f32(opacities, "opacity")
        scales, rotations = _dev_f32(scales, "scales"), _dev_f32(rotations, "rotations")
        cov3Ds_precomp = _dev_f32(cov3Ds_precomp, "transMat_precomp")
        bg_probs = _dev_f32(bg_probs, "background")
        view, proj, campos = _dev_f32(s.viewmatrix, "viewmatrix"), _dev_f32(s.projmatrix, "projmatrix"), _dev_f32(s.campos, "campos")
        H, W = int(s.image_height), int(s.image_width)
        dev = means3D.device
        L = _lib.lib()
        with torch.cuda.device(dev):
            st = _stream()
            u8 = dict(dtype=torch.uint8, device=dev)
            probs = torch.empty((n, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            img_buf = torch.empty((_lib.size(L.surfel_image_bytes(W, H), "surfel_image_bytes"),), **u8)
            geom_buf = torch.empty((_lib.size(L.surfel_geometry_bytes(P), "surfel_geometry_bytes") if P else 0,), **u8)
            R = C.c_int64(0)
            if P:
                dummy_colors = torch.zeros((P, 3), dtype=torch.float32, device=dev)   # the projection wants a colour source
                _lib.check(L.surfel_forward_prepare(
                    P, 0, 0, W, H, _ptr(means3D), None, _ptr(dummy_colors), _ptr(opacities), _ptr(scales),
                    float(s.scale_modifier), _ptr(rotations), _ptr(cov3Ds_precomp), _ptr(view), _ptr(proj), _ptr(campos),
                    float(s.tanfovx), float(s.tanfovy), int(bool(s.prefiltered)), _ptr(radii), _ptr(geom_buf), C.byref(R), st,
                    int(bool(s.debug))), "surfel_forward_prepare")
            num_rendered = int(R.value)
            bin_buf = torch.empty((_lib.size(L.surfel_binning_bytes(num_rendered), "surfel_binning_bytes") if num_rendered else 0,), **u8)
            _lib.check(L.surfel_forward_bin(P, W, H, num_rendered, _ptr(radii), _ptr(geom_buf), _ptr(bin_buf), _ptr(img_buf), st,
                                            int(bool(s.debug))), "surfel_forward_bin")
            _lib.check(L.surfel_classes_set_labels(P, _ptr(labels), _ptr(geom_buf), st), "surfel_classes_set_labels")
            _lib.check(L.surfel_classes_render(P, W, H, num_rendered, n, _ptr(bg_probs), _ptr(geom_buf), _ptr(bin_buf),
                                               _ptr(img_buf), _ptr(probs), st, int(bool(s.debug))), "surfel_classes_render")
        ctx.raster_settings, ctx.num_rendered, ctx.n = s, num_rendered, n
        ctx.save_for_backward(means3D, scales, rotations, cov3Ds_precomp, radii, labels, bg_probs, geom_buf, bin_buf, img_buf,
                              view, proj, campos)
        ctx.mark_non_differentiable(radii)
        return probs, radii

    @staticmethod
    def backward(ctx, g_probs, g_radii):
        s, n = ctx.raster_settings, ctx.n
        (means3D, scales, rotations, cov3Ds_precomp, radii, labels, bg_probs, geom_buf, bin_buf, img_buf, view, proj,
         campos) = ctx.saved_tensors
        P = int(means3D.size(0))
        H, W = int(s.image_height), int(s.image_width)
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        if g_probs is None:
            g_probs = torch.zeros((n, H, W), **f32)
        g_probs = _dev_f32(g_probs, "dL_dprobs")
        L = _lib.lib()
        d_means3D, d_means2D = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32)
        d_opacity, d_transMat = torch.empty((P, 1), **f32), torch.empty((P, 9), **f32)
        d_scales, d_rot, d_color = torch.empty((P, 2), **f32), torch.empty((P, 4), **f32), torch.empty((P, 3), **f32)
        if P:
            with torch.cuda.device(dev):
                st = _stream()
                scratch = torch.empty((_lib.size(L.surfel_grad_scratch_bytes(P), "surfel_grad_scratch_bytes"),),
                                      dtype=torch.uint8, device=dev)
                _lib.check(L.surfel_classes_set_labels(P, _ptr(labels), _ptr(geom_buf), st), "surfel_classes_set_labels")
                _lib.check(L.surfel_classes_backward_blend(P, W, H, ctx.num_rendered, n, _ptr(bg_probs), _ptr(geom_buf),
                                                           _ptr(bin_buf), _ptr(img_buf), _ptr(g_probs), _ptr(scratch), 1, st,
                                                           int(bool(s.debug))), "surfel_classes_backward_blend")
                have_sr = scales.numel() != 0
                _lib.check(L.surfel_pass_backward_geometry(
                    P, W, H, _ptr(means3D), _ptr(scales) if have_sr else None, _ptr(rotations) if have_sr else None,
                    _ptr(cov3Ds_precomp), _ptr(view), _ptr(proj), _ptr(campos), float(s.tanfovx), float(s.tanfovy),
                    _ptr(radii), _ptr(geom_buf), _ptr(scratch), _ptr(d_means2D), None, _ptr(d_opacity), _ptr(d_color),
                    _ptr(d_means3D), _ptr(d_transMat), _ptr(d_scales), _ptr(d_rot), st, int(bool(s.debug))),
                    "surfel_pass_backward_geometry")
        return d_means3D, d_means2D, d_opacity, d_scales, d_rot, d_transMat, None, None, None


def rasterize_class_probabilities(raster_settings, means3D, means2D, opacities, labels, bg_probs, scales=None,
                                  rotations=None, cov3D_precomp=None):
    """-> ``(probs [n_classes,H,W], radii [P])`` with ``n_classes = len(bg_probs)``: channel k is what the reference's
    rasterizer renders for the colour ``(labels == k)`` over background ``bg_probs[k]``."""
    have_sr = scales is not None or rotations is not None
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (have_sr and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
    scales = empty if scales is None else scales
    rotations = empty if rotations is None else rotations
    cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
    return _RasterizeClasses.apply(means3D, means2D, opacities, scales, rotations, cov3D_precomp, labels, bg_probs,
                                   raster_settings)
