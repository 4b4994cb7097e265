"""Fused training-loss block ("next" row 3).

Drop-in for the reference's ``utils/loss_utils.py`` (``l1_loss`` :17-18, ``ssim`` :33-64 -- same names, arguments and
values) plus the statements of ``train.py:113-136`` that combine them (``training_loss``).  The reference runs ~60
PyTorch kernels per iteration here (sky composite, |.|.mean(), five 11x11 grouped convolutions, the SSIM map and their
autograd, two regulariser means); this module runs three CUDA kernels forward and two backward (csrc/loss.cu) through
the C ABI.  There is no fallback: CPU tensors, a missing library or a failing call raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _image(t, name, channels):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dim() != 3 or t.shape[0] != channels:
        raise RuntimeError(f"{name} must have shape [{channels},H,W], got {tuple(t.shape)}")
    return t.contiguous().float()


class _Photometric(torch.autograd.Function):
    """(render, rend_alpha | None, sky | None, gt) -> tensor [2] = (mean |img1 - gt|, mean ssim_map(img1, gt)) with
    img1 = render + sky * (1 - rend_alpha)."""

    @staticmethod
    def forward(ctx, render, rend_alpha, sky, gt):
        render = _image(render, "network_output / img1", 3)
        gt = _image(gt, "gt / img2", 3)
        if (sky is None) != (rend_alpha is None):
            raise RuntimeError("sky and rend_alpha go together")
        if sky is not None:
            sky, rend_alpha = _image(sky, "sky_image", 3), _image(rend_alpha, "rend_alpha", 1)
        if gt.shape != render.shape or (sky is not None and (sky.shape != render.shape or rend_alpha.shape[1:] != render.shape[1:])):
            raise RuntimeError("image shapes differ")
        _, H, W = render.shape
        dev = render.device
        need_grad = any(ctx.needs_input_grad[:3])
        if ctx.needs_input_grad[3]:
            raise RuntimeError("gradients with respect to the ground-truth image are not supported")
        L = _lib.lib()
        with torch.cuda.device(dev):
            deriv = torch.empty((9, H, W), dtype=torch.float32, device=dev) if need_grad else None
            scratch = torch.empty(_lib.size(L.surfel_loss_scratch_bytes(W, H), "surfel_loss_scratch_bytes"),
                                  dtype=torch.uint8, device=dev)
            means = torch.empty(2, dtype=torch.float32, device=dev)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.surfel_loss_photometric_forward(W, H, _p(render), _p(rend_alpha), _p(sky), _p(gt), _p(deriv),
                                                         _p(scratch), _p(means), st), "surfel_loss_photometric_forward")
        ctx.save_for_backward(render, rend_alpha, sky, gt, deriv)
        return means

    @staticmethod
    def backward(ctx, g_means):
        render, rend_alpha, sky, gt, deriv = ctx.saved_tensors
        _, H, W = render.shape
        up = g_means.contiguous().float()
        d_render = torch.empty_like(render)
        d_alpha = torch.empty_like(rend_alpha) if sky is not None and ctx.needs_input_grad[1] else None
        d_sky = torch.empty_like(sky) if sky is not None and ctx.needs_input_grad[2] else None
        with torch.cuda.device(render.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_loss_photometric_backward(W, H, _p(render), _p(rend_alpha), _p(sky), _p(gt),
                                                                   _p(deriv), _p(up), _p(d_render), _p(d_alpha),
                                                                   _p(d_sky), st), "surfel_loss_photometric_backward")
        return d_render, d_alpha, d_sky, None


class _Regulariser(torch.autograd.Function):
    """(rend_normal, surf_normal, rend_dist) -> tensor [2] = (mean(1 - <rend_normal, surf_normal>), mean(rend_dist))."""

    @staticmethod
    def forward(ctx, rend_normal, surf_normal, rend_dist):
        rn, sn = _image(rend_normal, "rend_normal", 3), _image(surf_normal, "surf_normal", 3)
        dist = _image(rend_dist, "rend_dist", 1)
        if rn.shape != sn.shape or dist.shape[1:] != rn.shape[1:]:
            raise RuntimeError("image shapes differ")
        _, H, W = rn.shape
        L = _lib.lib()
        with torch.cuda.device(rn.device):
            scratch = torch.empty(_lib.size(L.surfel_loss_scratch_bytes(W, H), "surfel_loss_scratch_bytes"),
                                  dtype=torch.uint8, device=rn.device)
            means = torch.empty(2, dtype=torch.float32, device=rn.device)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.surfel_loss_regulariser_forward(W, H, _p(rn), _p(sn), _p(dist), _p(scratch), _p(means), st),
                       "surfel_loss_regulariser_forward")
        ctx.save_for_backward(rn, sn)
        return means

    @staticmethod
    def backward(ctx, g_means):
        rn, sn = ctx.saved_tensors
        _, H, W = rn.shape
        up = g_means.contiguous().float()
        d_rn, d_sn = torch.empty_like(rn), torch.empty_like(sn)
        d_dist = torch.empty((1, H, W), dtype=torch.float32, device=rn.device)
        with torch.cuda.device(rn.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_loss_regulariser_backward(W, H, _p(rn), _p(sn), _p(up), _p(d_rn), _p(d_sn),
                                                                   _p(d_dist), st), "surfel_loss_regulariser_backward")
        return d_rn, d_sn, d_dist


def l1_and_ssim(img1, img2, rend_alpha=None, sky_image=None):
    """Both photometric terms from one kernel: (Ll1, Lssim) of ``img1 [+ sky_image * (1 - rend_alpha)]`` against img2."""
    means = _Photometric.apply(img1, rend_alpha, sky_image, img2)
    return means[0], means[1]


def l1_loss(network_output, gt):
    """utils/loss_utils.py:17-18."""
    return l1_and_ssim(network_output, gt)[0]


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:33-64 for the configuration every reference call site uses (window 11, mean over all)."""
    if window_size != 11 or not size_average:
        raise RuntimeError("the fused ssim implements window_size=11, size_average=True (the reference's call sites)")
    return l1_and_ssim(img1, img2)[1]


class _TrainingLoss(torch.autograd.Function):
    """(render, rend_alpha | None, sky | None, gt, rend_normal, surf_normal, rend_dist, lambdas) ->
    tensor [5] = (loss, l1, ssim, Lnormal, Ldist) in three launches forward and two backward; the upstream gradient of
    all five values is mixed with the lambdas on the device."""

    @staticmethod
    def forward(ctx, render, rend_alpha, sky, gt, rend_normal, surf_normal, rend_dist, lambda_dssim, lambda_normal,
                lambda_dist):
        render, gt = _image(render, "render", 3), _image(gt, "gt_image", 3)
        rn, sn = _image(rend_normal, "rend_normal", 3), _image(surf_normal, "surf_normal", 3)
        dist = _image(rend_dist, "rend_dist", 1)
        if (sky is None) != (rend_alpha is None):
            raise RuntimeError("sky and rend_alpha go together")
        if sky is not None:
            sky, rend_alpha = _image(sky, "sky_image", 3), _image(rend_alpha, "rend_alpha", 1)
        shapes = [gt.shape, rn.shape, sn.shape] + ([sky.shape] if sky is not None else [])
        planes = [dist.shape[1:]] + ([rend_alpha.shape[1:]] if sky is not None else [])
        if any(x != render.shape for x in shapes) or any(x != render.shape[1:] for x in planes):
            raise RuntimeError("image shapes differ")
        if ctx.needs_input_grad[3]:
            raise RuntimeError("gradients with respect to the ground-truth image are not supported")
        _, H, W = render.shape
        dev = render.device
        need_grad = any(ctx.needs_input_grad[:7])
        lam = (float(lambda_dssim), float(lambda_normal), float(lambda_dist))
        L = _lib.lib()
        with torch.cuda.device(dev):
            deriv = torch.empty((9, H, W), dtype=torch.float32, device=dev) if need_grad else None
            scratch = torch.empty(_lib.size(L.surfel_loss_scratch_bytes(W, H), "surfel_loss_scratch_bytes"),
                                  dtype=torch.uint8, device=dev)
            out5 = torch.empty(5, dtype=torch.float32, device=dev)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.surfel_loss_training_forward(W, H, _p(render), _p(rend_alpha), _p(sky), _p(gt), _p(rn), _p(sn),
                                                      _p(dist), lam[0], lam[1], lam[2], _p(deriv), _p(scratch), _p(out5),
                                                      st), "surfel_loss_training_forward")
        ctx.save_for_backward(render, rend_alpha, sky, gt, rn, sn, deriv)
        ctx.lam = lam
        return out5

    @staticmethod
    def backward(ctx, g_out5):
        render, rend_alpha, sky, gt, rn, sn, deriv = ctx.saved_tensors
        _, H, W = render.shape
        g = g_out5.contiguous().float()
        need = ctx.needs_input_grad
        d_render = torch.empty_like(render)
        d_alpha = torch.empty_like(rend_alpha) if sky is not None and need[1] else None
        d_sky = torch.empty_like(sky) if sky is not None and need[2] else None
        d_rn, d_sn = torch.empty_like(rn), torch.empty_like(sn)
        d_dist = torch.empty((1, H, W), dtype=torch.float32, device=render.device)
        with torch.cuda.device(render.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_loss_training_backward(
                W, H, _p(render), _p(rend_alpha), _p(sky), _p(gt), _p(rn), _p(sn), _p(deriv), _p(g), ctx.lam[0], ctx.lam[1],
                ctx.lam[2], _p(d_render), _p(d_alpha), _p(d_sky), _p(d_rn), _p(d_sn), _p(d_dist), st),
                "surfel_loss_training_backward")
        return d_render, d_alpha, d_sky, None, d_rn, d_sn, d_dist, None, None, None


def training_loss(render_pkg, sky_image, gt_image, lambda_dssim, lambda_normal=0.0, lambda_dist=0.0):
    """train.py:113-136: ``(loss, loss_dict)`` with the reference's keys ``l1, ssim, Lnormal, Ldist``.
    ``sky_image`` may be None (composite = render).  The shrink term (train.py:138-141) is a mean over the opacity
    parameters, not an image operation, and stays with the caller."""
    alpha = render_pkg["rend_alpha"] if sky_image is not None else None
    out5 = _TrainingLoss.apply(render_pkg["render"], alpha, sky_image, gt_image, render_pkg["rend_normal"],
                               render_pkg["surf_normal"], render_pkg["rend_dist"], lambda_dssim, lambda_normal, lambda_dist)
    loss, Ll1, Lssim, normal_loss, dist_loss = out5.unbind(0)
    return loss, {"l1": Ll1, "ssim": Lssim, "Lnormal": normal_loss, "Ldist": dist_loss}
