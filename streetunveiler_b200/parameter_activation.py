"""Fused GaussianModel activations and SH packing -- the caller-side row in front of the rasterizer.

Every ``render()`` of the reference starts by evaluating four ``GaussianModel`` properties
(scene/gaussian_model.py:101-127): ``get_scaling = exp(_scaling)``, ``get_rotation = normalize(_rotation)``,
``get_opacity = sigmoid(_opacity)`` and ``get_features = cat((_features_dc, _features_rest), dim=1)`` -- about a dozen
PyTorch kernels forward and fifteen backward, two of which copy the whole 192 B/Gaussian SH block.  ``activate`` computes
all four with one CUDA kernel each way (csrc/activate.cu); ``ActivatedGaussians`` wraps a reference ``GaussianModel`` so
that ``render()`` / ``render_semantic()`` callers see the same property names.  No fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


def _dev(t, name, shape_tail):
    """float32 CUDA tensor of shape [P, *shape_tail] (None = any size) -> contiguous."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    ok = t.dtype == torch.float32 and t.dim() == 1 + len(shape_tail) and all(
        want is None or want == got for want, got in zip(shape_tail, t.shape[1:]))
    if not ok:
        tail = "".join("," + ("*" if w is None else str(w)) for w in shape_tail)
        raise RuntimeError(f"{name} must be float32 with shape [P{tail}], got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


class _Activate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scaling_raw, rotation_raw, opacity_raw, features_dc, features_rest):
        s = _dev(scaling_raw, "_scaling", (2,))
        q = _dev(rotation_raw, "_rotation", (4,))
        o = _dev(opacity_raw, "_opacity", (1,))
        dc = _dev(features_dc, "_features_dc", (1, 3))
        rest = _dev(features_rest, "_features_rest", (None, 3))
        P, R = int(s.shape[0]), int(rest.shape[1])
        if any(int(t.shape[0]) != P for t in (q, o, dc, rest)):
            raise RuntimeError("parameter tensors disagree on the number of Gaussians")
        f32 = dict(dtype=torch.float32, device=s.device)
        scaling, rotation, opacity = torch.empty((P, 2), **f32), torch.empty((P, 4), **f32), torch.empty((P, 1), **f32)
        features = torch.empty((P, 1 + R, 3), **f32)
        with torch.cuda.device(s.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_activate_forward(P, R, _p(s), _p(q), _p(o), _p(dc), _p(rest), _p(scaling), _p(rotation),
                                                          _p(opacity), _p(features), st), "surfel_activate_forward")
        ctx.save_for_backward(q, scaling, opacity)
        ctx.R = R
        return scaling, rotation, opacity, features

    @staticmethod
    def backward(ctx, g_scaling, g_rotation, g_opacity, g_features):
        q, scaling, opacity = ctx.saved_tensors
        P, R = int(q.shape[0]), ctx.R
        f32 = dict(dtype=torch.float32, device=q.device)
        zeros = lambda shape: torch.zeros(shape, **f32)
        g_scaling = zeros((P, 2)) if g_scaling is None else g_scaling.contiguous().float()
        g_rotation = zeros((P, 4)) if g_rotation is None else g_rotation.contiguous().float()
        g_opacity = zeros((P, 1)) if g_opacity is None else g_opacity.contiguous().float()
        g_features = None if g_features is None else g_features.contiguous().float()
        d_s, d_q, d_o = torch.empty((P, 2), **f32), torch.empty((P, 4), **f32), torch.empty((P, 1), **f32)
        if g_features is None:
            d_dc, d_rest = zeros((P, 1, 3)), zeros((P, R, 3))
        else:
            d_dc, d_rest = torch.empty((P, 1, 3), **f32), torch.empty((P, R, 3), **f32)
        with torch.cuda.device(q.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.lib().surfel_activate_backward(P, R, _p(q), _p(scaling), _p(opacity), _p(g_scaling), _p(g_rotation),
                                                           _p(g_opacity), _p(g_features), _p(d_s), _p(d_q), _p(d_o), _p(d_dc),
                                                           _p(d_rest), st), "surfel_activate_backward")
        return d_s, d_q, d_o, d_dc, d_rest


def activate(scaling_raw, rotation_raw, opacity_raw, features_dc, features_rest):
    """-> ``(get_scaling [P,2], get_rotation [P,4], get_opacity [P,1], get_features [P,1+R,3])`` of
    scene/gaussian_model.py:101-127 for the raw parameters ``_scaling, _rotation, _opacity, _features_dc, _features_rest``."""
    return _Activate.apply(scaling_raw, rotation_raw, opacity_raw, features_dc, features_rest)


class ActivatedGaussians:
    """The view of a reference ``GaussianModel`` that ``render()`` needs, with the four activated properties computed
    once by the fused kernel.  Everything else (``get_xyz``, ``active_sh_degree``, semantics, ...) is forwarded."""

    def __init__(self, pc):
        self._pc = pc
        (self.get_scaling, self.get_rotation, self.get_opacity, self.get_features) = activate(
            pc._scaling, pc._rotation, pc._opacity, pc._features_dc, pc._features_rest)

    def __getattr__(self, name):
        return getattr(self._pc, name)
