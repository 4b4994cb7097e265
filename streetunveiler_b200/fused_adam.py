"""Fused per-iteration parameter update ("next" row 4).

``FusedAdam`` is a drop-in for the optimiser the reference builds at scene/gaussian_model.py:171-180
(``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` over six single-tensor groups): same constructor arguments, same
``param_groups`` and per-parameter ``state`` keys (``step``, ``exp_avg``, ``exp_avg_sq``), so the reference's optimiser
surgery (``replace_tensor_to_optimizer``, ``_prune_optimizer``, ``cat_tensors_to_optimizer``, gaussian_model.py:384-470),
``state_dict()`` / ``load_state_dict()`` and the learning-rate schedule (:218-224) work on it unchanged.  ``step()`` is
one CUDA kernel launch for all groups (csrc/adam.cu) instead of PyTorch's ~12 multi-tensor kernels.

``densification_stats`` fuses train.py:168 and gaussian_model.py:555-557 (five indexed PyTorch statements, each a
boolean-mask gather/scatter with a host sync for the mask size) into one launch.

No fallback: CPU tensors, a missing library or a failing call raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_MAX_GROUPS = 8


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise RuntimeError("FusedAdam implements the reference's configuration: weight_decay=0, amsgrad=False")
        if not 0.0 <= lr or not 0.0 <= eps or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # one launch per distinct (betas, eps, device); the reference has exactly one
        batches = {}
        keep = []
        stepped = []   # step counters advance only after every parameter validated and every launch succeeded
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                state = self.state[p]
                if len(state) == 0:   # torch/optim/adam.py _init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if m.shape != p.shape or v.shape != p.shape or not m.is_contiguous() or not v.is_contiguous():
                    raise RuntimeError("optimizer state does not match its parameter")
                g = p.grad.contiguous()
                if g.dtype != torch.float32:
                    g = g.float()
                keep.append(g)
                key = (group["betas"][0], group["betas"][1], group["eps"], p.device)
                batches.setdefault(key, []).append(
                    _lib.AdamGroup(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(group["lr"]),
                                   int(state["step"]) + 1))
                stepped.append(state)
        L = _lib.lib()
        for (b1, b2, eps, dev), items in batches.items():
            with torch.cuda.device(dev):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for i in range(0, len(items), _MAX_GROUPS):
                    part = items[i:i + _MAX_GROUPS]
                    arr = (_lib.AdamGroup * len(part))(*part)
                    _lib.check(L.surfel_adam_step(len(part), arr, float(b1), float(b2), float(eps), st), "surfel_adam_step")
        for state in stepped:
            state["step"] += 1
        del keep
        return loss


@torch.no_grad()
def densification_stats(radii, viewspace_point_grad, max_radii2D, xyz_gradient_accum, denom):
    """In place, for the Gaussians with ``radii > 0`` (the reference's ``visibility_filter``):
    ``max_radii2D = max(max_radii2D, radii)`` (train.py:168), ``xyz_gradient_accum += ||grad||``, ``denom += 1``
    (scene/gaussian_model.py:555-557).  ``viewspace_point_grad`` is ``viewspace_point_tensor.grad`` [P,3]."""
    P = int(radii.shape[0])
    for name, t, n, dt in (("radii", radii, P, torch.int32), ("viewspace_point_grad", viewspace_point_grad, 3 * P, torch.float32),
                           ("max_radii2D", max_radii2D, P, torch.float32), ("xyz_gradient_accum", xyz_gradient_accum, P, torch.float32),
                           ("denom", denom, P, torch.float32)):
        if not t.is_cuda or t.dtype != dt or t.numel() != n or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous CUDA tensor of {n} x {dt}")
    with torch.cuda.device(radii.device):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().surfel_densification_stats(P, C.c_void_p(radii.data_ptr()), C.c_void_p(viewspace_point_grad.data_ptr()),
                                                         C.c_void_p(max_radii2D.data_ptr()), C.c_void_p(xyz_gradient_accum.data_ptr()),
                                                         C.c_void_p(denom.data_ptr()), st), "surfel_densification_stats")
