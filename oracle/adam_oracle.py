"""CPU (numpy float32) restatement of the per-iteration parameter update -- SURVEY.md section 8(f) row 4.

TEST INFRASTRUCTURE ONLY (same rules as surfel_oracle.c): never imported by streetunveiler_b200/.

What the reference runs after every backward:
  train.py:168                      max_radii2D[vis] = max(max_radii2D[vis], radii[vis])      (vis = radii > 0)
  scene/gaussian_model.py:555-557   xyz_gradient_accum[vis] += ||viewspace_points.grad[vis]||_2 ;  denom[vis] += 1
  train.py:197                      gaussians.optimizer.step() with the optimiser of scene/gaussian_model.py:171-180:
                                    torch.optim.Adam(six groups with their own lr, lr=0.0, eps=1e-15)
The Adam arithmetic is a THIRD-PARTY algorithm (PyTorch 2.11, torch/optim/adam.py, betas (0.9, 0.999), no weight decay,
no amsgrad, not maximize); it is restated here in the operation order of torch's implementation
    exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps;  param.addcdiv_(exp_avg, denom, value=-lr / (1 - beta1^t))
and pinned against tests/golden/adam_*.npz, produced by torch.optim.Adam itself on CPU over the reference's group
list (tests/golden/make_golden_adam.py).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """One update of one tensor; `step` is the 1-based count AFTER the increment.  Returns new (param, exp_avg, exp_avg_sq)."""
    p, g, m, v = (np.asarray(a, F) for a in (param, grad, exp_avg, exp_avg_sq))
    m = (m + F(1.0 - beta1) * (g - m)).astype(F)
    v = (v * F(beta2)).astype(F)
    v = (v + (F(1.0 - beta2) * g).astype(F) * g).astype(F)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = F(-(lr / bc1))
    denom = (np.sqrt(v).astype(F) / F(math.sqrt(bc2))).astype(F) + F(eps)
    p = (p + step_size * (m / denom).astype(F)).astype(F)
    return p, m, v


def densification_stats(radii, viewspace_grad, max_radii2D, xyz_gradient_accum, denom):
    """train.py:168 + gaussian_model.py:555-557.  Returns new (max_radii2D [P], xyz_gradient_accum [P,1], denom [P,1])."""
    radii = np.asarray(radii)
    vis = radii > 0
    mr = np.asarray(max_radii2D, F).copy()
    mr[vis] = np.maximum(mr[vis], radii[vis].astype(F))
    g = np.asarray(viewspace_grad, F)
    norm = np.sqrt((g.astype(np.float64) ** 2).sum(-1, keepdims=True)).astype(F)
    acc = np.asarray(xyz_gradient_accum, F).copy()
    acc[vis] = acc[vis] + norm[vis]
    dn = np.asarray(denom, F).copy()
    dn[vis] = dn[vis] + F(1.0)
    return mr, acc, dn
