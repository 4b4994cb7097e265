"""CPU (numpy, float64 internally) restatement of the training-loss block -- SURVEY.md section 8(f) row 3.

TEST INFRASTRUCTURE ONLY (same rules as surfel_oracle.c): never imported by streetunveiler_b200/.

Follows the reference line by line:
  utils/loss_utils.py:17-18     l1_loss = mean |a - b|
  utils/loss_utils.py:23-31     gaussian(11, 1.5) window: python-double exp -> float32 tensor -> / float32 sum
  utils/loss_utils.py:33-64     ssim: five zero-padded 11x11 depthwise convolutions (mu1, mu2, E[x^2], E[y^2], E[xy]),
                                C1 = 0.01^2, C2 = 0.03^2, mean over channels and pixels
  train.py:113-117              composite = render + sky * (1 - rend_alpha); Ll1; Lssim;
                                loss = (1 - lambda_dssim) Ll1 + lambda_dssim (1 - Lssim)
  train.py:122-129              normal_loss = lambda_normal * mean(1 - sum_c rend_normal * surf_normal)
  train.py:133-136              dist_loss = lambda_dist * mean(rend_dist)
The backward is written by hand (what autograd does for those lines).  Forward and backward are pinned against
tests/golden/loss_*.npz, produced by the reference's own utils/loss_utils.py functions under autograd on CPU
(tests/golden/make_golden_loss.py).
"""
from __future__ import annotations

import math

import numpy as np

WINDOW = 11
SIGMA = 1.5
C1 = 0.01 ** 2
C2 = 0.03 ** 2


def window_1d() -> np.ndarray:
    """loss_utils.py:23-25 -- float32 values, float32 normalisation."""
    g = np.array([math.exp(-(x - WINDOW // 2) ** 2 / float(2 * SIGMA ** 2)) for x in range(WINDOW)], np.float32)
    return (g / g.sum(dtype=np.float32)).astype(np.float32)


def _blur(img: np.ndarray) -> np.ndarray:
    """Depthwise zero-padded 11x11 Gaussian (separable form of F.conv2d(..., padding=5, groups=C)); img [C,H,W] f64."""
    w = window_1d().astype(np.float64)
    C, H, W = img.shape
    r = WINDOW // 2
    p = np.zeros((C, H, W + 2 * r))
    p[:, :, r:r + W] = img
    h = sum(w[k] * p[:, :, k:k + W] for k in range(WINDOW))
    p = np.zeros((C, H + 2 * r, W))
    p[:, r:r + H, :] = h
    return sum(w[k] * p[:, k:k + H, :] for k in range(WINDOW))


def composite(render, alpha, sky):
    """train.py:113 -- evaluated in float32 with separately rounded 1 - a, product and sum, as the reference's
    three PyTorch ops do (where the composite equals gt exactly, |.| sits on its kink and sign(0) = 0)."""
    render = np.asarray(render, np.float32)
    if sky is None:
        return render.astype(np.float64)
    one_m = (np.float32(1.0) - np.asarray(alpha, np.float32)).astype(np.float32)
    prod = (np.asarray(sky, np.float32) * one_m).astype(np.float32)
    return (render + prod).astype(np.float32).astype(np.float64)


def ssim_terms(x, y):
    mu1, mu2 = _blur(x), _blur(y)
    s1 = _blur(x * x) - mu1 * mu1
    s2 = _blur(y * y) - mu2 * mu2
    s12 = _blur(x * y) - mu1 * mu2
    a = 2 * mu1 * mu2 + C1
    b = 2 * s12 + C2
    c = mu1 * mu1 + mu2 * mu2 + C1
    d = s1 + s2 + C2
    return mu1, mu2, a, b, c, d


def photometric_forward(render, alpha, sky, gt):
    """-> (Ll1, Lssim) as python floats."""
    x, y = composite(render, alpha, sky), np.asarray(gt, np.float64)
    _, _, a, b, c, d = ssim_terms(x, y)
    return float(np.abs(x - y).mean()), float(((a * b) / (c * d)).mean())


def photometric_backward(render, alpha, sky, gt, g_l1, g_ssim):
    """Gradients of g_l1 * Ll1 + g_ssim * Lssim -> (d_render, d_alpha | None, d_sky | None)."""
    x, y = composite(render, alpha, sky), np.asarray(gt, np.float64)
    n = x.size
    mu1, mu2, a, b, c, d = ssim_terms(x, y)
    f = (a * b) / (c * d)
    df_ds1 = -f / d                                  # d ssim_map / d sigma1_sq
    df_ds12 = 2 * a / (c * d)                        # d ssim_map / d sigma12
    df_dmu1 = 2 * mu2 * b / (c * d) - 2 * mu1 * f / c   # direct dependence through a and c
    A = df_dmu1 - 2 * mu1 * df_ds1 - mu2 * df_ds12   # + through sigma1_sq = E[x^2] - mu1^2, sigma12 = E[xy] - mu1 mu2
    # the adjoint of a zero-padded convolution with a symmetric window is the same convolution
    dx = (g_ssim / n) * (_blur(A) + 2 * x * _blur(df_ds1) + y * _blur(df_ds12))
    dx = dx + (g_l1 / n) * np.sign(x - y)
    if sky is None:
        return dx, None, None
    sky = np.asarray(sky, np.float64)
    alpha = np.asarray(alpha, np.float64)
    return dx, -(dx * sky).sum(0, keepdims=True), dx * (1.0 - alpha)


def regulariser_forward(rend_normal, surf_normal, rend_dist):
    """-> (mean(1 - <rend_normal, surf_normal>), mean(rend_dist))  (train.py:125-126,134 before the lambdas)."""
    rn, sn = np.asarray(rend_normal, np.float64), np.asarray(surf_normal, np.float64)
    return float((1.0 - (rn * sn).sum(0)).mean()), float(np.asarray(rend_dist, np.float64).mean())


def regulariser_backward(rend_normal, surf_normal, rend_dist, g_normal, g_dist):
    rn, sn = np.asarray(rend_normal, np.float64), np.asarray(surf_normal, np.float64)
    hw = rn.shape[1] * rn.shape[2]
    return -(g_normal / hw) * sn, -(g_normal / hw) * rn, np.full(np.shape(rend_dist), g_dist / hw)


def training_loss(pkg, sky, gt, lambda_dssim, lambda_normal, lambda_dist):
    """train.py:113-136 -> dict(loss, l1, ssim, Lnormal, Ldist) and the gradients of `loss`."""
    l1, ss = photometric_forward(pkg["render"], pkg["rend_alpha"], sky, gt)
    nm, dm = regulariser_forward(pkg["rend_normal"], pkg["surf_normal"], pkg["rend_dist"])
    out = {"l1": l1, "ssim": ss, "Lnormal": lambda_normal * nm, "Ldist": lambda_dist * dm}
    out["loss"] = (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ss) + out["Lnormal"] + out["Ldist"]
    d_render, d_alpha, d_sky = photometric_backward(pkg["render"], pkg["rend_alpha"], sky, gt, 1.0 - lambda_dssim,
                                                    -lambda_dssim)
    d_rn, d_sn, d_dist = regulariser_backward(pkg["rend_normal"], pkg["surf_normal"], pkg["rend_dist"], lambda_normal,
                                              lambda_dist)
    grads = {"render": d_render, "rend_alpha": d_alpha, "sky": d_sky, "rend_normal": d_rn, "surf_normal": d_sn,
             "rend_dist": d_dist}
    return out, grads
