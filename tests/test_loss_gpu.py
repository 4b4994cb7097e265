"""GPU parity tests of the fused training-loss block (csrc/loss.cu) through the C ABI / loss_block.py."""
import os
from math import exp

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import harness as hz
from loss_cases import GRAD_KEYS, LOSS_CASES, build_loss_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def torch_ssim(img1, img2):
    """Plain PyTorch fp32 formulation of the same op (what utils/loss_utils.py:33-64 computes: 2-D window, grouped conv)."""
    g = torch.tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float()[None, None].expand(3, 1, 11, 11).contiguous().to(img1.device)
    conv = lambda t: F.conv2d(t, window, padding=5, groups=3)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = conv(img1 * img1) - mu1_sq
    s2 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def torch_training_loss(pkg, sky, gt, lam, lam_n, lam_d):
    comp = pkg["render"] + sky * (1 - pkg["rend_alpha"]) if sky is not None else pkg["render"]
    Ll1 = torch.abs(comp - gt).mean()
    Lssim = torch_ssim(comp, gt)
    normal_loss = lam_n * (1 - (pkg["rend_normal"] * pkg["surf_normal"]).sum(dim=0))[None].mean()
    dist_loss = lam_d * pkg["rend_dist"].mean()
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - Lssim) + normal_loss + dist_loss
    return loss, {"l1": Ll1, "ssim": Lssim, "Lnormal": normal_loss, "Ldist": dist_loss}


def run(case, fn):
    dev = torch.device("cuda")
    pkg = {k: v.to(dev).clone().requires_grad_(True) for k, v in case["pkg"].items()}
    sky = case["sky"].to(dev).clone().requires_grad_(True) if case["sky"] is not None else None
    loss, d = fn(pkg, sky, case["gt"].to(dev), case["lambda_dssim"], case["lambda_normal"], case["lambda_dist"])
    loss.backward()
    out = {k: float(v) for k, v in d.items()}
    out["loss"] = float(loss)
    grads = {k: (v.grad.cpu().numpy() if v.grad is not None else None) for k, v in pkg.items()}
    grads["sky"] = sky.grad.cpu().numpy() if sky is not None else None
    return out, grads


def fused(pkg, sky, gt, lam, lam_n, lam_d):
    from streetunveiler_b200.loss_block import training_loss
    return training_loss(pkg, sky, gt, lam, lam_n, lam_d)


@pytest.mark.parametrize("name", LOSS_CASES)
def test_fused_loss_matches_reference_golden(name):
    g = dict(np.load(os.path.join(GOLD, f"loss_{name}.npz")))
    c = build_loss_case(name)
    out, grads = run(c, fused)
    for k in ("loss", "l1", "ssim", "Lnormal", "Ldist"):
        assert abs(out[k] - float(g[k])) <= 5e-6 * max(1.0, abs(float(g[k]))), (k, out[k], float(g[k]))
    for k in GRAD_KEYS:
        if c["sky"] is None and k in ("sky", "rend_alpha"):
            assert grads[k] is None          # no composite: alpha receives nothing from this block
            continue
        assert hz.rel_err(grads[k], g["g_" + k]) <= 1e-4, (k, hz.rel_err(grads[k], g["g_" + k]))


def full_size_case(seed=5):
    from loss_cases import _images
    pkg, sky, gt = _images(1280, 1920, seed, True)
    return dict(pkg=pkg, sky=sky, gt=gt, lambda_dssim=0.2, lambda_normal=0.05, lambda_dist=100.0)


def test_fused_loss_full_size_against_torch_ops_and_oracle():
    from oracle import loss_oracle as lo
    c = full_size_case()
    out, grads = run(c, fused)
    tout, tgrads = run(c, torch_training_loss)
    for k in out:
        assert abs(out[k] - tout[k]) <= 1e-5 * max(1.0, abs(tout[k])), (k, out[k], tout[k])
    for k in GRAD_KEYS:
        assert hz.rel_err(grads[k], tgrads[k]) <= 1e-4, (k, hz.rel_err(grads[k], tgrads[k]))
    pkg = {k: v.numpy() for k, v in c["pkg"].items()}
    oout, ograds = lo.training_loss(pkg, c["sky"].numpy(), c["gt"].numpy(), 0.2, 0.05, 100.0)
    for k in out:
        assert abs(out[k] - oout[k]) <= 1e-5 * max(1.0, abs(oout[k])), (k, out[k], oout[k])
    for k in GRAD_KEYS:
        assert hz.rel_err(grads[k], ograds[k]) <= 1e-4, (k, hz.rel_err(grads[k], ograds[k]))


def test_fused_loss_is_deterministic():
    c = full_size_case(6)
    a, ga = run(c, fused)
    b, gb = run(c, fused)
    assert a == b
    for k in GRAD_KEYS:
        assert np.array_equal(ga[k], gb[k]), k


def test_loss_utils_surface():
    """l1_loss / ssim keep the names, argument order and values of utils/loss_utils.py."""
    from streetunveiler_b200 import loss_block as lb
    dev = torch.device("cuda")
    c = build_loss_case("photo_sky")
    a = c["pkg"]["render"].to(dev).requires_grad_(True)
    gt = c["gt"].to(dev)
    l1, ss = lb.l1_loss(a, gt), lb.ssim(a, gt)
    assert abs(float(l1) - float(torch.abs(a - gt).mean())) <= 1e-6
    assert abs(float(ss) - float(torch_ssim(a.detach(), gt))) <= 5e-6
    (0.8 * l1 + 0.2 * (1 - ss)).backward()
    b = a.detach().clone().requires_grad_(True)
    (0.8 * torch.abs(b - gt).mean() + 0.2 * (1 - torch_ssim(b, gt))).backward()
    assert hz.rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) <= 1e-4
    with pytest.raises(RuntimeError):
        lb.ssim(a, gt, window_size=7)
    with pytest.raises(RuntimeError):
        lb.l1_loss(a.cpu(), gt.cpu())          # no CPU fallback
    with pytest.raises(RuntimeError):
        lb.l1_loss(a[:2], gt[:2])
    with torch.no_grad():                       # inference: no derivative maps are produced
        assert abs(float(lb.ssim(a, gt)) - float(ss)) == 0.0
