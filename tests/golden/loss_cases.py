"""Seeded inputs of the training-loss golden cases (see make_golden_loss.py)."""
import torch

LOSS_CASES = ["photo_sky", "no_sky_tiny", "full_block", "exact_match"]
PKG_KEYS = ["render", "rend_alpha", "rend_normal", "surf_normal", "rend_dist"]
GRAD_KEYS = ["render", "rend_alpha", "sky", "rend_normal", "surf_normal", "rend_dist"]


def _images(H, W, seed, sky):
    g = torch.Generator().manual_seed(seed)
    f64 = dict(generator=g, dtype=torch.float64)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(xx / 5.0 + c) * torch.cos(yy / (4.0 + c)) for c in range(3)])
    gt = (base + 0.05 * torch.randn(3, H, W, **f64)).clamp(0, 1)
    render = (base + 0.15 * torch.randn(3, H, W, **f64)).clamp(0, 1.2)
    render[:, H // 2:, : W // 3] = 0.0                      # flat region: sigma ~ 0, C2 decides
    alpha = (0.2 + 0.8 * torch.rand(1, H, W, **f64)).clamp(0, 1)
    alpha[:, : H // 4, W // 2:] = 0.0
    n1 = torch.randn(3, H, W, **f64)
    n1 = n1 / n1.norm(dim=0, keepdim=True)
    n2 = n1 + 0.3 * torch.randn(3, H, W, **f64)
    n2 = n2 / n2.norm(dim=0, keepdim=True) * alpha
    pkg = {"render": render * alpha, "rend_alpha": alpha, "rend_normal": n1 * alpha, "surf_normal": n2,
           "rend_dist": 0.02 * torch.rand(1, H, W, **f64) * alpha}
    pkg = {k: v.float().contiguous() for k, v in pkg.items()}
    sky_img = torch.rand(3, H, W, **f64).float().contiguous() if sky else None
    return pkg, sky_img, gt.float().contiguous()


def build_loss_case(name):
    """-> dict(pkg, sky, gt, lambda_dssim, lambda_normal, lambda_dist)"""
    if name == "photo_sky":         # train.py defaults before the regularisers switch on
        pkg, sky, gt = _images(53, 75, 11, True)
        lam = (0.2, 0.0, 0.0)
    elif name == "no_sky_tiny":     # image smaller than the 11x11 window in one direction, no sky composite
        pkg, sky, gt = _images(7, 37, 12, False)
        lam = (0.2, 0.05, 100.0)
    elif name == "full_block":      # everything on, sizes that are not multiples of the 32x32 CUDA tile
        pkg, sky, gt = _images(70, 97, 13, True)
        lam = (0.2, 0.05, 100.0)
    elif name == "exact_match":     # composite == gt on a patch: |.| has a kink there, torch's sign(0) = 0
        pkg, sky, gt = _images(40, 64, 14, True)
        comp = pkg["render"] + sky * (1 - pkg["rend_alpha"])
        gt[:, 5:20, 8:40] = comp[:, 5:20, 8:40]
        lam = (0.35, 0.05, 10.0)
    else:
        raise KeyError(name)
    return dict(pkg=pkg, sky=sky, gt=gt, lambda_dssim=lam[0], lambda_normal=lam[1], lambda_dist=lam[2])
