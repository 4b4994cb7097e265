"""Golden vectors for the training-loss block, generated on CPU in this container with the
reference's OWN functions ``utils.loss_utils.l1_loss`` / ``ssim`` (utils/loss_utils.py:17-64) under autograd,
combined by the statements of train.py:113-136 (restated below verbatim in meaning; train.py itself is a script
with a CUDA-only training loop and cannot be imported as a function).

    python tests/golden/make_golden_loss.py        # needs /root/reference, no GPU

Outputs: tests/golden/loss_*.npz (inputs are regenerated from seeds by tests/golden/loss_cases.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from loss_cases import LOSS_CASES, build_loss_case  # noqa: E402

REF = "/root/reference"


def run_reference_loss(lu, case):
    pkg = {k: v.clone().requires_grad_(True) for k, v in case["pkg"].items()}
    sky = case["sky"].clone().requires_grad_(True) if case["sky"] is not None else None
    gt = case["gt"]
    # train.py:113-117
    composite_image = pkg["render"] + sky * (1 - pkg["rend_alpha"]) if sky is not None else pkg["render"]
    Ll1 = lu.l1_loss(composite_image, gt)
    Lssim = lu.ssim(composite_image, gt)
    loss = (1.0 - case["lambda_dssim"]) * Ll1 + case["lambda_dssim"] * (1.0 - Lssim)
    # train.py:122-129
    normal_error = (1 - (pkg["rend_normal"] * pkg["surf_normal"]).sum(dim=0))[None]
    normal_loss = case["lambda_normal"] * (normal_error).mean()
    loss = loss + normal_loss
    # train.py:133-136
    dist_loss = case["lambda_dist"] * (pkg["rend_dist"]).mean()
    loss = loss + dist_loss
    loss.backward()
    out = {"loss": loss.item(), "l1": Ll1.item(), "ssim": Lssim.item(), "Lnormal": float(normal_loss),
           "Ldist": float(dist_loss)}
    out = {k: np.float64(v) for k, v in out.items()}
    for k, v in pkg.items():
        out["g_" + k] = v.grad.numpy() if v.grad is not None else np.zeros(v.shape, np.float32)
    if sky is not None:
        out["g_sky"] = sky.grad.numpy()
    return out


def main():
    sys.path.insert(0, REF)
    from utils import loss_utils as lu
    for name in LOSS_CASES:
        out = run_reference_loss(lu, build_loss_case(name))
        np.savez_compressed(os.path.join(HERE, f"loss_{name}.npz"), **out)
        print(name, {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


if __name__ == "__main__":
    main()
