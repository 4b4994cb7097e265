"""Pins the training-loss oracle (oracle/loss_oracle.py) against golden vectors produced by the reference's own
utils/loss_utils.py functions under autograd on CPU (tests/golden/make_golden_loss.py)."""
import os

import numpy as np
import pytest

import harness as hz
from loss_cases import GRAD_KEYS, LOSS_CASES, build_loss_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_window_matches_reference_values():
    from oracle import loss_oracle as lo
    w = lo.window_1d()
    assert w.dtype == np.float32 and w.shape == (11,)
    assert abs(float(w.sum()) - 1.0) < 1e-6 and np.array_equal(w, w[::-1]) and w.argmax() == 5


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_oracle_matches_reference(name):
    from oracle import loss_oracle as lo
    g = dict(np.load(os.path.join(GOLD, f"loss_{name}.npz")))
    c = build_loss_case(name)
    pkg = {k: v.numpy() for k, v in c["pkg"].items()}
    sky = c["sky"].numpy() if c["sky"] is not None else None
    out, grads = lo.training_loss(pkg, sky, c["gt"].numpy(), c["lambda_dssim"], c["lambda_normal"], c["lambda_dist"])
    for k in ("loss", "l1", "ssim", "Lnormal", "Ldist"):
        assert abs(out[k] - float(g[k])) <= 2e-6 * max(1.0, abs(float(g[k]))), (k, out[k], float(g[k]))
    for k in GRAD_KEYS:
        if grads[k] is None:
            assert "g_" + k not in g or not np.any(g["g_" + k])
            continue
        ref = g["g_" + k]
        assert grads[k].shape == ref.shape, k
        assert hz.rel_err(grads[k], ref) <= 1e-4, (k, hz.rel_err(grads[k], ref))
