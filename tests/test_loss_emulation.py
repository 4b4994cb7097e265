"""CPU check of the CUDA loss kernels' tile code: streetunveiler_b200/csrc/loss_tile.cuh holds the kernel bodies as
__host__ __device__ functions; tests/emul/loss_emul.cu runs them on the host (one "thread", tile by tile).  This
verifies halo / zero-padding / indexing / chain-rule logic of the very code the GPU runs against the golden vectors of
the reference's loss functions and against the oracle -- the GPU tests then only have to confirm the parallel execution."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as hz
from loss_cases import LOSS_CASES, build_loss_case

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def emul_training_loss(L, case):
    pkg = {k: np.ascontiguousarray(v.numpy()) for k, v in case["pkg"].items()}
    sky = np.ascontiguousarray(case["sky"].numpy()) if case["sky"] is not None else None
    gt = np.ascontiguousarray(case["gt"].numpy())
    _, H, W = gt.shape
    alpha = pkg["rend_alpha"] if sky is not None else None
    deriv = np.full((9, H, W), np.nan, np.float32)
    means = np.zeros(2, np.float32)
    L.emul_loss_photometric_forward(W, H, _p(pkg["render"]), _p(alpha), _p(sky), _p(gt), _p(deriv), _p(means))
    assert np.isfinite(deriv).all()          # every pixel of every map written
    lam = case["lambda_dssim"]
    up = np.array([1.0 - lam, -lam], np.float32)
    d_render = np.full((3, H, W), np.nan, np.float32)
    d_alpha = np.full((1, H, W), np.nan, np.float32) if sky is not None else None
    d_sky = np.full((3, H, W), np.nan, np.float32) if sky is not None else None
    L.emul_loss_photometric_backward(W, H, _p(pkg["render"]), _p(alpha), _p(sky), _p(gt), _p(deriv), _p(up), _p(d_render),
                                     _p(d_alpha), _p(d_sky))
    rmeans = np.zeros(2, np.float32)
    upr = np.array([case["lambda_normal"], case["lambda_dist"]], np.float32)
    d_rn, d_sn = np.full((3, H, W), np.nan, np.float32), np.full((3, H, W), np.nan, np.float32)
    d_dist = np.full((1, H, W), np.nan, np.float32)
    L.emul_loss_regulariser(W, H, _p(pkg["rend_normal"]), _p(pkg["surf_normal"]), _p(pkg["rend_dist"]), _p(upr), _p(rmeans),
                            _p(d_rn), _p(d_sn), _p(d_dist))
    out = {"l1": float(means[0]), "ssim": float(means[1]), "Lnormal": case["lambda_normal"] * float(rmeans[0]),
           "Ldist": case["lambda_dist"] * float(rmeans[1])}
    out["loss"] = (1 - lam) * out["l1"] + lam * (1 - out["ssim"]) + out["Lnormal"] + out["Ldist"]
    grads = {"render": d_render, "rend_alpha": d_alpha, "sky": d_sky, "rend_normal": d_rn, "surf_normal": d_sn,
             "rend_dist": d_dist}
    return out, grads


@pytest.mark.parametrize("name", LOSS_CASES)
def test_cuda_tile_code_on_host_matches_reference_golden(emul, name):
    g = dict(np.load(os.path.join(GOLD, f"loss_{name}.npz")))
    out, grads = emul_training_loss(emul, build_loss_case(name))
    for k in ("loss", "l1", "ssim", "Lnormal", "Ldist"):
        assert abs(out[k] - float(g[k])) <= 5e-6 * max(1.0, abs(float(g[k]))), (k, out[k], float(g[k]))
    for k, v in grads.items():
        if v is None:
            continue
        assert np.isfinite(v).all(), k
        assert hz.rel_err(v, g["g_" + k]) <= 1e-4, (k, hz.rel_err(v, g["g_" + k]))
