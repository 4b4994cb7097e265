"""CPU (numpy, float64 internally) restatement of the GaussianModel activations -- the caller-side row in front of the
rasterizer's projection kernel.

TEST INFRASTRUCTURE ONLY (same rules as surfel_oracle.c): never imported by streetunveiler_b200/.

Follows scene/gaussian_model.py of the reference:
  :31-39     scaling_activation = torch.exp, opacity_activation = torch.sigmoid,
             rotation_activation = torch.nn.functional.normalize   (x / max(||x||_2, 1e-12) along dim 1)
  :101-127   get_scaling, get_rotation, get_opacity, get_features = cat((_features_dc, _features_rest), dim=1)
The backward is written by hand (what autograd does for those lines).  Pinned against tests/golden/activation_*.npz,
produced by the reference's own GaussianModel properties under autograd on CPU (tests/golden/make_golden_activation.py).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12


def forward(scaling_raw, rotation_raw, opacity_raw, features_dc, features_rest):
    s, q, o = (np.asarray(a, np.float64) for a in (scaling_raw, rotation_raw, opacity_raw))
    n = np.sqrt((q * q).sum(1, keepdims=True))
    return {"scaling": np.exp(s), "rotation": q / np.maximum(n, EPS), "opacity": 1.0 / (1.0 + np.exp(-o)),
            "features": np.concatenate([np.asarray(features_dc, np.float64), np.asarray(features_rest, np.float64)], 1)}


def backward(scaling_raw, rotation_raw, opacity_raw, g):
    """g: upstream gradients of the four outputs -> gradients of the five raw parameters."""
    s, q, o = (np.asarray(a, np.float64) for a in (scaling_raw, rotation_raw, opacity_raw))
    # autograd multiplies by the SAVED fp32 outputs (ExpBackward: grad * result; SigmoidBackward: grad * (1 - y) * y), so
    # 1 - y carries the rounding of a saturated fp32 sigmoid -- restated here, not "improved"
    y = (1.0 / (1.0 + np.exp(-o))).astype(np.float32).astype(np.float64)
    e = np.exp(s).astype(np.float32).astype(np.float64)
    n = np.sqrt((q * q).sum(1, keepdims=True))
    d = np.maximum(n, EPS)
    gq = np.asarray(g["rotation"], np.float64)
    dot = (gq * q).sum(1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where((n >= EPS) & (n > 0), dot / (d * d * n), 0.0)
    gf = np.asarray(g["features"], np.float64)
    return {"scaling_raw": np.asarray(g["scaling"], np.float64) * e, "rotation_raw": gq / d - q * k,
            "opacity_raw": np.asarray(g["opacity"], np.float64) * (1.0 - y) * y, "features_dc": gf[:, :1], "features_rest": gf[:, 1:]}
