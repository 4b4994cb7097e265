"""Parameter activations on CPU: the numpy oracle and the CUDA kernels' bodies compiled for the host
(tests/emul/activate_emul.cu) against golden vectors produced by the reference's own GaussianModel properties under
autograd (tests/golden/make_golden_activation.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as hz
from activation_cases import ACTIVATION_CASES, OUT, RAW, build_activation_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(out, grads, g):
    for k in OUT:
        assert out[k].shape == g[k].shape, k
        assert hz.rel_err(out[k], g[k]) <= 2e-6, (k, hz.rel_err(out[k], g[k]))
    assert np.array_equal(np.asarray(out["features"], np.float32), g["features"])          # a pure copy: exact
    for k in RAW:
        assert grads[k].shape == g["g_" + k].shape, k
        assert hz.rel_err(grads[k], g["g_" + k]) <= 5e-6, (k, hz.rel_err(grads[k], g["g_" + k]))


@pytest.mark.parametrize("name", list(ACTIVATION_CASES))
def test_activation_oracle_matches_reference_model(name):
    from oracle import activation_oracle as ao
    g = dict(np.load(os.path.join(GOLD, f"activation_{name}.npz")))
    c = build_activation_case(name)
    raw = {k: v.numpy() for k, v in c["raw"].items()}
    out = ao.forward(*[raw[k] for k in RAW])
    grads = ao.backward(raw["scaling_raw"], raw["rotation_raw"], raw["opacity_raw"], {k: v.numpy() for k, v in c["upstream"].items()})
    check(out, grads, g)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", list(ACTIVATION_CASES))
def test_cuda_activation_code_on_host_matches_reference_model(emul, name):
    g = dict(np.load(os.path.join(GOLD, f"activation_{name}.npz")))
    c = build_activation_case(name)
    P, R = c["P"], c["R"]
    raw = {k: np.ascontiguousarray(v.numpy()) for k, v in c["raw"].items()}
    up = {k: np.ascontiguousarray(v.numpy()) for k, v in c["upstream"].items()}
    out = {"scaling": np.full((P, 2), np.nan, np.float32), "rotation": np.full((P, 4), np.nan, np.float32),
           "opacity": np.full((P, 1), np.nan, np.float32), "features": np.full((P, 1 + R, 3), np.nan, np.float32)}
    emul.emul_activate_forward(P, R, *[ptr(raw[k]) for k in RAW], *[ptr(out[k]) for k in OUT])
    grads = {k: np.full(raw[k].shape, np.nan, np.float32) for k in RAW}
    emul.emul_activate_backward(P, R, ptr(raw["rotation_raw"]), ptr(out["scaling"]), ptr(out["opacity"]), ptr(up["scaling"]),
                                ptr(up["rotation"]), ptr(up["opacity"]), ptr(up["features"]), ptr(grads["scaling_raw"]),
                                ptr(grads["rotation_raw"]), ptr(grads["opacity_raw"]), ptr(grads["features_dc"]),
                                ptr(grads["features_rest"]))
    assert all(np.isfinite(v).all() for v in list(out.values()) + list(grads.values()))     # every element written
    check(out, grads, g)
