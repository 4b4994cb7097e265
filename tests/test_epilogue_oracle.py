"""Pins the render()-epilogue oracle (oracle/epilogue_oracle.py) against golden vectors produced by the
reference's unchanged gaussian_renderer.render on CPU (tests/golden/make_golden_epilogue.py)."""
import os

import numpy as np
import pytest

import harness as hz
from epilogue_cases import EPILOGUE_CASES, KEYS, build_epilogue_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", EPILOGUE_CASES)
def test_epilogue_oracle_matches_reference(name):
    from oracle import epilogue_oracle as eo
    g = dict(np.load(os.path.join(GOLD, f"epilogue_{name}.npz")))
    c = build_epilogue_case(name)
    cam = c["cam"]
    out = eo.forward(c["allmap"].numpy(), cam.viewmatrix.numpy(), cam.fovx, cam.fovy, c["depth_ratio"])
    for k in KEYS:
        assert out[k].shape == g[k].shape, k
        assert hz.rel_err(out[k], g[k]) <= 2e-5, (k, hz.rel_err(out[k], g[k]))
    up = {k: v.numpy() for k, v in c["upstream"].items()}
    ga = eo.backward(c["allmap"].numpy(), cam.viewmatrix.numpy(), cam.fovx, cam.fovy, c["depth_ratio"], up)
    ref = g["g_allmap"]
    assert np.array_equal(np.isnan(ga), np.isnan(ref))           # alpha == 0 pixels: 0/0 in both
    if name == "mixed_with_holes":
        assert np.isnan(ref).any()
    ga, ref = np.nan_to_num(ga), np.nan_to_num(ref)
    assert hz.rel_err(ga, ref) <= 1e-4, hz.rel_err(ga, ref)
    for ch in range(7):   # per channel as well: the depth channels are much smaller than the others
        if np.abs(ref[ch]).max() > 0:
            assert hz.rel_err(ga[ch], ref[ch]) <= 2e-4, (ch, hz.rel_err(ga[ch], ref[ch]))
