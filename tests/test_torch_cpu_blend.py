"""BASELINE configs[0]: the pure-PyTorch CPU alpha-blend (oracle/torch_cpu_blend.py, bench.py's `cpu_baseline_config1`)
pinned against the golden vectors captured from the unmodified reference extension on a B200
(tests/golden/config1_box10k.npz, made by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np

import harness as hz
from streetunveiler_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_box10k.npz")


def test_pure_pytorch_cpu_blend_matches_reference_golden():
    from oracle import torch_cpu_blend as tb
    g = dict(np.load(GOLD))
    cam, scene = syn.cam_s(), syn.box_scene(10_000, 3, 0)
    grads = syn.upstream_grads(cam.width, cam.height, "color_alpha")
    out, seconds = tb.forward_backward(scene, cam, grads)
    assert out["num_rendered"] == int(g["num_rendered"])                  # integer outputs of the binning: exact
    assert int((out["radii"] != g["radii"]).sum()) <= 2                   # ceil() of a radius within 1 ulp of an integer
    for k in ("color", "allmap"):
        scale = float(np.abs(g[k]).max())
        frac = float(np.mean(np.abs(out[k] - g[k]) > 1e-4 * scale))
        assert frac <= 1e-3, (k, frac)                                    # bulk within the 1e-4 bar ...
        assert hz.rel_err(out[k], g[k]) <= 2e-2, (k, hz.rel_err(out[k], g[k]))   # ... isolated alpha/T threshold flips
    # autograd gives the TRUE gradient; the reference's hand-written backward equals it for the colour and opacity
    # paths (its deviations -- SURVEY 8a quirks 1-3, 7 -- sit in the rotation / scale / position chain)
    assert hz.rel_err(out["g_shs"], g["g_shs"]) <= 1e-3
    assert hz.rel_err(out["g_opacities"], g["g_opacities"]) <= 2e-3
    assert seconds < 300
