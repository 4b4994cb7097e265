"""Property tests (hypothesis) of the CUDA tile code compiled for the host (tests/emul/) against the numpy oracles on
random shapes: image sizes around the 32x32 tile and the 11-tap window, element counts around the 128-bit word and the
4096-element chunk, with and without the optional inputs.  CPU only; complements the fixed golden cases."""
import ctypes as C

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import harness as hz

COMMON = dict(deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@settings(max_examples=12, **COMMON)
@given(H=st.integers(1, 70), W=st.integers(1, 70), sky=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_loss_tile_code_matches_oracle_on_random_sizes(emul, H, W, sky, seed):
    from oracle import loss_oracle as lo
    rng = np.random.default_rng(seed)
    f = lambda c: np.ascontiguousarray(rng.random((c, H, W), dtype=np.float32))
    render, gt, alpha = f(3), f(3), f(1)
    sky_img = f(3) if sky else None
    a = alpha if sky else None
    deriv = np.full((9, H, W), np.nan, np.float32)
    means = np.zeros(2, np.float32)
    emul.emul_loss_photometric_forward(W, H, ptr(render), ptr(a), ptr(sky_img), ptr(gt), ptr(deriv), ptr(means))
    l1, ss = lo.photometric_forward(render, a, sky_img, gt)
    assert abs(means[0] - l1) <= 2e-6 * max(1.0, abs(l1)) and abs(means[1] - ss) <= 1e-5 * max(1.0, abs(ss))
    up = np.array([0.7, -0.3], np.float32)
    d_render = np.full((3, H, W), np.nan, np.float32)
    d_alpha = np.full((1, H, W), np.nan, np.float32) if sky else None
    d_sky = np.full((3, H, W), np.nan, np.float32) if sky else None
    emul.emul_loss_photometric_backward(W, H, ptr(render), ptr(a), ptr(sky_img), ptr(gt), ptr(deriv), ptr(up), ptr(d_render),
                                        ptr(d_alpha), ptr(d_sky))
    o_render, o_alpha, o_sky = lo.photometric_backward(render, a, sky_img, gt, 0.7, -0.3)
    assert np.isfinite(d_render).all() and hz.rel_err(d_render, o_render) <= 2e-4
    if sky:
        assert hz.rel_err(d_alpha, o_alpha) <= 2e-4 and hz.rel_err(d_sky, o_sky) <= 2e-4


class Group(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("lr", C.c_double), ("step", C.c_int)]


@settings(max_examples=15, **COMMON)
@given(sizes=st.lists(st.sampled_from([0, 1, 3, 4, 5, 4095, 4096, 4097, 9000]), min_size=1, max_size=8),
       step=st.integers(1, 2000), zero_frac=st.floats(0.0, 1.0), seed=st.integers(0, 2 ** 16))
def test_adam_code_matches_oracle_on_random_groups(emul, sizes, step, zero_frac, seed):
    from oracle import adam_oracle as ao
    rng = np.random.default_rng(seed)
    groups, keep, expect = [], [], []
    for i, n in enumerate(sizes):
        p = rng.standard_normal(n).astype(np.float32)
        g = (rng.standard_normal(n) * 10.0 ** rng.uniform(-20, 0, n)).astype(np.float32)
        g[rng.random(n) < zero_frac] = 0.0
        m = (rng.standard_normal(n) * 1e-3).astype(np.float32)
        v = (rng.random(n) * 10.0 ** rng.uniform(-40, -2, n)).astype(np.float32)     # down to the denormals and zero
        v[rng.random(n) < zero_frac] = 0.0
        lr = 10.0 ** rng.uniform(-5, -1)
        expect.append(ao.adam_step(p, g, m, v, step, lr))
        keep.append((p, g, m, v))
        groups.append(Group(ptr(p), ptr(g), ptr(m), ptr(v), n, lr, step))
    emul.emul_adam_step(len(groups), (Group * len(groups))(*groups), C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
    for (p, g, m, v), (ep, em, ev) in zip(keep, expect):
        if p.size == 0:
            continue
        assert np.isfinite(p).all()
        assert hz.rel_err(m, em) <= 2e-6 and hz.rel_err(v, ev) <= 2e-6 and hz.rel_err(p, ep) <= 2e-6


@settings(max_examples=15, **COMMON)
@given(P=st.integers(1, 700), R=st.sampled_from([0, 3, 8, 15]), seed=st.integers(0, 2 ** 16))
def test_activation_code_matches_oracle_on_random_sizes(emul, P, R, seed):
    from oracle import activation_oracle as ao
    rng = np.random.default_rng(seed)
    raw = [rng.standard_normal((P, 2)).astype(np.float32) - 3, rng.standard_normal((P, 4)).astype(np.float32),
           (rng.standard_normal((P, 1)) * 4).astype(np.float32), rng.standard_normal((P, 1, 3)).astype(np.float32),
           rng.standard_normal((P, R, 3)).astype(np.float32)]
    out = [np.full(s, np.nan, np.float32) for s in [(P, 2), (P, 4), (P, 1), (P, 1 + R, 3)]]
    emul.emul_activate_forward(P, R, *[ptr(a) for a in raw], *[ptr(a) for a in out])
    ref = ao.forward(*raw)
    for a, k in zip(out, ["scaling", "rotation", "opacity", "features"]):
        assert np.isfinite(a).all() and hz.rel_err(a, ref[k]) <= 2e-6, k
    up = {"scaling": rng.standard_normal((P, 2)).astype(np.float32), "rotation": rng.standard_normal((P, 4)).astype(np.float32),
          "opacity": rng.standard_normal((P, 1)).astype(np.float32), "features": rng.standard_normal((P, 1 + R, 3)).astype(np.float32)}
    grads = [np.full(a.shape, np.nan, np.float32) for a in raw]
    emul.emul_activate_backward(P, R, ptr(raw[1]), ptr(out[0]), ptr(out[2]), ptr(up["scaling"]), ptr(up["rotation"]),
                                ptr(up["opacity"]), ptr(up["features"]), *[ptr(a) for a in grads])
    refg = ao.backward(raw[0], raw[1], raw[2], up)
    for a, k in zip(grads, ["scaling_raw", "rotation_raw", "opacity_raw", "features_dc", "features_rest"]):
        if not a.size:
            continue
        assert np.isfinite(a).all(), k
        if k == "opacity_raw":
            # grad * (1 - y) * y with the fp32 y: one ulp of a saturated sigmoid (6e-8) is a large RELATIVE change of 1 - y,
            # in the kernel as in PyTorch -- bound the error by ulps of y, not by the size of the (tiny) result
            assert (np.abs(a - refg[k]) <= 5e-6 * np.abs(refg[k]) + 2.5e-7 * np.abs(up["opacity"])).all()
        else:
            assert hz.rel_err(a, refg[k]) <= 5e-6, k
