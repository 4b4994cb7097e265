"""Parameter update on CPU: the numpy oracle and the CUDA kernels' bodies compiled for the host
(tests/emul/adam_emul.cu) against golden vectors produced by torch.optim.Adam itself over the reference's group list
(tests/golden/make_golden_adam.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as hz
from adam_cases import ADAM_CASES, GROUPS, LRS, build_adam_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(got, g, tol=2e-6):
    for k in GROUPS:
        for pre in ("p_", "m_", "v_"):
            assert got[pre + k].shape == g[pre + k].shape
            assert hz.rel_err(got[pre + k], g[pre + k]) <= tol, (pre + k, hz.rel_err(got[pre + k], g[pre + k]))
    assert np.array_equal(got["max_radii2D"], g["max_radii2D"])
    assert np.array_equal(got["denom"], g["denom"])
    assert hz.rel_err(got["xyz_gradient_accum"], g["xyz_gradient_accum"]) <= 1e-6


@pytest.mark.parametrize("name", list(ADAM_CASES))
def test_adam_oracle_matches_torch_adam(name):
    from oracle import adam_oracle as ao
    g = dict(np.load(os.path.join(GOLD, f"adam_{name}.npz")))
    c = build_adam_case(name)
    P = c["P"]
    p = {k: v.numpy().copy() for k, v in c["params"].items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    mr, acc, dn = np.zeros(P, np.float32), np.zeros((P, 1), np.float32), np.zeros((P, 1), np.float32)
    for s in range(c["steps"]):
        mr, acc, dn = ao.densification_stats(c["radii"][s].numpy(), c["vgrads"][s].numpy(), mr, acc, dn)
        for k in GROUPS:
            p[k], m[k], v[k] = ao.adam_step(p[k], c["grads"][s][k].numpy(), m[k], v[k], s + 1, LRS[k])
    got = {"max_radii2D": mr, "xyz_gradient_accum": acc, "denom": dn}
    for k in GROUPS:
        got["p_" + k], got["m_" + k], got["v_" + k] = p[k], m[k], v[k]
    check(got, g)


class Group(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("lr", C.c_double), ("step", C.c_int)]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", list(ADAM_CASES))
@pytest.mark.parametrize("misalign", [False, True])
def test_cuda_update_code_on_host_matches_torch_adam(emul, name, misalign):
    g = dict(np.load(os.path.join(GOLD, f"adam_{name}.npz")))
    c = build_adam_case(name)
    P = c["P"]

    def buf(a):   # optionally 4 bytes off a 16-byte boundary: exercises the scalar path of the kernel body
        a = np.asarray(a, np.float32)
        store = np.zeros(a.size + 8, np.float32)
        off = (-(store.ctypes.data // 4)) % 4 + (1 if misalign else 0)
        view = store[off:off + a.size].reshape(a.shape)
        view[...] = a
        assert (view.ctypes.data % 16 == 0) != misalign
        return view

    p = {k: buf(v.numpy()) for k, v in c["params"].items()}
    m = {k: buf(np.zeros_like(v)) for k, v in p.items()}
    v = {k: buf(np.zeros_like(x)) for k, x in p.items()}
    mr, acc, dn = np.zeros(P, np.float32), np.zeros((P, 1), np.float32), np.zeros((P, 1), np.float32)
    for s in range(c["steps"]):
        radii = np.ascontiguousarray(c["radii"][s].numpy())
        vg = np.ascontiguousarray(c["vgrads"][s].numpy())
        emul.emul_densification_stats(P, ptr(radii), ptr(vg), ptr(mr), ptr(acc), ptr(dn))
        grads = {k: buf(c["grads"][s][k].numpy()) for k in GROUPS}
        arr = (Group * len(GROUPS))(*[Group(ptr(p[k]), ptr(grads[k]), ptr(m[k]), ptr(v[k]), p[k].size, LRS[k], s + 1)
                                      for k in GROUPS])
        chunks = emul.emul_adam_step(len(GROUPS), arr, C.c_double(0.9), C.c_double(0.999), C.c_double(1e-15))
        assert chunks == sum((p[k].size + 4095) // 4096 for k in GROUPS)
    got = {"max_radii2D": mr, "xyz_gradient_accum": acc, "denom": dn}
    for k in GROUPS:
        got["p_" + k], got["m_" + k], got["v_" + k] = p[k], m[k], v[k]
    check(got, g)
