"""GPU parity tests of the fused parameter activations (csrc/activate.cu) through parameter_activation.py."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from torch import nn

import harness as hz
from activation_cases import ACTIVATION_CASES, OUT, RAW, build_activation_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def torch_activate(s, q, o, dc, rest):
    """scene/gaussian_model.py:101-127 written out with the same torch ops."""
    return torch.exp(s), torch.nn.functional.normalize(q), torch.sigmoid(o), torch.cat((dc, rest), dim=1)


def run(raw, up, fn):
    dev = torch.device("cuda")
    leaves = {k: raw[k].to(dev).clone().requires_grad_(True) for k in RAW}
    outs = dict(zip(OUT, fn(*[leaves[k] for k in RAW])))
    sum((outs[k] * up[k].to(dev)).sum() for k in OUT).backward()
    return {k: v.detach().cpu().numpy() for k, v in outs.items()}, {k: v.grad.cpu().numpy() for k, v in leaves.items()}


@pytest.mark.parametrize("name", list(ACTIVATION_CASES))
def test_fused_activations_match_reference_model_golden(name):
    from streetunveiler_b200.parameter_activation import activate
    g = dict(np.load(os.path.join(GOLD, f"activation_{name}.npz")))
    c = build_activation_case(name)
    out, grads = run(c["raw"], c["upstream"], activate)
    for k in OUT:
        assert hz.rel_err(out[k], g[k]) <= 2e-6, (k, hz.rel_err(out[k], g[k]))
    assert np.array_equal(out["features"], g["features"])
    for k in RAW:
        assert hz.rel_err(grads[k], g["g_" + k]) <= 5e-6, (k, hz.rel_err(grads[k], g["g_" + k]))


def test_fused_activations_large_against_torch_ops():
    from streetunveiler_b200.parameter_activation import activate
    P = 300_001
    g = torch.Generator().manual_seed(4)
    raw = {"scaling_raw": torch.randn(P, 2, generator=g) - 3, "rotation_raw": torch.randn(P, 4, generator=g),
           "opacity_raw": torch.randn(P, 1, generator=g) * 3, "features_dc": torch.randn(P, 1, 3, generator=g),
           "features_rest": torch.randn(P, 15, 3, generator=g)}
    up = {"scaling": torch.randn(P, 2, generator=g), "rotation": torch.randn(P, 4, generator=g),
          "opacity": torch.randn(P, 1, generator=g), "features": torch.randn(P, 16, 3, generator=g)}
    a, ga = run(raw, up, activate)
    b, gb = run(raw, up, torch_activate)
    for k in OUT:
        assert hz.rel_err(a[k], b[k]) <= 2e-6, k
    assert np.array_equal(a["features"], b["features"]) and np.array_equal(ga["features_dc"], gb["features_dc"])
    assert np.array_equal(ga["features_rest"], gb["features_rest"])
    for k in RAW:
        assert hz.rel_err(ga[k], gb[k]) <= 5e-6, k


def test_activated_gaussians_wrapper_and_errors():
    from streetunveiler_b200.parameter_activation import ActivatedGaussians, activate
    dev = torch.device("cuda")
    c = build_activation_case("p36_sh1")
    raw = {k: nn.Parameter(v.to(dev)) for k, v in c["raw"].items()}
    pc = SimpleNamespace(_scaling=raw["scaling_raw"], _rotation=raw["rotation_raw"], _opacity=raw["opacity_raw"],
                         _features_dc=raw["features_dc"], _features_rest=raw["features_rest"], get_xyz=torch.zeros(36, 3, device=dev),
                         active_sh_degree=1)
    view = ActivatedGaussians(pc)
    assert view.get_xyz is pc.get_xyz and view.active_sh_degree == 1 and view.get_features.shape == (36, 4, 3)
    assert torch.allclose(view.get_rotation.norm(dim=1), torch.ones(36, device=dev), atol=1e-6)
    view.get_opacity.sum().backward()                     # only one output used: the others get zero gradients
    assert raw["opacity_raw"].grad is not None and not torch.any(raw["features_rest"].grad)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        activate(*[v.cpu() for v in c["raw"].values()])
    with pytest.raises(RuntimeError, match="shape"):
        activate(raw["scaling_raw"], raw["rotation_raw"][:, :3], raw["opacity_raw"], raw["features_dc"], raw["features_rest"])
