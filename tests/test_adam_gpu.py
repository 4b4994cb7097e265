"""GPU parity tests of the fused parameter update (csrc/adam.cu) through fused_adam.py / the C ABI."""
import os

import numpy as np
import pytest
import torch
from torch import nn

import harness as hz
from adam_cases import ADAM_CASES, GROUPS, LRS, SHAPES, build_adam_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_optimizer(cls, params):
    # scene/gaussian_model.py:171-180
    l = [{"params": [params[k]], "lr": LRS[k], "name": k} for k in GROUPS]
    return cls(l, lr=0.0, eps=1e-15)


def run_case(case, cls, stats):
    dev = torch.device("cuda")
    P = case["P"]
    params = {k: nn.Parameter(v.to(dev).clone()) for k, v in case["params"].items()}
    opt = make_optimizer(cls, params)
    mr, acc, dn = torch.zeros(P, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)
    for s in range(case["steps"]):
        for k in GROUPS:
            params[k].grad = case["grads"][s][k].to(dev)
        stats(case["radii"][s].to(dev), case["vgrads"][s].to(dev), mr, acc, dn)
        opt.step()
        opt.zero_grad(set_to_none=True)
    out = {"max_radii2D": mr.cpu().numpy(), "xyz_gradient_accum": acc.cpu().numpy(), "denom": dn.cpu().numpy()}
    for k in GROUPS:
        st = opt.state[params[k]]
        out["p_" + k], out["m_" + k], out["v_" + k] = (t.detach().cpu().numpy() for t in (params[k], st["exp_avg"], st["exp_avg_sq"]))
    return out, opt, params


def torch_stats(radii, vgrad, mr, acc, dn):
    vf = radii > 0
    mr[vf] = torch.max(mr[vf], radii[vf])                      # train.py:168
    acc[vf] += torch.norm(vgrad[vf], dim=-1, keepdim=True)     # scene/gaussian_model.py:556
    dn[vf] += 1


def fused_stats(radii, vgrad, mr, acc, dn):
    from streetunveiler_b200.fused_adam import densification_stats
    densification_stats(radii, vgrad, mr, acc, dn)


def check(got, ref, tol=2e-6):
    for k in GROUPS:
        for pre in ("p_", "m_", "v_"):
            assert hz.rel_err(got[pre + k], ref[pre + k]) <= tol, (pre + k, hz.rel_err(got[pre + k], ref[pre + k]))
    assert np.array_equal(got["max_radii2D"], ref["max_radii2D"]) and np.array_equal(got["denom"], ref["denom"])
    assert hz.rel_err(got["xyz_gradient_accum"], ref["xyz_gradient_accum"]) <= 1e-6


@pytest.mark.parametrize("name", list(ADAM_CASES))
def test_fused_update_matches_torch_adam_golden(name):
    from streetunveiler_b200.fused_adam import FusedAdam
    g = dict(np.load(os.path.join(GOLD, f"adam_{name}.npz")))
    got, _, _ = run_case(build_adam_case(name), FusedAdam, fused_stats)
    check(got, g)


def test_fused_update_large_against_torch_adam_on_gpu():
    """300k Gaussians (17.4M parameters, every chunk shape incl. tails), 4 steps, against torch.optim.Adam on the GPU."""
    from streetunveiler_b200.fused_adam import FusedAdam
    P, steps = 300_001, 4
    g = torch.Generator().manual_seed(9)
    case = dict(P=P, steps=steps, params={k: torch.randn((P,) + SHAPES[k], generator=g) for k in GROUPS}, grads=[], radii=[], vgrads=[])
    for s in range(steps):
        case["grads"].append({k: 0.01 * torch.randn((P,) + SHAPES[k], generator=g) for k in GROUPS})
        r = torch.randint(0, 40, (P,), generator=g, dtype=torch.int32)
        case["radii"].append(r)
        case["vgrads"].append(torch.randn(P, 3, generator=g) * 1e-3)
    got, _, _ = run_case(case, FusedAdam, fused_stats)
    ref, _, _ = run_case(case, torch.optim.Adam, torch_stats)
    check(got, ref)


def test_optimizer_surgery_of_the_reference_works_unchanged():
    """The statements of scene/gaussian_model.py:402-416 (_prune_optimizer) and :452-470 (cat_tensors_to_optimizer)
    run on FusedAdam exactly as on torch.optim.Adam, and state_dict round-trips between the two classes."""
    from streetunveiler_b200.fused_adam import FusedAdam
    case = build_adam_case("p37_steps5")
    dev = torch.device("cuda")

    def surgery_run(cls):
        _, opt, params = run_case(case, cls, torch_stats)
        mask = torch.arange(case["P"], device=dev) % 3 != 0
        for group in opt.param_groups:                              # _prune_optimizer
            stored_state = opt.state.get(group["params"][0], None)
            stored_state["exp_avg"] = stored_state["exp_avg"][mask]
            stored_state["exp_avg_sq"] = stored_state["exp_avg_sq"][mask]
            del opt.state[group["params"][0]]
            group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
            opt.state[group["params"][0]] = stored_state
        for group in opt.param_groups:                              # cat_tensors_to_optimizer
            ext = torch.full((5,) + tuple(group["params"][0].shape[1:]), 0.25, device=dev)
            stored_state = opt.state.get(group["params"][0], None)
            stored_state["exp_avg"] = torch.cat((stored_state["exp_avg"], torch.zeros_like(ext)), dim=0)
            stored_state["exp_avg_sq"] = torch.cat((stored_state["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
            del opt.state[group["params"][0]]
            group["params"][0] = nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
            opt.state[group["params"][0]] = stored_state
        for group in opt.param_groups:
            p = group["params"][0]
            p.grad = torch.linspace(-1, 1, p.numel(), device=dev).reshape(p.shape) * 1e-2
        opt.step()
        return opt

    a, b = surgery_run(FusedAdam), surgery_run(torch.optim.Adam)
    for ga, gb in zip(a.param_groups, b.param_groups):
        pa, pb = ga["params"][0], gb["params"][0]
        assert hz.rel_err(pa.detach().cpu().numpy(), pb.detach().cpu().numpy()) <= 2e-6
        assert float(a.state[pa]["step"]) == float(b.state[pb]["step"]) == case["steps"] + 1
    sd = b.state_dict()
    a.load_state_dict(sd)                                           # torch.optim.Adam checkpoint -> FusedAdam
    assert set(a.state_dict()["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_fused_update_errors():
    from streetunveiler_b200.fused_adam import FusedAdam, densification_stats
    p = nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p], lr=1e-3).step()                              # CPU parameter: no fallback
    with pytest.raises(RuntimeError):
        FusedAdam([p], lr=1e-3, amsgrad=True)
    dev = torch.device("cuda")
    with pytest.raises(RuntimeError):
        densification_stats(torch.zeros(4, device=dev), torch.zeros(4, 3, device=dev), torch.zeros(4, device=dev),
                            torch.zeros(4, 1, device=dev), torch.zeros(4, 1, device=dev))   # radii must be int32
