"""Golden vectors for the per-iteration parameter update, generated on CPU in this container by
``torch.optim.Adam`` itself (PyTorch 2.11 -- the third-party implementation the reference calls) over the group list of
scene/gaussian_model.py:171-180, with the statistics statements of train.py:168 / gaussian_model.py:555-557 executed
verbatim on CPU tensors.

    python tests/golden/make_golden_adam.py

Outputs: tests/golden/adam_*.npz (inputs are regenerated from seeds by tests/golden/adam_cases.py).
"""
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from adam_cases import ADAM_CASES, GROUPS, LRS, build_adam_case  # noqa: E402


def run_reference_update(case):
    P = case["P"]
    params = {k: nn.Parameter(v.clone()) for k, v in case["params"].items()}
    # scene/gaussian_model.py:171-180
    l = [{"params": [params[k]], "lr": LRS[k], "name": k} for k in GROUPS]
    optimizer = torch.optim.Adam(l, lr=0.0, eps=1e-15)
    max_radii2D = torch.zeros(P)
    xyz_gradient_accum = torch.zeros(P, 1)
    denom = torch.zeros(P, 1)
    for s in range(case["steps"]):
        for k in GROUPS:
            params[k].grad = case["grads"][s][k].clone()
        radii, vgrad = case["radii"][s], case["vgrads"][s]
        visibility_filter = radii > 0
        # train.py:168
        max_radii2D[visibility_filter] = torch.max(max_radii2D[visibility_filter], radii[visibility_filter])
        # scene/gaussian_model.py:555-557
        xyz_gradient_accum[visibility_filter] += torch.norm(vgrad[visibility_filter], dim=-1, keepdim=True)
        denom[visibility_filter] += 1
        optimizer.step()                       # train.py:197
        optimizer.zero_grad(set_to_none=True)
    out = {"max_radii2D": max_radii2D.numpy(), "xyz_gradient_accum": xyz_gradient_accum.numpy(), "denom": denom.numpy()}
    for k in GROUPS:
        st = optimizer.state[params[k]]
        out["p_" + k] = params[k].detach().numpy()
        out["m_" + k] = st["exp_avg"].numpy()
        out["v_" + k] = st["exp_avg_sq"].numpy()
    return out


if __name__ == "__main__":
    for name in ADAM_CASES:
        out = run_reference_update(build_adam_case(name))
        np.savez_compressed(os.path.join(HERE, f"adam_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("p_")})
