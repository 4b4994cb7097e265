"""Seeded inputs of the parameter-update golden cases (see make_golden_adam.py)."""
import torch

GROUPS = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]
SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (2,), "rotation": (4,)}
# arguments/__init__.py defaults: position_lr_init 1.6e-4 (x spatial_lr_scale), feature_lr 2.5e-3, /20 for f_rest,
# opacity_lr 0.05, scaling_lr 0.005, rotation_lr 0.001   (scene/gaussian_model.py:171-178)
LRS = {"xyz": 1.6e-4 * 5.3, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20.0, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
ADAM_CASES = {"p1000_steps3": (1000, 3, 21), "p37_steps5": (37, 5, 22)}    # name -> (P, steps, seed)


def build_adam_case(name):
    P, steps, seed = ADAM_CASES[name]
    g = torch.Generator().manual_seed(seed)
    params = {k: torch.randn((P,) + SHAPES[k], generator=g) for k in GROUPS}
    grads, radii, vgrads = [], [], []
    for s in range(steps):
        scale = 10.0 ** (-2 - s)              # several orders of magnitude: eps = 1e-15 must not matter, tiny v must
        gs = {k: scale * torch.randn((P,) + SHAPES[k], generator=g) for k in GROUPS}
        vis = torch.rand(P, generator=g) < 0.7
        for k in GROUPS:                      # Gaussians outside the frustum get exactly zero gradients
            gs[k][~vis] = 0.0
        grads.append(gs)
        r = torch.randint(1, 60, (P,), generator=g, dtype=torch.int32)
        r[~vis] = 0
        radii.append(r)
        vg = torch.randn(P, 3, generator=g) * 1e-3
        vg[:, 2] = 0.0
        vg[~vis] = 0.0
        vgrads.append(vg)
    return dict(P=P, steps=steps, params=params, grads=grads, radii=radii, vgrads=vgrads)
