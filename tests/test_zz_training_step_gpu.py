"""The fused training iteration (streetunveiler_b200/training.py) against the same iteration written with this repo's
rasterizer and the PyTorch formulation of every other row (the composition tools/bench_iteration.py times).  Named to run
last: it only combines pieces that the other GPU tests pin individually."""
from types import SimpleNamespace

import pytest
import torch
from torch import nn

import harness as hz
from adam_cases import LRS
from streetunveiler_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
LAM = (0.2, 0.05, 10.0)


def make_model(sc, dev):
    inv_sig = lambda x: torch.log(x / (1 - x))
    P = sc["means3D"].shape[0]
    raw = {"_xyz": sc["means3D"], "_features_dc": sc["shs"][:, :1].contiguous(), "_features_rest": sc["shs"][:, 1:].contiguous(),
           "_opacity": inv_sig(sc["opacities"].clamp(1e-4, 1 - 1e-4)), "_scaling": torch.log(sc["scales"]), "_rotation": sc["rotations"]}
    m = SimpleNamespace(**{k: nn.Parameter(v.to(dev).clone()) for k, v in raw.items()})
    m.active_sh_degree = 3
    m.max_radii2D = torch.zeros(P, device=dev)
    m.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
    m.denom = torch.zeros(P, 1, device=dev)
    m.get_xyz = m._xyz
    return m


def groups(m):   # scene/gaussian_model.py:171-178
    return [{"params": [getattr(m, "_" + k)], "lr": LRS[n], "name": n} for k, n in
            [("xyz", "xyz"), ("features_dc", "f_dc"), ("features_rest", "f_rest"), ("opacity", "opacity"), ("scaling", "scaling"),
             ("rotation", "rotation")]]


def test_fused_training_step_tracks_the_torch_formulation():
    from streetunveiler_b200.fused_adam import FusedAdam
    from streetunveiler_b200.training import fused_training_step
    from test_adam_gpu import torch_stats
    from test_epilogue_gpu import torch_epilogue, view_of
    from test_loss_gpu import torch_training_loss
    dev = torch.device("cuda")
    cam = syn.cam_tilted(320, 208, 260.0)
    view = view_of(cam, dev)
    sc = syn.box_scene(20_000, 71, 3)
    g = torch.Generator().manual_seed(5)
    gt, sky = torch.rand(3, cam.height, cam.width, generator=g).to(dev), torch.rand(3, cam.height, cam.width, generator=g).to(dev)
    pipe = SimpleNamespace(debug=False, depth_ratio=0.0, convert_SHs_python=False, compute_cov3D_python=False)
    bg = torch.zeros(3, device=dev)

    a = make_model(sc, dev)
    opt_a = FusedAdam(groups(a), lr=0.0, eps=1e-15)
    b = make_model(sc, dev)
    opt_b = torch.optim.Adam(groups(b), lr=0.0, eps=1e-15)
    mod = hz.ours_module()
    losses_a, losses_b = [], []
    for it in range(3):
        loss, loss_dict, pkg = fused_training_step(a, view, pipe, bg, gt, sky, *LAM, optimizer=opt_a)
        losses_a.append(float(loss.detach()))
        assert set(loss_dict) == {"l1", "ssim", "Lnormal", "Ldist"} and pkg["render"].shape == (3, cam.height, cam.width)
        # the same iteration with torch ops around this repo's rasterizer (gaussian_model.py:101-127, train.py:109-199)
        scaling, rotation, opacity = torch.exp(b._scaling), torch.nn.functional.normalize(b._rotation), torch.sigmoid(b._opacity)
        features = torch.cat((b._features_dc, b._features_rest), dim=1)
        means2D = torch.zeros_like(b._xyz, requires_grad=True) + 0
        means2D.retain_grad()
        color, radii, allmap = mod.GaussianRasterizer(hz._settings(mod, cam, torch.zeros(3), 3, 1.0, dev))(
            means3D=b._xyz, means2D=means2D, opacities=opacity, shs=features, scales=scaling, rotations=rotation)
        ref_pkg = {"render": color}
        ref_pkg.update(torch_epilogue(allmap, view, 0.0))
        ref_loss, _ = torch_training_loss(ref_pkg, sky, gt, *LAM)
        ref_loss.backward()
        if it == 0:   # identical parameters on both sides (the fused activations may differ from torch's in the last ulp)
            assert int((pkg["radii"] != radii).sum()) <= 2
        torch_stats(radii, means2D.grad, b.max_radii2D, b.xyz_gradient_accum, b.denom)
        opt_b.step()
        opt_b.zero_grad(set_to_none=True)
        losses_b.append(float(ref_loss.detach()))
    for x, y in zip(losses_a, losses_b):
        assert abs(x - y) <= 1e-4 * abs(y), (losses_a, losses_b)
    # both arms saw (almost) the same set of visible Gaussians in every iteration
    da, db = float(a.denom.sum()), float(b.denom.sum())
    assert db > 0 and abs(da - db) <= 0.01 * db, (da, db)
