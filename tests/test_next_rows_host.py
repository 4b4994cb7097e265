"""CPU checks of the host side of the "next" rows: argument validation, reference-compatible surfaces
and the no-fallback rule (CPU tensors raise; nothing silently computes on the host)."""
import inspect

import pytest
import torch
from torch import nn


def test_loss_block_surface_and_no_cpu_fallback():
    from streetunveiler_b200 import loss_block as lb
    # same names and argument lists as utils/loss_utils.py:17,33
    assert list(inspect.signature(lb.l1_loss).parameters) == ["network_output", "gt"]
    assert list(inspect.signature(lb.ssim).parameters) == ["img1", "img2", "window_size", "size_average"]
    a, b = torch.rand(3, 8, 8), torch.rand(3, 8, 8)
    for call in (lambda: lb.l1_loss(a, b), lambda: lb.ssim(a, b),
                 lambda: lb.training_loss({"render": a, "rend_alpha": a[:1], "rend_normal": a, "surf_normal": a, "rend_dist": a[:1]},
                                          None, b, 0.2)):
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            call()
    with pytest.raises(RuntimeError, match="window_size=11"):
        lb.ssim(a, b, window_size=7)


def test_fused_adam_is_a_torch_optimizer_with_adam_state_layout():
    from streetunveiler_b200.fused_adam import FusedAdam
    p = nn.Parameter(torch.zeros(4, 3))
    # the reference's construction (scene/gaussian_model.py:171-180): named groups with their own lr, lr=0.0, eps=1e-15
    opt = FusedAdam([{"params": [p], "lr": 1.6e-4, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert isinstance(opt, torch.optim.Optimizer)
    g = opt.param_groups[0]
    assert g["name"] == "xyz" and g["lr"] == 1.6e-4 and g["eps"] == 1e-15 and g["betas"] == (0.9, 0.999)
    ref = torch.optim.Adam([{"params": [nn.Parameter(torch.zeros(4, 3))], "lr": 1.6e-4, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert set(opt.state_dict()["param_groups"][0]) >= {"lr", "betas", "eps", "name", "params"}
    opt.load_state_dict(ref.state_dict())                      # a torch.optim.Adam checkpoint loads
    opt.step()                                                   # no gradients: nothing to do, nothing touched
    p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()                                               # CPU parameter: no fallback
    for bad in (dict(weight_decay=0.1), dict(amsgrad=True)):
        with pytest.raises(RuntimeError):
            FusedAdam([p], **bad)
    with pytest.raises(ValueError):
        FusedAdam([p], betas=(1.0, 0.999))


def test_densification_stats_validates_its_arguments():
    from streetunveiler_b200.fused_adam import densification_stats
    z = torch.zeros
    with pytest.raises(RuntimeError, match="radii"):
        densification_stats(z(4, dtype=torch.int32), z(4, 3), z(4), z(4, 1), z(4, 1))    # CPU tensors


def test_class_pass_validates_before_touching_the_device():
    from streetunveiler_b200.diff_surfel_rasterization import GaussianRasterizationSettings
    from streetunveiler_b200.diff_surfel_rasterization.class_pass import MAX_CLASSES, rasterize_class_probabilities
    s = GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    z = torch.zeros(4, 3)
    lab = torch.zeros(4, dtype=torch.int32)
    kw = dict(scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
    assert MAX_CLASSES == 8
    with pytest.raises(Exception, match="scale/rotation pair"):
        rasterize_class_probabilities(s, z, z, torch.zeros(4, 1), lab, torch.zeros(6))
    with pytest.raises(RuntimeError, match="classes per pass"):
        rasterize_class_probabilities(s, z, z, torch.zeros(4, 1), lab, torch.zeros(9), **kw)
    with pytest.raises(RuntimeError, match="labels"):
        rasterize_class_probabilities(s, z, z, torch.zeros(4, 1), lab, torch.zeros(6), **kw)   # CPU labels: no fallback


def test_training_step_surface():
    from streetunveiler_b200.training import fused_training_step
    params = list(inspect.signature(fused_training_step).parameters)
    assert params[:7] == ["gaussians", "viewpoint_cam", "pipe", "background", "gt_image", "sky_image", "lambda_dssim"]
    assert {"optimizer", "update_densification_stats", "lambda_normal", "lambda_dist"} <= set(params)
