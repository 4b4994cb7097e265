"""Edge cases of the "next" rows on the GPU: empty and fully culled scenes, tiny images, degenerate
optimiser groups, debug mode -- the situations the reference's callers can produce (densification can empty a semantic
subset; a camera can look away from everything)."""
import pytest
import torch
from torch import nn

import harness as hz
from streetunveiler_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _settings(cam, dev, debug=False):
    return hz._settings(hz.ours_module(), cam, torch.zeros(3), 0, 1.0, dev, debug=debug)


def test_semantic_ops_on_empty_and_culled_scenes():
    from streetunveiler_b200.diff_surfel_rasterization.class_pass import rasterize_class_probabilities
    from streetunveiler_b200.diff_surfel_rasterization.color_passes import rasterize_color_passes
    dev = torch.device("cuda")
    cam = syn.cam_s(70, 45, 60.0)
    st = _settings(cam, dev)
    bg6 = torch.tensor([0., 0., 0., 0., 1., 0.], device=dev)
    # P == 0: background only, empty gradients, no crash in backward
    e3 = torch.zeros(0, 3, device=dev, requires_grad=True)
    m2 = torch.zeros(0, 3, device=dev, requires_grad=True)
    kw = dict(scales=torch.zeros(0, 2, device=dev), rotations=torch.zeros(0, 4, device=dev))
    probs, radii = rasterize_class_probabilities(st, e3, m2, torch.zeros(0, 1, device=dev), torch.zeros(0, dtype=torch.int32, device=dev),
                                                 bg6, **kw)
    assert radii.numel() == 0 and probs.shape == (6, cam.height, cam.width)
    assert torch.equal(probs[4], torch.ones_like(probs[4])) and float(probs[:4].abs().max()) == 0
    probs.sum().backward()
    assert e3.grad.shape == (0, 3)
    imgs, radii, allmap = rasterize_color_passes(st, e3, m2, torch.zeros(0, 1, device=dev),
                                                 [torch.zeros(0, 3, device=dev)] * 2, [bg6[:3], bg6[3:]], **kw)
    assert float(imgs[0].abs().max()) == 0 and torch.equal(imgs[1][1], torch.ones_like(imgs[1][1])) and float(allmap.abs().max()) == 0
    # everything behind the camera
    sc = syn.box_scene(300, 1, 0)
    sc["means3D"][:, 2] = -sc["means3D"][:, 2]
    p = {k: sc[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations")}
    m2 = torch.zeros_like(p["means3D"], requires_grad=True)
    labels = torch.randint(0, 6, (300,), dtype=torch.int32).to(dev)
    probs, radii = rasterize_class_probabilities(st, p["means3D"], m2, p["opacities"], labels, bg6, scales=p["scales"],
                                                 rotations=p["rotations"])
    assert not (radii > 0).any() and torch.equal(probs[4], torch.ones_like(probs[4]))
    (probs * torch.randn_like(probs)).sum().backward()
    assert all(not torch.any(v.grad) for v in p.values())


def test_class_pass_debug_mode_and_argument_errors():
    from streetunveiler_b200.diff_surfel_rasterization.class_pass import rasterize_class_probabilities
    dev = torch.device("cuda")
    cam = syn.cam_tilted(96, 64, 90.0)
    sc = syn.box_scene(900, 13, 0)
    labels = torch.randint(0, 3, (900,), dtype=torch.int32).to(dev)
    bg = torch.tensor([0.2, 0.3, 0.5], device=dev)
    args = [sc["means3D"].to(dev), torch.zeros(900, 3, device=dev), sc["opacities"].to(dev), labels, bg]
    kw = dict(scales=sc["scales"].to(dev), rotations=sc["rotations"].to(dev))
    a, _ = rasterize_class_probabilities(_settings(cam, dev), *args, **kw)
    b, _ = rasterize_class_probabilities(_settings(cam, dev, debug=True), *args, **kw)    # per-stage sync: same result
    assert torch.equal(a, b)
    assert abs(float((a.sum(0) - 1).abs().max())) < 1e-5            # bg sums to 1: the class images partition unity
    with pytest.raises(RuntimeError):                                 # 9 classes in one pass
        rasterize_class_probabilities(_settings(cam, dev), *args[:4], torch.zeros(9, device=dev), **kw)
    with pytest.raises(RuntimeError):                                 # labels must be int32
        rasterize_class_probabilities(_settings(cam, dev), *args[:3], labels.long(), bg, **kw)
    with pytest.raises(Exception, match="scale/rotation pair"):
        rasterize_class_probabilities(_settings(cam, dev), *args)


def test_loss_block_on_tiny_images():
    """1x1 and 3x2 images: the whole image lies inside every window (all zero padding), against the torch formulation."""
    from test_loss_gpu import torch_training_loss, fused
    dev = torch.device("cuda")
    for H, W in [(1, 1), (3, 2), (33, 1)]:
        g = torch.Generator().manual_seed(H * 10 + W)
        mk = lambda c: torch.rand(c, H, W, generator=g).to(dev)
        vals, grads = [], []
        for fn in (fused, torch_training_loss):
            g.manual_seed(H * 10 + W)
            pkg = {k: mk(c).requires_grad_(True) for k, c in [("render", 3), ("rend_alpha", 1), ("rend_normal", 3), ("surf_normal", 3), ("rend_dist", 1)]}
            sky, gt = mk(3).requires_grad_(True), mk(3)
            loss, _ = fn(pkg, sky, gt, 0.2, 0.05, 10.0)
            loss.backward()
            vals.append(float(loss))
            grads.append({**{k: v.grad.cpu().numpy() for k, v in pkg.items()}, "sky": sky.grad.cpu().numpy()})
        assert abs(vals[0] - vals[1]) <= 1e-5 * max(1.0, abs(vals[1])), (H, W, vals)
        for k in grads[0]:
            assert hz.rel_err(grads[0][k], grads[1][k]) <= 1e-4, (H, W, k)


def test_fused_adam_degenerate_groups():
    from streetunveiler_b200.fused_adam import FusedAdam
    dev = torch.device("cuda")
    ps = [nn.Parameter(torch.randn(n, device=dev)) for n in (0, 1, 5, 4099)]
    frozen = nn.Parameter(torch.randn(7, device=dev))                  # never receives a gradient
    ref = [nn.Parameter(p.detach().clone()) for p in ps]
    a = FusedAdam([{"params": [p], "lr": 0.01 * (i + 1)} for i, p in enumerate(ps)] + [{"params": [frozen], "lr": 1.0}], eps=1e-15)
    b = torch.optim.Adam([{"params": [p], "lr": 0.01 * (i + 1)} for i, p in enumerate(ref)], eps=1e-15)
    before = frozen.detach().clone()
    for step in range(3):
        for p, r in zip(ps, ref):
            p.grad = torch.full_like(p, 0.1 * (step + 1))
            p.grad[::2] = 0.0                                          # exactly-zero gradients: the fast-path rescaling
            r.grad = p.grad.clone()
        a.step()
        b.step()
    for p, r in zip(ps, ref):
        assert hz.rel_err(p.detach().cpu().numpy(), r.detach().cpu().numpy()) <= 2e-6
    assert torch.equal(frozen, before) and len(a.state[frozen]) == 0
    # eleven groups: more than one launch (8 groups per call)
    many = [nn.Parameter(torch.randn(100, device=dev)) for _ in range(11)]
    many_ref = [nn.Parameter(p.detach().clone()) for p in many]
    a, b = FusedAdam(many, lr=1e-2), torch.optim.Adam(many_ref, lr=1e-2)
    for p, r in zip(many, many_ref):
        p.grad = torch.randn_like(p)
        r.grad = p.grad.clone()
    a.step()
    b.step()
    for p, r in zip(many, many_ref):
        assert hz.rel_err(p.detach().cpu().numpy(), r.detach().cpu().numpy()) <= 2e-6
