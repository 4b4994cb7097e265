"""GPU parity tests of the shared-binning colour passes: rasterize_color_passes and render_semantic
against the formulation the reference runs -- one complete rasterizer call per colour set -- with this repo's operator
and, where oracle/_ref exists, with the UNMODIFIED reference extension."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import harness as hz
from streetunveiler_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def pass_inputs(P, H, W, n, seed):
    g = torch.Generator().manual_seed(seed)
    colors = [torch.rand(P, 3, generator=g) for _ in range(n)]
    bgs = [torch.rand(3, generator=g) for _ in range(n)]
    up_c = [torch.randn(3, H, W, generator=g) / (3 * H * W) for _ in range(n)]
    up_a = torch.randn(7, H, W, generator=g) / (7 * H * W)
    return colors, bgs, up_c, up_a


def leaves_of(scene, dev, transmat=None):
    p = {k: scene[k].to(dev).clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations")}
    p["means2D"] = torch.zeros_like(p["means3D"], requires_grad=True)
    if transmat is not None:
        p["transMat"] = transmat.to(dev).clone().requires_grad_(True)
    return p


def geom_kw(p):
    if "transMat" in p:
        return dict(cov3D_precomp=p["transMat"])
    return dict(scales=p["scales"], rotations=p["rotations"])


def grads_of(p, colors):
    out = {k: v.grad.cpu().numpy() for k, v in p.items() if v.grad is not None}
    for i, c in enumerate(colors):
        if c.grad is not None:
            out[f"colors{i}"] = c.grad.cpu().numpy()
    return out


def run_separate(mod, scene, cam, colors, bgs, up_c, up_a, transmat=None, color_grads=True):
    """The reference's formulation: one GaussianRasterizer call per colour set; allmap taken from the first."""
    dev = torch.device("cuda")
    p = leaves_of(scene, dev, transmat)
    cs = [c.to(dev).clone().requires_grad_(color_grads) for c in colors]
    images, allmap, radii = [], None, None
    for i, c in enumerate(cs):
        rast = mod.GaussianRasterizer(hz._settings(mod, cam, bgs[i], 0, 1.0, dev))
        img, r, am = rast(means3D=p["means3D"], means2D=p["means2D"], opacities=p["opacities"], colors_precomp=c, **geom_kw(p))
        images.append(img)
        if i == 0:
            allmap, radii = am, r
    torch.autograd.backward(images + [allmap], [u.to(dev) for u in up_c] + [up_a.to(dev)])
    return [i.detach().cpu().numpy() for i in images], allmap.detach().cpu().numpy(), radii.cpu().numpy(), grads_of(p, cs)


def run_passes(scene, cam, colors, bgs, up_c, up_a, transmat=None, color_grads=True):
    from streetunveiler_b200.diff_surfel_rasterization.color_passes import rasterize_color_passes
    dev = torch.device("cuda")
    mod = hz.ours_module()
    p = leaves_of(scene, dev, transmat)
    cs = [c.to(dev).clone().requires_grad_(color_grads) for c in colors]
    images, radii, allmap = rasterize_color_passes(hz._settings(mod, cam, bgs[0], 0, 1.0, dev), p["means3D"], p["means2D"],
                                                   p["opacities"], cs, [b.to(dev) for b in bgs], **geom_kw(p))
    torch.autograd.backward(list(images) + [allmap], [u.to(dev) for u in up_c] + [up_a.to(dev)])
    return [i.detach().cpu().numpy() for i in images], allmap.detach().cpu().numpy(), radii.cpu().numpy(), grads_of(p, cs)


def assert_same(a, b, tol_img, tol_grad):
    (ia, aa, ra, ga), (ib, ab, rb, gb) = a, b
    assert np.array_equal(ra, rb)
    for x, y in zip(ia, ib):
        assert hz.rel_err(x, y) <= tol_img, hz.rel_err(x, y)
    assert hz.rel_err(aa, ab) <= tol_img
    assert set(ga) == set(gb), set(ga) ^ set(gb)
    for k in ga:
        assert hz.rel_err(ga[k], gb[k]) <= tol_grad, (k, hz.rel_err(ga[k], gb[k]))


@pytest.mark.parametrize("n", [1, 2, 3])
def test_color_passes_equal_separate_calls(n):
    cam = syn.cam_tilted(320, 208, 260.0)
    sc = syn.box_scene(20_000, 41, 0)
    args = pass_inputs(20_000, cam.height, cam.width, n, 5)
    a = run_passes(sc, cam, *args)
    b = run_separate(hz.ours_module(), sc, cam, *args)
    # same lists, same order, same per-pixel arithmetic: images bit-identical; gradients differ by float-add order only
    assert_same(a, b, 0.0, 2e-5)
    assert all(f"colors{i}" in a[3] for i in range(n))


def test_color_passes_without_color_gradients_and_precomputed_transmat():
    cam = syn.cam_tilted(200, 136, 170.0)
    sc = syn.box_scene(5_000, 43, 0)
    from cases import _transmat_like
    P = 5_000
    tm = _transmat_like(sc, cam)      # [P,9] splat->pixel matrices (the cov3D_precomp input of the operator)
    args = pass_inputs(P, cam.height, cam.width, 2, 6)
    a = run_passes(sc, cam, *args, transmat=tm, color_grads=False)
    b = run_separate(hz.ours_module(), sc, cam, *args, transmat=tm, color_grads=False)
    assert_same(a, b, 0.0, 2e-5)
    assert "transMat" in a[3] and "colors0" not in a[3]


@pytest.mark.skipif(not hz.reference_available(), reason="oracle/_ref (reference extension) not built")
def test_color_passes_against_reference_extension():
    cam = syn.cam_a()
    sc = syn.street_scene(200_000, 3, 0)
    args = pass_inputs(200_000, cam.height, cam.width, 2, 7)
    a = run_passes(sc, cam, *args)
    b = run_separate(hz.reference_module(), sc, cam, *args)
    assert_same(a, b, 1e-6, 1e-4)


def fake_model(P, dev, seed):
    sc = syn.box_scene(P, seed, 0)
    g = torch.Generator().manual_seed(seed)
    tags = torch.randint(0, 6, (P, 1), generator=g, dtype=torch.int32)
    p = {k: sc[k].to(dev).clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations")}
    pc = SimpleNamespace(get_xyz=p["means3D"], get_opacity=p["opacities"], get_scaling=p["scales"], get_rotation=p["rotations"],
                         get_semantics=tags.to(dev), get_semantics_32bit=(1 << tags.to(dev)), active_sh_degree=0)
    return pc, p


def reference_formulation_semantic(view, pc, filter_bit, reverse):
    """gaussian_renderer/__init__.py:371-446 written out with this repo's single-pass operator."""
    mod = hz.ours_module()
    dev = pc.get_xyz.device
    if filter_bit is None:
        m = slice(None)
    else:
        m = ((pc.get_semantics_32bit & filter_bit) != 0).reshape(-1)
        if reverse is False:
            m = ~m
    means3D, opacity, scales, rots, tag = pc.get_xyz[m], pc.get_opacity[m], pc.get_scaling[m], pc.get_rotation[m], pc.get_semantics[m]
    out = []
    bg_prob = [0., 0., 0., 0., 1., 0.]
    for i in range(0, 6, 3):
        semantic_3 = torch.zeros_like(means3D)
        for j in range(3):
            semantic_3[(tag == (i + j)).reshape(-1), j] = 1.0
        st = mod.GaussianRasterizationSettings(
            image_height=view.image_height, image_width=view.image_width, tanfovx=np.tan(view.FoVx * 0.5),
            tanfovy=np.tan(view.FoVy * 0.5), bg=torch.tensor(bg_prob[i:i + 3], device=dev), scale_modifier=1.0,
            viewmatrix=view.world_view_transform, projmatrix=view.full_proj_transform, sh_degree=0,
            campos=view.camera_center, prefiltered=False, debug=False)
        img, _, _ = mod.GaussianRasterizer(st)(means3D=means3D, means2D=torch.zeros_like(means3D), opacities=opacity,
                                               colors_precomp=semantic_3, scales=scales, rotations=rots)
        out.append(img)
    return torch.cat(out, 0)


@pytest.mark.parametrize("single_pass", [True, False])
@pytest.mark.parametrize("filter_bit,reverse", [(None, None), (1 << 2, True), (1 << 4, False)])
def test_render_semantic_matches_reference_formulation(filter_bit, reverse, single_pass):
    from streetunveiler_b200.semantic_passes import render_semantic
    from test_epilogue_gpu import view_of
    dev = torch.device("cuda")
    cam = syn.cam_tilted(256, 160, 210.0)
    view = view_of(cam, dev)
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False)
    gt = torch.softmax(torch.randn(6, cam.height, cam.width, generator=torch.Generator().manual_seed(1)), 0).to(dev)
    w = torch.tensor([1.0, 1.0, 1.0, 1.0, 0.2, 1.0], device=dev)

    def loss_of(sem):   # train.py:88
        return torch.nn.functional.cross_entropy(sem.unsqueeze(0), gt.unsqueeze(0), weight=w)

    pc, p = fake_model(15_000, dev, 51)
    pkg = render_semantic(view, pc, pipe, torch.zeros(3, device=dev), semantic_filter_bit=filter_bit, reverse_semantic=reverse,
                          single_pass=single_pass)
    assert pkg["render_semantics"].shape == (6, cam.height, cam.width)
    assert pkg["semantic_rgb"].shape == (3, cam.height, cam.width) and pkg["semantic_uncertainty"].shape == (cam.height, cam.width)
    loss_of(pkg["render_semantics"]).backward()
    pc2, p2 = fake_model(15_000, dev, 51)
    sem2 = reference_formulation_semantic(view, pc2, filter_bit, reverse)
    loss_of(sem2).backward()
    assert torch.equal(pkg["render_semantics"].detach(), sem2.detach())
    for k in p:
        assert hz.rel_err(p[k].grad.cpu().numpy(), p2[k].grad.cpu().numpy()) <= 2e-5, k


@pytest.mark.parametrize("n_classes", [1, 6, 8])
def test_class_probability_pass_equals_one_hot_colour_passes(n_classes):
    """One traversal with labels == one rasterizer call per three one-hot colours (images bit-identical), including a
    label outside the class range (contributes to nothing) and an arbitrary background vector."""
    from streetunveiler_b200.diff_surfel_rasterization.class_pass import rasterize_class_probabilities
    from streetunveiler_b200.semantic_passes import one_hot_colors
    dev = torch.device("cuda")
    mod = hz.ours_module()
    cam = syn.cam_tilted(320, 208, 260.0)
    P = 20_000
    sc = syn.box_scene(P, 61, 0)
    g = torch.Generator().manual_seed(n_classes)
    labels = torch.randint(-1, n_classes + 1, (P,), generator=g, dtype=torch.int32).to(dev)
    bg = torch.rand(n_classes, generator=g).to(dev)
    up = (torch.randn(n_classes, cam.height, cam.width, generator=g) / (cam.height * cam.width)).to(dev)

    p = leaves_of(sc, dev)
    probs, radii = rasterize_class_probabilities(hz._settings(mod, cam, torch.zeros(3), 0, 1.0, dev), p["means3D"], p["means2D"],
                                                 p["opacities"], labels, bg, scales=p["scales"], rotations=p["rotations"])
    probs.backward(up)

    q = leaves_of(sc, dev)
    imgs = []
    for i in range(0, n_classes, 3):
        k = min(3, n_classes - i)
        bg3 = torch.cat([bg[i:i + k], torch.zeros(3 - k, device=dev)])
        rast = mod.GaussianRasterizer(hz._settings(mod, cam, bg3.cpu(), 0, 1.0, dev))
        img, r2, _ = rast(means3D=q["means3D"], means2D=q["means2D"], opacities=q["opacities"],
                          colors_precomp=one_hot_colors(labels.reshape(-1, 1), i, n_classes), scales=q["scales"], rotations=q["rotations"])
        imgs.append(img[:k])
    ref = torch.cat(imgs, 0)
    ref.backward(up)
    assert torch.equal(radii, r2) and torch.equal(probs.detach(), ref.detach())
    for k in p:
        assert hz.rel_err(p[k].grad.cpu().numpy(), q[k].grad.cpu().numpy()) <= 2e-5, (k, hz.rel_err(p[k].grad.cpu().numpy(), q[k].grad.cpu().numpy()))


@pytest.mark.skipif(not hz.reference_available(), reason="oracle/_ref (reference extension) not built")
def test_class_probability_pass_against_reference_extension():
    from streetunveiler_b200.diff_surfel_rasterization.class_pass import rasterize_class_probabilities
    from streetunveiler_b200.semantic_passes import one_hot_colors
    dev = torch.device("cuda")
    cam = syn.cam_a()
    P = 200_000
    sc = syn.street_scene(P, 5, 0)
    g = torch.Generator().manual_seed(3)
    labels = torch.randint(0, 6, (P,), generator=g, dtype=torch.int32).to(dev)
    bg = torch.tensor([0., 0., 0., 0., 1., 0.], device=dev)
    up = (torch.randn(6, cam.height, cam.width, generator=g) / (cam.height * cam.width)).to(dev)
    p = leaves_of(sc, dev)
    probs, _ = rasterize_class_probabilities(hz._settings(hz.ours_module(), cam, torch.zeros(3), 0, 1.0, dev), p["means3D"],
                                             p["means2D"], p["opacities"], labels, bg, scales=p["scales"], rotations=p["rotations"])
    probs.backward(up)
    ref_mod = hz.reference_module()
    q = leaves_of(sc, dev)
    imgs = []
    for i in (0, 3):
        rast = ref_mod.GaussianRasterizer(hz._settings(ref_mod, cam, bg[i:i + 3].cpu(), 0, 1.0, dev))
        img, _, _ = rast(means3D=q["means3D"], means2D=q["means2D"], opacities=q["opacities"],
                         colors_precomp=one_hot_colors(labels.reshape(-1, 1), i, 6), scales=q["scales"], rotations=q["rotations"])
        imgs.append(img)
    ref = torch.cat(imgs, 0)
    ref.backward(up)
    assert hz.rel_err(probs.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= 1e-6
    for k in p:
        assert hz.rel_err(p[k].grad.cpu().numpy(), q[k].grad.cpu().numpy()) <= 1e-4, (k, hz.rel_err(p[k].grad.cpu().numpy(), q[k].grad.cpu().numpy()))
