"""GPU parity tests of the fused render() epilogue (csrc/epilogue.cu) through the C ABI."""
import math
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import harness as hz
from epilogue_cases import EPILOGUE_CASES, KEYS, build_epilogue_case, synthetic_allmap
from streetunveiler_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def view_of(cam, dev):
    return SimpleNamespace(image_width=cam.width, image_height=cam.height, FoVx=cam.fovx, FoVy=cam.fovy,
                           world_view_transform=cam.viewmatrix.to(dev), full_proj_transform=cam.projmatrix.to(dev),
                           camera_center=cam.campos.to(dev))


def torch_epilogue(allmap, view, depth_ratio):
    """Plain PyTorch fp32 version of the same op (the reference's formulation, written out with torch ops)."""
    W, H = view.image_width, view.image_height
    alpha = allmap[1:2]
    normal = (allmap[2:5].permute(1, 2, 0) @ view.world_view_transform[:3, :3].T).permute(2, 0, 1)
    med = torch.nan_to_num(allmap[5:6], 0, 0)
    exp = torch.nan_to_num(allmap[0:1] / alpha, 0, 0)
    sd = exp * (1 - depth_ratio) + depth_ratio * med
    c2w = view.world_view_transform.T.inverse()
    fx, fy = W / (2 * math.tan(view.FoVx / 2)), H / (2 * math.tan(view.FoVy / 2))
    K = torch.tensor([[fx, 0, W / 2], [0, fy, H / 2], [0, 0, 1.0]], device=allmap.device)
    gx, gy = torch.meshgrid(torch.arange(W, device=allmap.device), torch.arange(H, device=allmap.device), indexing="xy")
    pix = torch.stack([gx, gy, torch.ones_like(gx)], -1).reshape(-1, 3).float()
    rays = pix @ K.inverse().T @ c2w[:3, :3].T
    pts = (sd.reshape(-1, 1) * rays + c2w[:3, 3]).reshape(H, W, 3)
    out = torch.zeros_like(pts)
    dx = pts[2:, 1:-1] - pts[:-2, 1:-1]
    dy = pts[1:-1, 2:] - pts[1:-1, :-2]
    out[1:-1, 1:-1] = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return {"rend_alpha": alpha, "rend_normal": normal, "rend_dist": allmap[6:7], "surf_depth": sd,
            "surf_normal": out.permute(2, 0, 1) * alpha.detach(), "surf_point": pts.permute(2, 0, 1)}


def run_fused(allmap_cpu, cam, ratio, upstream):
    from streetunveiler_b200.surface_epilogue import render_epilogue
    dev = torch.device("cuda")
    a = allmap_cpu.to(dev).clone().requires_grad_(True)
    out = render_epilogue(a, view_of(cam, dev), ratio)
    loss = sum((out[k] * upstream[k].to(dev)).sum() for k in KEYS)
    loss.backward()
    res = {k: out[k].detach().cpu().numpy() for k in KEYS}
    res["g_allmap"] = a.grad.cpu().numpy()
    return res


def assert_close(res, ref, tol_f=2e-5, tol_g=1e-4, tol_normal=None):
    for k in KEYS:
        assert res[k].shape == ref[k].shape, k
        tol = tol_normal if (k == "surf_normal" and tol_normal) else tol_f
        assert hz.rel_err(res[k], ref[k]) <= tol, (k, hz.rel_err(res[k], ref[k]))
    ga, gr = res["g_allmap"], ref["g_allmap"]
    assert np.array_equal(np.isnan(ga), np.isnan(gr))
    ga, gr = np.nan_to_num(ga), np.nan_to_num(gr)
    for ch in range(7):
        scale = np.abs(gr[ch]).max()
        if scale > 0:
            assert np.abs(ga[ch] - gr[ch]).max() <= tol_g * scale, (ch, np.abs(ga[ch] - gr[ch]).max() / scale)
        else:
            assert not np.any(ga[ch]), ch


@pytest.mark.parametrize("name", EPILOGUE_CASES)
def test_fused_epilogue_matches_reference_golden(name):
    g = dict(np.load(os.path.join(GOLD, f"epilogue_{name}.npz")))
    c = build_epilogue_case(name)
    assert_close(run_fused(c["allmap"], c["cam"], c["depth_ratio"], c["upstream"]), g)


def test_fused_epilogue_full_size_against_oracle_and_torch():
    from oracle import epilogue_oracle as eo
    cam = syn.cam_a()
    H, W = cam.height, cam.width
    allmap = synthetic_allmap(H, W, 7, holes=True)
    g = torch.Generator().manual_seed(3)
    up = {k: (torch.randn(n, H, W, generator=g) / (H * W)) for k, n in
          [("rend_alpha", 1), ("rend_normal", 3), ("rend_dist", 1), ("surf_depth", 1), ("surf_normal", 3), ("surf_point", 3)]}
    res = run_fused(allmap, cam, 0.4, up)
    ref = eo.forward(allmap.numpy(), cam.viewmatrix.numpy(), cam.fovx, cam.fovy, 0.4)
    ref["g_allmap"] = eo.backward(allmap.numpy(), cam.viewmatrix.numpy(), cam.fovx, cam.fovy, 0.4,
                                  {k: v.numpy() for k, v in up.items()})
    # surf_normal differences neighbouring world points that are ~2.5 mm apart at 1920x1280 and ~5 m from the
    # origin: fp32 (the reference's precision too) leaves ~1e-4 relative in them against the fp64 oracle
    assert_close(res, ref, tol_f=5e-5, tol_g=1e-3, tol_normal=5e-4)
    # and against the same op written with plain torch ops on the GPU (fp32)
    dev = torch.device("cuda")
    a = allmap.to(dev).clone().requires_grad_(True)
    out = torch_epilogue(a, view_of(cam, dev), 0.4)
    sum((out[k] * up[k].to(dev)).sum() for k in KEYS).backward()
    tref = {k: out[k].detach().cpu().numpy() for k in KEYS}
    tref["g_allmap"] = a.grad.cpu().numpy()
    assert_close(res, tref, tol_f=5e-5, tol_g=2e-3, tol_normal=1e-3)


def test_render_frontend_end_to_end():
    """surface_epilogue.render == rasterizer + torch-op epilogue, including gradients into the Gaussians."""
    from streetunveiler_b200.surface_epilogue import render
    dev = torch.device("cuda")
    cam = syn.cam_tilted(320, 208, 260.0)
    sc = syn.box_scene(20_000, 31, 3)
    view = view_of(cam, dev)
    pipe = SimpleNamespace(debug=False, depth_ratio=0.25, convert_SHs_python=False, compute_cov3D_python=False)

    def run(fused):
        p = {k: v.to(dev).clone().requires_grad_(True) for k, v in sc.items() if isinstance(v, torch.Tensor)}
        pc = SimpleNamespace(get_xyz=p["means3D"], get_opacity=p["opacities"], get_scaling=p["scales"],
                             get_rotation=p["rotations"], get_features=p["shs"], active_sh_degree=3, max_sh_degree=3)
        if fused:
            rets = render(view, pc, pipe, torch.tensor([0.1, 0.2, 0.3], device=dev))
        else:
            mod = hz.ours_module()
            st = hz._settings(mod, cam, torch.tensor([0.1, 0.2, 0.3]), 3, 1.0, dev)
            color, radii, allmap = mod.GaussianRasterizer(st)(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"]),
                                                              opacities=p["opacities"], shs=p["shs"], scales=p["scales"],
                                                              rotations=p["rotations"])
            rets = {"render": color, "radii": radii}
            rets.update(torch_epilogue(allmap, view, pipe.depth_ratio))
        # a loss of the kind train.py:109-146 builds: colour + normal consistency + distortion
        normal_err = (1 - (rets["rend_normal"] * rets["surf_normal"]).sum(0)).mean()
        loss = rets["render"].mean() + 0.05 * normal_err + 10.0 * rets["rend_dist"].mean() + 0.01 * rets["surf_depth"].mean()
        loss.backward()
        return float(loss), {k: v.grad.cpu().numpy() for k, v in p.items()}, rets["radii"].cpu().numpy()

    lf, gf, rf = run(True)
    lt, gt, rt = run(False)
    assert abs(lf - lt) <= 1e-5 * abs(lt) and np.array_equal(rf, rt)
    for k in gf:
        assert hz.rel_err(gf[k], gt[k]) <= 2e-4, (k, hz.rel_err(gf[k], gt[k]))
