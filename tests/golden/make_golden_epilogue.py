"""Golden vectors for the render() epilogue, generated on CPU in this container by
running the reference's UNCHANGED ``gaussian_renderer.render`` (gaussian_renderer/__init__.py:18-188,
which calls utils/point_utils.py:8-37) with only the rasterizer stubbed out: the stub returns seeded
(color, radii, allmap) tensors, everything after line 145 is the reference's own code.

    python tests/golden/make_golden_epilogue.py        # needs /root/reference, no GPU

CUDA-only calls in that code (``.cuda()``, ``device="cuda"``) are redirected to the CPU for the duration.
Outputs: tests/golden/epilogue_*.npz (inputs are regenerated from seeds by tests/golden/epilogue_cases.py).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from epilogue_cases import EPILOGUE_CASES, build_epilogue_case  # noqa: E402

REF = "/root/reference"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def import_reference_renderer():
    for name in ["plyfile", "simple_knn", "simple_knn._C", "tinycudann", "sh_encoder", "sh_encoder._shencoder",
                 "scene.dataset_readers", "diff_surfel_rasterization"]:
        sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    import gaussian_renderer
    return gaussian_renderer


def run_reference_epilogue(gr, case):
    """-> (outputs dict, gradient of the upstream-weighted sum w.r.t. allmap)"""
    allmap = case["allmap"].clone().requires_grad_(True)
    cam, H, W = case["cam"], case["cam"].height, case["cam"].width

    class FakeRasterizer:  # stands in for GaussianRasterizer: returns the seeded tensors
        def __init__(self, raster_settings):
            pass

        def __call__(self, **kw):
            return torch.zeros(3, H, W), torch.ones(4, dtype=torch.int32), allmap

    view = types.SimpleNamespace(FoVx=cam.fovx, FoVy=cam.fovy, image_height=H, image_width=W,
                                 world_view_transform=cam.viewmatrix, full_proj_transform=cam.projmatrix,
                                 camera_center=cam.campos)
    pc = types.SimpleNamespace(get_xyz=torch.zeros(4, 3), get_opacity=torch.zeros(4, 1), get_scaling=torch.ones(4, 2),
                               get_rotation=torch.ones(4, 4), get_features=torch.zeros(4, 1, 3), active_sh_degree=0,
                               max_sh_degree=0)
    pipe = types.SimpleNamespace(convert_SHs_python=False, compute_cov3D_python=False, depth_ratio=case["depth_ratio"],
                                 debug=False)
    saved = (gr.GaussianRasterizer, gr.GaussianRasterizationSettings, torch.Tensor.cuda, torch.zeros_like)
    gr.GaussianRasterizer = FakeRasterizer
    gr.GaussianRasterizationSettings = lambda **kw: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    zl = torch.zeros_like
    torch.zeros_like = lambda *a, **k: zl(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    try:
        rets = gr.render(view, pc, pipe, torch.zeros(3))
    finally:
        gr.GaussianRasterizer, gr.GaussianRasterizationSettings, torch.Tensor.cuda, torch.zeros_like = saved
    keys = ["rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal", "surf_point"]
    loss = sum((rets[k] * case["upstream"][k]).sum() for k in keys)
    loss.backward()
    out = {k: rets[k].detach().numpy() for k in keys}
    out["g_allmap"] = allmap.grad.numpy()
    return out


def main():
    gr = import_reference_renderer()
    for name in EPILOGUE_CASES:
        case = build_epilogue_case(name)
        out = run_reference_epilogue(gr, case)
        np.savez_compressed(os.path.join(HERE, f"epilogue_{name}.npz"), **out)
        print(name, {k: tuple(v.shape) for k, v in out.items()}, "surf_normal |max|", float(np.abs(out["surf_normal"]).max()))


if __name__ == "__main__":
    main()
