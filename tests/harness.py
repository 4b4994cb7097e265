"""Shared harness for parity tests, golden generation, smoke() and bench.py.

Three implementations behind one call shape (inputs: synthetic.scene dict + Camera):
  * ``run_ours``       streetunveiler_b200.diff_surfel_rasterization on cuda (the product)
  * ``run_reference``  the UNMODIFIED reference extension from oracle/_ref on cuda (checker / baseline)
  * ``run_oracle``     the CPU restatement oracle/ (checker)
Each returns a dict of numpy arrays: color, allmap, radii, num_rendered and (if grads were
requested) means3D, means2D, shs|colors, opacities, scales, rotations gradients.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from streetunveiler_b200 import synthetic  # noqa: E402

_REF_MOD = None


def reference_available() -> bool:
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "diff_surfel_rasterization", "_C.so"))


def reference_module():
    """The reference's own Python package + its compiled extension, imported under a private name."""
    global _REF_MOD
    if _REF_MOD is None:
        pkg = os.path.join(ROOT, "oracle", "_ref", "diff_surfel_rasterization")
        spec = importlib.util.spec_from_file_location("ref_diff_surfel_rasterization", os.path.join(pkg, "__init__.py"),
                                                      submodule_search_locations=[pkg])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_diff_surfel_rasterization"] = mod
        spec.loader.exec_module(mod)
        _REF_MOD = mod
    return _REF_MOD


def ours_module():
    from streetunveiler_b200 import diff_surfel_rasterization as mod
    return mod


def _settings(mod, cam, bg, sh_degree, scale_modifier, device, debug=False):
    return mod.GaussianRasterizationSettings(
        image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=bg.to(device), scale_modifier=scale_modifier, viewmatrix=cam.viewmatrix.to(device),
        projmatrix=cam.projmatrix.to(device), sh_degree=sh_degree, campos=cam.campos.to(device),
        prefiltered=False, debug=debug)


def run_torch_impl(mod, scene, cam, bg=None, grads=None, scale_modifier=1.0, use_precomp_color=False,
                   transmat_precomp=None, device="cuda", keep=False):
    """Run a torch-facing implementation (ours or the reference) forward (+ backward if ``grads``)."""
    bg = torch.zeros(3) if bg is None else bg
    dev = torch.device(device)
    need_grad = grads is not None
    p = {k: v.to(dev).clone().requires_grad_(need_grad) for k, v in scene.items() if isinstance(v, torch.Tensor)}
    means2D = torch.zeros_like(p["means3D"], requires_grad=need_grad)
    rast = mod.GaussianRasterizer(_settings(mod, cam, bg, int(scene["sh_degree"]), scale_modifier, dev))
    kw = {}
    if use_precomp_color:
        colors = p["shs"][:, 0, :].detach().abs().clone().requires_grad_(need_grad)
        kw["colors_precomp"] = colors
    else:
        kw["shs"] = p["shs"]
    if transmat_precomp is not None:
        tm = transmat_precomp.to(dev).clone().requires_grad_(need_grad)
        kw["cov3D_precomp"] = tm
    else:
        kw["scales"] = p["scales"]
        kw["rotations"] = p["rotations"]
    color, radii, allmap = rast(means3D=p["means3D"], means2D=means2D, opacities=p["opacities"], **kw)
    out = {"color": color.detach().cpu().numpy(), "allmap": allmap.detach().cpu().numpy(),
           "radii": radii.detach().cpu().numpy()}
    fn = color.grad_fn
    out["num_rendered"] = int(fn.num_rendered) if fn is not None and hasattr(fn, "num_rendered") else -1
    if need_grad:
        d_color, d_all = grads
        torch.autograd.backward([color, allmap], [d_color.to(dev), d_all.to(dev)])
        out["g_means3D"] = p["means3D"].grad.cpu().numpy()
        out["g_means2D"] = means2D.grad.cpu().numpy()
        out["g_opacities"] = p["opacities"].grad.cpu().numpy()
        if use_precomp_color:
            out["g_colors"] = kw["colors_precomp"].grad.cpu().numpy()
        else:
            out["g_shs"] = p["shs"].grad.cpu().numpy()
        if transmat_precomp is not None:
            out["g_transMat"] = kw["cov3D_precomp"].grad.cpu().numpy()
        else:
            out["g_scales"] = p["scales"].grad.cpu().numpy()
            out["g_rotations"] = p["rotations"].grad.cpu().numpy()
    if keep:
        out["_ctx"] = fn
    return out


def run_ours(scene, cam, **kw):
    return run_torch_impl(ours_module(), scene, cam, **kw)


def run_reference(scene, cam, **kw):
    return run_torch_impl(reference_module(), scene, cam, **kw)


def run_oracle(scene, cam, bg=None, grads=None, scale_modifier=1.0, use_precomp_color=False, transmat_precomp=None):
    from oracle import oracle
    bg = torch.zeros(3) if bg is None else bg
    colors = scene["shs"][:, 0, :].abs().contiguous() if use_precomp_color else None
    shs = None if use_precomp_color else scene["shs"]
    scales = None if transmat_precomp is not None else scene["scales"]
    rots = None if transmat_precomp is not None else scene["rotations"]
    f = oracle.rasterize_forward(bg, scene["means3D"], colors, scene["opacities"], scales, rots, scale_modifier,
                                 transmat_precomp, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy,
                                 cam.height, cam.width, shs, int(scene["sh_degree"]), cam.campos)
    out = {"color": f.color, "allmap": f.allmap, "radii": f.radii, "num_rendered": f.num_rendered, "_fwd": f}
    if grads is not None:
        g = oracle.rasterize_backward(f, grads[0], grads[1])
        out["g_means3D"] = g["means3D"]
        out["g_means2D"] = g["means2D"]
        out["g_opacities"] = g["opacities"]
        if use_precomp_color:
            out["g_colors"] = g["colors"]
        else:
            out["g_shs"] = g["sh"]
        if transmat_precomp is not None:
            out["g_transMat"] = g["transMat"]
        else:
            out["g_scales"] = g["scales"]
            out["g_rotations"] = g["rotations"]
    return out


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max|a-b| / max|b|  -- the parity criterion of SURVEY.md 8(d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    denom = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / (denom if denom > 0 else 1.0)


def compare(a: dict, b: dict, keys=None) -> dict:
    keys = keys or [k for k in a if not k.startswith("_") and k in b and k not in ("radii", "num_rendered")]
    return {k: rel_err(a[k], b[k]) for k in keys}
