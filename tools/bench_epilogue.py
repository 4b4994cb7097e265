"""Fused render() epilogue vs the same op as plain PyTorch kernels, 1920x1280, fwd+bwd (GPU box)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from test_epilogue_gpu import torch_epilogue, view_of, KEYS
from epilogue_cases import synthetic_allmap
from streetunveiler_b200 import synthetic as syn
from streetunveiler_b200.surface_epilogue import render_epilogue

dev = torch.device("cuda")
cam = syn.cam_a(); H, W = cam.height, cam.width
allmap = synthetic_allmap(H, W, 7, holes=False).to(dev)
view = view_of(cam, dev)
g = torch.Generator().manual_seed(3)
up = {k: (torch.randn(n, H, W, generator=g) / (H * W)).to(dev) for k, n in
      [("rend_alpha", 1), ("rend_normal", 3), ("rend_dist", 1), ("surf_depth", 1), ("surf_normal", 3), ("surf_point", 3)]}

def step(fn):
    a = allmap.clone().requires_grad_(True)
    out = fn(a, view, 0.3)
    torch.autograd.backward([out[k] for k in KEYS], [up[k] for k in KEYS])
    return a.grad

def timeit(fn, n=30):
    for _ in range(5): step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(fn)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_fused, t_torch = timeit(render_epilogue), timeit(torch_epilogue)
HW = H * W
alg = (17 + 23) * 4 * HW      # fwd 17 planes, bwd 23 planes
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"op": "render() epilogue fwd+bwd 1920x1280 (incl. autograd glue and the clone)", "fused_ms": round(t_fused, 4),
                  "torch_ops_ms": round(t_torch, 4), "speedup": round(t_torch / t_fused, 2), "alg_bytes": alg,
                  "fused_gbs": round(alg / t_fused / 1e6, 1), "frac_of_hbm_peak": round(alg / t_fused / 1e6 / peak, 3)}))
