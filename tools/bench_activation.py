"""Fused GaussianModel activations + SH packing vs the torch ops of scene/gaussian_model.py:101-127, 2M Gaussians,
fwd+bwd (GPU box)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from streetunveiler_b200.parameter_activation import activate
from test_activation_gpu import torch_activate

dev = torch.device("cuda")
P = 2_000_000
g = torch.Generator().manual_seed(2)
raw = [torch.randn(P, 2, generator=g) - 3, torch.randn(P, 4, generator=g), torch.randn(P, 1, generator=g), torch.randn(P, 1, 3, generator=g),
       torch.randn(P, 15, 3, generator=g)]
raw = [t.to(dev).requires_grad_(True) for t in raw]
ups = [torch.randn(P, 2, generator=g), torch.randn(P, 4, generator=g), torch.randn(P, 1, generator=g), torch.randn(P, 16, 3, generator=g)]
ups = [t.to(dev) for t in ups]

def step(fn):
    for t in raw: t.grad = None
    torch.autograd.backward(list(fn(*raw)), ups)

def timeit(fn, n=20):
    for _ in range(3): step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(fn)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_fused, t_torch = timeit(activate), timeit(torch_activate)
alg = P * 4 * ((58 + 55) + (55 + 7 + 58))   # fwd: read 58 raw, write 7 + 48; bwd: read 55 upstream + 7 saved, write 58
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"op": "GaussianModel activations + SH packing fwd+bwd, 2M Gaussians (scene/gaussian_model.py:101-127)", "fused_ms": round(t_fused, 4),
                  "torch_ops_ms": round(t_torch, 4), "speedup": round(t_torch / t_fused, 2), "alg_bytes": alg,
                  "fused_gbs": round(alg / t_fused / 1e6, 1), "frac_of_hbm_peak": round(alg / t_fused / 1e6 / peak, 3)}))
