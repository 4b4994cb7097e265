"""Fused training-loss block vs the same ops as plain PyTorch kernels, 1920x1280, fwd+bwd (GPU box)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from test_loss_gpu import torch_training_loss, fused, full_size_case

dev = torch.device("cuda")
c = full_size_case()
pkg0 = {k: v.to(dev) for k, v in c["pkg"].items()}
sky0, gt = c["sky"].to(dev), c["gt"].to(dev)

def step(fn):
    pkg = {k: v.detach().requires_grad_(True) for k, v in pkg0.items()}
    sky = sky0.detach().requires_grad_(True)
    loss, _ = fn(pkg, sky, gt, 0.2, 0.05, 100.0)
    loss.backward()
    return loss

def timeit(fn, n=30):
    for _ in range(5): step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(fn)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_fused, t_torch = timeit(fused), timeit(torch_training_loss)
HW = 1280 * 1920
# compulsory traffic: fwd reads render3+alpha1+sky3+gt3 and writes 9 derivative planes, reads normals 6 + dist 1;
# bwd reads 9 + the same 10 image planes, writes d_render3+d_alpha1+d_sky3, reads normals 6, writes 6+1
alg = ((10 + 9 + 7) + (9 + 10 + 7 + 6 + 7)) * 4 * HW
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"op": "training loss block fwd+bwd 1920x1280 (train.py:113-136; incl. autograd glue)", "fused_ms": round(t_fused, 4),
                  "torch_ops_ms": round(t_torch, 4), "speedup": round(t_torch / t_fused, 2), "alg_bytes": alg,
                  "fused_gbs": round(alg / t_fused / 1e6, 1), "frac_of_hbm_peak": round(alg / t_fused / 1e6 / peak, 3)}))
