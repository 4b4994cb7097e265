"""Backward-blend kernel variants (surfel_set_option "bwd_variant": 0 = one kernel, 96 registers; 1 = device-side choice
between the colour+alpha-only kernel at 72 registers and the full one at 96; 2 = both at 72 (default); 3 = 56 | 96;
4 = 92 | 80) at benchmark size: stage time of render_bwd for
colour+alpha gradients (BASELINE configs 2/3/5) and for all ten gradient planes (config 4).  GPU box."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz
from streetunveiler_b200 import _lib, synthetic as syn

dev = torch.device("cuda")
cam = syn.cam_a()
sc = syn.street_scene(2_000_000, 1, 3)
mod = hz.ours_module()
p = {k: sc[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2 = torch.zeros_like(p["means3D"], requires_grad=True)
rast = mod.GaussianRasterizer(hz._settings(mod, cam, torch.zeros(3), 3, 1.0, dev))
res = {}
for mode in ("color_alpha", "all"):
    gc, ga = (g.to(dev) for g in syn.upstream_grads(cam.width, cam.height, mode))
    for variant in range(5):
        _lib.set_option("bwd_variant", variant)
        def step():
            for t in list(p.values()) + [m2]: t.grad = None
            c, r, a = rast(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"], scales=p["scales"], rotations=p["rotations"])
            torch.autograd.backward([c, a], [gc, ga])
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): step()
        e1.record(); torch.cuda.synchronize()
        total = e0.elapsed_time(e1) / 10
        _lib.set_option("time_stages", 1)
        for _ in range(5): step()
        st = _lib.stage_times()
        _lib.set_option("time_stages", 0)
        res[f"{mode}/v{variant}"] = {"step_ms": round(total, 3), "render_bwd_ms": round(st["render_bwd"][0] / st["render_bwd"][1], 3)}
_lib.set_option("bwd_variant", 0)
print(json.dumps(res))
