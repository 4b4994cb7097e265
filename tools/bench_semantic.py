"""render_semantic-shaped workload (2 colour passes, one-hot colours, cross-entropy-like upstream gradients) on 2M
surfels at 1920x1280: shared binning (rasterize_color_passes) vs one complete rasterizer call per pass (GPU box)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import harness as hz
from streetunveiler_b200 import synthetic as syn
from streetunveiler_b200.diff_surfel_rasterization.color_passes import rasterize_color_passes
from streetunveiler_b200.semantic_passes import one_hot_colors
from streetunveiler_b200.diff_surfel_rasterization.class_pass import rasterize_class_probabilities

dev = torch.device("cuda")
P = 2_000_000
cam = syn.cam_a(); H, W = cam.height, cam.width
sc = syn.street_scene(P, 1, 0)
g = torch.Generator().manual_seed(4)
tags = torch.randint(0, 6, (P, 1), generator=g, dtype=torch.int32).to(dev)
colors = [one_hot_colors(tags, i, 6) for i in (0, 3)]
bgs = [torch.tensor([0., 0., 0.], device=dev), torch.tensor([0., 1., 0.], device=dev)]
ups = [(torch.randn(3, H, W, generator=g) / (3 * H * W)).to(dev) for _ in range(2)]
leaves = {k: sc[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations")}
m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)

def impls():
    out = {"ours": hz.ours_module()}
    if hz.reference_available():
        out["reference_ext"] = hz.reference_module()
    return out

def step_separate(mod):
    def f():
        for t in list(leaves.values()) + [m2]: t.grad = None
        imgs = []
        for i in range(2):
            rast = mod.GaussianRasterizer(hz._settings(mod, cam, bgs[i].cpu(), 0, 1.0, dev))
            img, _, _ = rast(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"], colors_precomp=colors[i],
                             scales=leaves["scales"], rotations=leaves["rotations"])
            imgs.append(img)
        torch.autograd.backward(imgs, ups)
    return f

def step_shared():
    for t in list(leaves.values()) + [m2]: t.grad = None
    st = hz._settings(hz.ours_module(), cam, bgs[0].cpu(), 0, 1.0, dev)
    imgs, _, _ = rasterize_color_passes(st, leaves["means3D"], m2, leaves["opacities"], colors, bgs, scales=leaves["scales"],
                                        rotations=leaves["rotations"])
    torch.autograd.backward(imgs, ups)

up6 = torch.cat(ups, 0)
bg6 = torch.tensor([0., 0., 0., 0., 1., 0.], device=dev)
labels = tags.reshape(-1).contiguous()

def step_single():
    for t in list(leaves.values()) + [m2]: t.grad = None
    st = hz._settings(hz.ours_module(), cam, bgs[0].cpu(), 0, 1.0, dev)
    probs, _ = rasterize_class_probabilities(st, leaves["means3D"], m2, leaves["opacities"], labels, bg6, scales=leaves["scales"],
                                             rotations=leaves["rotations"])
    probs.backward(up6)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {"op": "semantic render: 6 class channels fwd+bwd, 2M surfels, 1920x1280", "single_class_pass_ms": round(timeit(step_single), 3),
       "shared_binning_ms": round(timeit(step_shared), 3)}
for name, mod in impls().items():
    res[f"separate_calls_{name}_ms"] = round(timeit(step_separate(mod)), 3)
res["single_pass_speedup_vs_separate_ours"] = round(res["separate_calls_ours_ms"] / res["single_class_pass_ms"], 2)
if "separate_calls_reference_ext_ms" in res:
    res["single_pass_speedup_vs_reference_ext"] = round(res["separate_calls_reference_ext_ms"] / res["single_class_pass_ms"], 2)
print(json.dumps(res))
