"""One StreetUnveiler training iteration at benchmark size (2M surfels, 1920x1280, SH3), everything between the
parameters and the optimiser step (train.py:109-199 without data loading, the sky model and logging):

    activations (gaussian_model.py:107-127)  ->  rasterizer  ->  render() epilogue  ->  loss block  ->  backward
    ->  densification statistics  ->  Adam step

Arm "fused":      fused activations + this repo's rasterizer + fused epilogue (8f row 1) + fused loss (row 3) + fused
                  update (row 4)
Arm "reference":  the UNMODIFIED reference extension (oracle/_ref; falls back to this repo's rasterizer if it is not
                  built, and says so) + the same steps as the PyTorch ops the reference runs.
The "reference" and "mixed" arms use the torch ops of scene/gaussian_model.py:101-127 for the model's activations.
GPU box only."""
import json, os, sys
import torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import harness as hz
from adam_cases import LRS
from streetunveiler_b200 import synthetic as syn
from streetunveiler_b200.fused_adam import FusedAdam, densification_stats
from streetunveiler_b200.loss_block import training_loss
from streetunveiler_b200.parameter_activation import activate
from streetunveiler_b200.surface_epilogue import render_epilogue
from test_adam_gpu import torch_stats
from test_epilogue_gpu import torch_epilogue, view_of
from test_loss_gpu import torch_training_loss

dev = torch.device("cuda")
P = int(os.environ.get("ITER_P", 2_000_000))
cam = syn.cam_a(); H, W = cam.height, cam.width
view = view_of(cam, dev)
sc = syn.street_scene(P, 1, 3)
g = torch.Generator().manual_seed(8)
gt = torch.rand(3, H, W, generator=g).to(dev)
sky = torch.rand(3, H, W, generator=g).to(dev)
LAM = (0.2, 0.05, 100.0)


def make_model():
    inv_sig = lambda x: torch.log(x / (1 - x))
    raw = {"xyz": sc["means3D"], "f_dc": sc["shs"][:, :1].contiguous(), "f_rest": sc["shs"][:, 1:].contiguous(),
           "opacity": inv_sig(sc["opacities"].clamp(1e-4, 1 - 1e-4)), "scaling": torch.log(sc["scales"]), "rotation": sc["rotations"]}
    return {k: nn.Parameter(v.to(dev).clone()) for k, v in raw.items()}


def iteration(arm, params, opt, stats):
    mod = arm["mod"]
    scaling, rotation, opacity, features = arm["activate"](params["scaling"], params["rotation"], params["opacity"],
                                                            params["f_dc"], params["f_rest"])
    means2D = torch.zeros_like(params["xyz"], requires_grad=True) + 0
    means2D.retain_grad()
    st = hz._settings(mod, cam, torch.zeros(3), 3, 1.0, dev)
    color, radii, allmap = mod.GaussianRasterizer(st)(means3D=params["xyz"], means2D=means2D, opacities=opacity, shs=features,
                                                      scales=scaling, rotations=rotation)
    pkg = {"render": color}
    pkg.update(arm["epilogue"](allmap, view, 0.0))
    loss, _ = arm["loss"](pkg, sky, gt, *LAM)
    loss.backward()
    arm["stats"](radii, means2D.grad, *stats)
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


def bench(arm, n=10):
    params = make_model()
    opt = arm["opt"]([{"params": [params[k]], "lr": LRS[k], "name": k} for k in params], lr=0.0, eps=1e-15)
    stats = (torch.zeros(P, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev))
    for _ in range(3):
        iteration(arm, params, opt, stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = iteration(arm, params, opt, stats)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(loss)


def torch_activate(s, q, o, dc, rest):   # scene/gaussian_model.py:101-127
    return torch.exp(s), torch.nn.functional.normalize(q), torch.sigmoid(o), torch.cat((dc, rest), dim=1)


fused = dict(activate=activate, mod=hz.ours_module(), epilogue=render_epilogue, loss=training_loss, stats=densification_stats, opt=FusedAdam)
have_ref = hz.reference_available()
ref = dict(activate=torch_activate, mod=hz.reference_module() if have_ref else hz.ours_module(), epilogue=torch_epilogue, loss=torch_training_loss,
           stats=torch_stats, opt=torch.optim.Adam)
mixed = dict(activate=torch_activate, mod=hz.ours_module(), epilogue=torch_epilogue, loss=torch_training_loss, stats=torch_stats, opt=torch.optim.Adam)
t_f, l_f = bench(fused)
t_r, l_r = bench(ref)
t_m, l_m = bench(mixed)
print(json.dumps({"op": f"training iteration (activations, rasterizer, epilogue, loss, backward, statistics, Adam), {P} surfels, 1920x1280, SH3",
                  "fused_ms": round(t_f, 3), "reference_formulation_ms": round(t_r, 3),
                  "reference_rasterizer": "oracle/_ref extension" if have_ref else "NOT BUILT: this repo's rasterizer",
                  "our_rasterizer_with_torch_op_rows_ms": round(t_m, 3), "speedup": round(t_r / t_f, 2),
                  "iterations_per_s_fused": round(1e3 / t_f, 1), "loss_after_13_iterations": {"fused": l_f, "reference": l_r, "mixed": l_m}}))
