"""Development script (GPU box): ours vs reference extension vs oracle, plus quick timings."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz  # noqa: E402
from streetunveiler_b200 import synthetic as syn  # noqa: E402
from streetunveiler_b200 import _lib  # noqa: E402


def report(tag, a, b):
    c = hz.compare(a, b)
    rad = int((a["radii"] != b["radii"]).sum())
    print(f"[{tag}] R {a['num_rendered']} vs {b['num_rendered']}  radii mismatches {rad}/{a['radii'].size}")
    print("    " + "  ".join(f"{k}={v:.2e}" for k, v in c.items()))
    return c


def small():
    for name, scene, cam, bgc in [
        ("box10k-sh0", syn.box_scene(10_000, 3, 0), syn.cam_s(), (0.0, 0.0, 0.0)),
        ("box2k-sh3-tilt", syn.box_scene(2_000, 7, 3), syn.cam_tilted(200, 136, 180.0), (0.3, 0.1, 0.7)),
    ]:
        bg = torch.tensor(bgc)
        grads = syn.upstream_grads(cam.width, cam.height, "all")
        o = hz.run_ours(scene, cam, bg=bg, grads=grads)
        r = hz.run_reference(scene, cam, bg=bg, grads=grads)
        c = hz.run_oracle(scene, cam, bg=bg, grads=grads)
        report(name + " ours-vs-ref", o, r)
        report(name + " oracle-vs-ref", c, r)
        r2 = hz.run_reference(scene, cam, bg=bg, grads=grads)
        report(name + " ref-vs-ref(noise)", r2, r)
        print("    fwd bit-equal ours/ref:", np.array_equal(o["color"], r["color"]), np.array_equal(o["allmap"], r["allmap"]))


def timing(mod, scene, cam, grads, iters=5):
    dev = torch.device("cuda")
    p = {k: v.to(dev).requires_grad_(True) for k, v in scene.items() if isinstance(v, torch.Tensor)}
    m2 = torch.zeros_like(p["means3D"], requires_grad=True)
    st = hz._settings(mod, cam, torch.zeros(3), int(scene["sh_degree"]), 1.0, dev)
    rast = mod.GaussianRasterizer(st)
    dc, da = grads[0].to(dev), grads[1].to(dev)
    tf, tb = [], []
    for i in range(iters + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        color, radii, allmap = rast(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                    scales=p["scales"], rotations=p["rotations"])
        e[1].record()
        torch.autograd.backward([color, allmap], [dc, da])
        e[2].record()
        torch.cuda.synchronize()
        if i >= 2:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[1].elapsed_time(e[2]))
        for t in list(p.values()) + [m2]:
            t.grad = None
    return float(np.median(tf)), float(np.median(tb))


def large(P, seed, mode):
    cam = syn.cam_a()
    scene = syn.street_scene(P, seed, 3)
    grads = syn.upstream_grads(cam.width, cam.height, mode)
    o = hz.run_ours(scene, cam, grads=grads)
    r = hz.run_reference(scene, cam, grads=grads)
    report(f"street{P}-{mode} ours-vs-ref", o, r)
    r2 = hz.run_reference(scene, cam, grads=grads)
    report(f"street{P}-{mode} ref-vs-ref(noise)", r2, r)
    print("    fwd bit-equal ours/ref:", np.array_equal(o["color"], r["color"]), np.array_equal(o["allmap"], r["allmap"]),
          " n diff px color:", int((o["color"] != r["color"]).sum()), " allmap:", int((o["allmap"] != r["allmap"]).sum()))
    _lib.lib().surfel_set_option(b"subtile_cull", 0)
    o2 = hz.run_ours(scene, cam, grads=grads)
    _lib.lib().surfel_set_option(b"subtile_cull", 1)
    print("    cull on/off fwd bit-equal:", np.array_equal(o["color"], o2["color"]), np.array_equal(o["allmap"], o2["allmap"]))
    report(f"street{P}-{mode} cull-on-vs-off", o, o2)
    del o, r, r2, o2
    t_ours = timing(hz.ours_module(), scene, cam, grads)
    t_ref = timing(hz.reference_module(), scene, cam, grads)
    print(f"    TIMING P={P}: ours fwd {t_ours[0]:.3f} ms bwd {t_ours[1]:.3f} ms | ref fwd {t_ref[0]:.3f} ms bwd {t_ref[1]:.3f} ms"
          f" | speedup {(t_ref[0]+t_ref[1])/(t_ours[0]+t_ours[1]):.2f}x")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    small()
    large(500_000, 0, "color_alpha")
    large(2_000_000, 1, "all")
