import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz
from streetunveiler_b200 import synthetic as syn

cam = syn.cam_a()
scene = syn.street_scene(2_000_000, 1, 3)
grads = syn.upstream_grads(cam.width, cam.height, "color_alpha")
r = hz.run_reference(scene, cam, grads=grads)
print("ref R", r["num_rendered"], "vis", int((r["radii"] > 0).sum()))
prev = None
for i in range(6):
    o = hz.run_ours(scene, cam, grads=grads if i % 2 == 0 else None)
    mism = np.nonzero(o["radii"] != r["radii"])[0]
    print(i, "ours R", o["num_rendered"], "vis", int((o["radii"] > 0).sum()), "radii mismatch", mism.size, mism[:10],
          "color equal", np.array_equal(o["color"], r["color"]))
    if mism.size:
        for j in mism[:5]:
            print("   idx", j, "ours", o["radii"][j], "ref", r["radii"][j], "xyz", scene["means3D"][j].tolist(), "scale", scene["scales"][j].tolist())
# now emulate bench: persistent leaves, repeated steps
import bench
wl = bench.Workload(2_000_000, 1, 1, 0, torch.device("cuda", 0))
step, leaves, m2, state = bench.make_step(hz.ours_module(), wl)
for i in range(6):
    s = step(); torch.cuda.synchronize()
    rad = s["radii"].cpu().numpy()
    print("bench-step", i, "R", s["R"], "vis", int((rad > 0).sum()), "radii mismatch vs ref", int((rad != r["radii"]).sum()))
