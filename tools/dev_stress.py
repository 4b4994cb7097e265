import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz
import bench
from streetunveiler_b200 import synthetic as syn, _lib
from streetunveiler_b200.diff_surfel_rasterization import _C
dev = torch.device("cuda", 0)
wl = bench.Workload(2_000_000, 1, 1, 0, dev)
step, leaves, m2, state = bench.make_step(hz.ours_module(), wl)
_C.KEEP_LAST = True
def snap():
    geom = _C.LAST[1]
    t, i, o, r = _C.debug_geometry(wl.P, geom)
    return t.cpu(), i.cpu(), o.cpu(), state["radii"].cpu().clone()
s = step(); torch.cuda.synchronize(); R0 = s["R"]; base = snap()
print("base R", R0, "sum tiles", int(base[0].sum()), "perm ok", bool((torch.sort(base[1])[0] == torch.arange(wl.P)).all()))
pin = {k: v.pin_memory() for k, v in wl.host.items() if isinstance(v, torch.Tensor)}
out_host = {}
bad = 0
for it in range(150):
    if it % 3 == 0:
        for k in leaves:
            leaves[k].data.copy_(pin[k], non_blocking=True)
    s = step()
    if it % 3 == 1:
        outs = {"color": s["color"], "allmap": s["allmap"], "radii": s["radii"], "g_means2D": m2.grad}
        outs.update({"g_" + k: v.grad for k, v in leaves.items()})
        for k, v in outs.items():
            if k not in out_host:
                out_host[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
            out_host[k].copy_(v.detach(), non_blocking=True)
    if s["R"] != R0:
        bad += 1
        torch.cuda.synchronize()
        t, i, o, r = snap()
        print("ANOMALY it", it, "R", s["R"], "sum tiles", int(t.sum()), "tiles diff", int((t != base[0]).sum()),
              "idx diff", int((i != base[1]).sum()), "perm ok", bool((torch.sort(i)[0] == torch.arange(wl.P)).all()),
              "offsets last", int(o[-1]), "radii diff", int((r != base[3]).sum()))
        d = torch.nonzero(t != base[0]).flatten()[:8]
        for j in d.tolist():
            print("    g", j, "tiles", int(t[j]), "base", int(base[0][j]), "radii", int(r[j]), int(base[3][j]))
print("done, anomalies:", bad)
