import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz
import bench
from streetunveiler_b200 import synthetic as syn, _lib
dev = torch.device("cuda", 0)
wl = bench.Workload(2_000_000, 1, 1, 0, dev)
step, leaves, m2, state = bench.make_step(hz.ours_module(), wl)
def show(tag):
    s = step(); torch.cuda.synchronize()
    rad = s["radii"].cpu().numpy()
    print(tag, "R", s["R"], "vis", int((rad > 0).sum()), "sum radii", int(rad.sum()))
show("initial")
pin = {k: v.pin_memory() for k, v in wl.host.items() if isinstance(v, torch.Tensor)}
for k in leaves:
    before = leaves[k].detach().clone()
    leaves[k].data.copy_(pin[k], non_blocking=True)
    torch.cuda.synchronize()
    print("  copy", k, "changed elements:", int((before != leaves[k].detach()).sum()))
show("after pinned copy")
_lib.set_option("time_stages", 1)
show("timing on")
show("timing on 2")
print(_lib.stage_times())
_lib.set_option("time_stages", 0)
show("timing off")
