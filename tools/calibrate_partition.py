#!/usr/bin/env python
"""Per-range blend time of the screen partition on ONE GPU (virtual ranks, like tests/test_exchange_gpu.py): for a
scene and G ranges, time surfel_window_{prepare,render,backward} of every range alone on the GPU and print them next
to the range's instance count -- the data behind the partition's cost model (sharded.COST_BASE).

    python tools/calibrate_partition.py [P_total=16000000] [G=8] [cost_base ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz  # noqa: E402
from streetunveiler_b200 import synthetic as syn  # noqa: E402
from streetunveiler_b200.sharded import NativeBackend  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    bases = [int(x) for x in sys.argv[3:]] or [16]
    dev = torch.device("cuda")
    cam = syn.cam_a()
    be = NativeBackend()
    s = hz._settings(hz.ours_module(), cam, torch.zeros(3), 3, 1.0, dev)
    gc, ga = [g.to(dev) for g in syn.upstream_grads(cam.width, cam.height, "color_alpha")]
    pre = []
    for r in range(G):
        sc = syn.street_scene(P // G, 2000 + r, 3)
        p = {k: v.to(dev) for k, v in sc.items() if isinstance(v, torch.Tensor)}
        pre.append(be.shard_preprocess(s, p["means3D"], p["shs"], p["opacities"], p["scales"], p["rotations"]))
        del p, sc
    hist = sum(be.tile_hist(s, rec, radii).to(torch.int64) for radii, rec, keys, clamped in pre).to(torch.int32)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for base in bases:
        cuts, wr = be.partition(s, hist, G, base)
        cuts_h, wr_h = cuts.cpu().tolist(), wr.cpu().tolist()
        routes = [be.route_count(s, rec, radii, cuts, G) for radii, rec, keys, clamped in pre]
        cmat = torch.stack([c[:G] for _, c in routes]).cpu()
        print(f"cost_base {base}: cuts {cuts_h}")
        print(f"  rows sent per visible Gaussian: {int(cmat.sum()) / sum(int((p[0] > 0).sum()) for p in pre):.3f}")
        tot_f, tot_b = [], []
        for d in range(G):
            rows = []
            for r, (radii, rec, keys, clamped) in enumerate(pre):
                sr, _ = be.route_scatter(rec, radii, keys, routes[r][0], routes[r][1], int(cmat[r].sum()), G)
                a = int(cmat[r, :d].sum())
                rows.append(sr[a:a + int(cmat[r, d])].clone())
                del sr
            rows = torch.cat(rows, 0)
            rec_w, radii_w, keys_w = be.unpack(rows)
            tf, tb = [], []
            for it in range(4):
                e0, e1, e2 = ev(), ev(), ev()
                e0.record()
                pl, st = be.window_forward(s, rec_w, radii_w, keys_w, cuts_h[d], cuts_h[d + 1], int(wr_h[d]))
                e1.record()
                be.window_backward(s, rec_w, st, gc, ga)
                e2.record()
                torch.cuda.synchronize()
                tf.append(e0.elapsed_time(e1))
                tb.append(e1.elapsed_time(e2))
            tot_f.append(min(tf))
            tot_b.append(min(tb))
            print(f"  range {d}: tiles {cuts_h[d + 1] - cuts_h[d]:5d}  received {rows.shape[0]:8d}  instances {wr_h[d]:9d}  "
                  f"window fwd {min(tf):.3f} ms  bwd {min(tb):.3f} ms")
            del rows, rec_w, radii_w, keys_w, pl, st
        t = [a + b for a, b in zip(tot_f, tot_b)]
        print(f"  fwd+bwd window: max {max(t):.3f}  mean {sum(t) / G:.3f}  imbalance {max(t) / (sum(t) / G):.3f}")


if __name__ == "__main__":
    main()
