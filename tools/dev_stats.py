"""Work statistics of the tile blend on the benchmark scene (dev only)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz
from streetunveiler_b200 import synthetic as syn
from streetunveiler_b200.diff_surfel_rasterization import _C
_C.KEEP_LAST = True
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cam = syn.cam_a(); scene = syn.street_scene(P, 1, 3)
o = hz.run_ours(scene, cam)
R, geom, binb, img = _C.LAST
ranges, plist = _C.debug_binning(cam.width, cam.height, R, binb, img)
tiles_t, idx_s, offs, recs = _C.debug_geometry(P, geom)
ranges = ranges.cpu().numpy(); plist = plist.cpu(); recs = recs.cpu()
n_per_tile = ranges[:, 1] - ranges[:, 0]
print("R", R, "tiles", len(n_per_tile), "instances/tile mean %.1f median %.0f p90 %.0f max %d" % (n_per_tile.mean(), np.median(n_per_tile), np.quantile(n_per_tile, .9), n_per_tile.max()))
rad = o["radii"]; vis = rad > 0
print("radius: mean %.1f median %.0f p90 %.0f" % (rad[vis].mean(), np.median(rad[vis]), np.quantile(rad[vis], .9)))
gx = (cam.width + 15) // 16
rng = np.random.default_rng(0)
sample = rng.choice(np.nonzero(n_per_tile > 0)[0], 400, replace=False)
tot = dict(inst=0, pair_bbox=0, pair_contrib=0, lane_contrib=0, lane_inbbox=0, iters_now=0, iters_H=0, iters_H4=0, chunks=0, visited=0)
for t in sample:
    a, b = ranges[t]
    rec = recs[plist[a:b]].numpy()  # [n,20]
    n = rec.shape[0]
    tx, ty = (t % gx) * 16, (t // gx) * 16
    px = (tx + np.arange(16))[None, :].repeat(16, 0).astype(np.float32)
    py = (ty + np.arange(16))[:, None].repeat(16, 1).astype(np.float32)
    Tu, Tv, Tw = rec[:, 0:3], rec[:, 3:6], rec[:, 6:9]
    k = px[None, :, :, None] * Tw[:, None, None, :] - Tu[:, None, None, :]
    l = py[None, :, :, None] * Tw[:, None, None, :] - Tv[:, None, None, :]
    p = np.cross(k, l)
    with np.errstate(all="ignore"):
        sx, sy = p[..., 0] / p[..., 2], p[..., 1] / p[..., 2]
        rho3 = sx * sx + sy * sy
        dx, dy = rec[:, 9, None, None] - px[None], rec[:, 10, None, None] - py[None]
        rho2 = 2 * (dx * dx + dy * dy)
        rho = np.minimum(rho3, rho2)
        depth = sx * Tw[:, None, None, 0] + sy * Tw[:, None, None, 1] + Tw[:, None, None, 2]
        alpha = np.minimum(0.99, rec[:, 11, None, None] * np.exp(-0.5 * rho))
    ok = (p[..., 2] != 0) & (depth >= 0.2) & (alpha >= 1 / 255)
    alpha_eff = np.where(ok, alpha, 0.0)
    Tcum = np.cumprod(1 - alpha_eff, axis=0)
    Tbefore = np.concatenate([np.ones((1, 16, 16)), Tcum[:-1]], 0)
    alive = (Tbefore * (1 - alpha_eff)) >= 1e-4
    alive = np.logical_and.accumulate(alive | ~ok, axis=0)  # done once a test fails
    contrib = ok & alive                                   # [n,16,16]
    visited = alive.any(axis=(1, 2))
    bb = rec[:, 18:20].copy().view(np.uint32)
    x0, x1, y0, y1 = bb[:, 0] & 0xffff, bb[:, 0] >> 16, bb[:, 1] & 0xffff, bb[:, 1] >> 16
    inb = (px[None] >= x0[:, None, None]) & (px[None] <= x1[:, None, None]) & (py[None] >= y0[:, None, None]) & (py[None] <= y1[:, None, None])
    assert not (contrib & ~inb).any(), "bbox not conservative!"
    # 8x4 blocks: [n, 4(yb), 4, 2(xb), 8]
    def blocks(m, bh, bw):
        return m.reshape(n, 16 // bh, bh, 16 // bw, bw).transpose(0, 1, 3, 2, 4).reshape(n, (16 // bh) * (16 // bw), bh * bw)
    inb_b = blocks(inb & alive, 4, 8); con_b = blocks(contrib, 4, 8)
    tot["inst"] += n; tot["visited"] += int(visited.sum())
    tot["pair_bbox"] += int(inb_b.any(2).sum()); tot["pair_contrib"] += int(con_b.any(2).sum())
    tot["lane_contrib"] += int(contrib.sum()); tot["lane_inbbox"] += int((inb & alive).sum())
    tot["iters_now"] += int(inb_b.any(2).sum())
    for c in range(0, n, 32):
        ch = inb_b[c:c + 32]                       # [<=32, 8 blocks, 32 px]
        tot["iters_H"] += int(ch.sum(0).max(1).sum())   # per block: max over pixels of #candidates
        tot["chunks"] += 8
for k, v in tot.items():
    print(k, v, "per instance %.2f" % (v / tot["inst"]))
print("lane efficiency now: %.3f   with per-lane walking (H): %.3f" % (tot["lane_contrib"] / (32 * tot["iters_now"]), tot["lane_contrib"] / (32 * tot["iters_H"])))
