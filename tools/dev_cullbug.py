import os, sys, math
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as hz
from streetunveiler_b200 import synthetic as syn, _lib
from streetunveiler_b200.diff_surfel_rasterization import _C
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
g = torch.Generator().manual_seed(100 + seed)
P = 60_000
cam = syn.cam_tilted(640, 400, 520.0, yaw=0.3, pitch=-0.2)
sc = syn.box_scene(P, 40 + seed, 2)
z = torch.rand(P, generator=g)
sc["means3D"][:, 2] = torch.where(z < 0.3, 0.15 + z * 2.0, 1.0 + z * 12.0)
sc["means3D"][:, :2] *= 2.0
ratio = torch.exp(torch.randn(P, generator=g) * 2.0)
base = torch.exp(torch.randn(P, generator=g) * 1.2 + math.log(0.05))
sc["scales"] = torch.stack([base * ratio.sqrt(), base / ratio.sqrt()], 1).clamp(1e-4, 8.0).contiguous()
sc["opacities"] = torch.where(torch.rand(P, 1, generator=g) < 0.2, torch.rand(P, 1, generator=g) * 0.01, torch.rand(P, 1, generator=g)).contiguous()
bg = torch.tensor([0.2, 0.4, 0.6])
_C.KEEP_LAST = True
a = hz.run_ours(sc, cam, bg=bg)
R, geom, binb, img = _C.LAST
ranges, plist = _C.debug_binning(cam.width, cam.height, R, binb, img)
_, _, _, recs = _C.debug_geometry(P, geom)
ranges = ranges.cpu().numpy(); plist = plist.cpu().numpy(); recs = recs.cpu().numpy()
_lib.set_option("subtile_cull", 0)
b = hz.run_ours(sc, cam, bg=bg)
_lib.set_option("subtile_cull", 1)
diff = np.argwhere((a["color"] != b["color"]).any(0))
print("differing pixels:", len(diff), diff[:10].tolist())
gx = (cam.width + 15) // 16

def block_test(rec, rx0, rx1, ry0, ry1):
    mx, my = rec[9], rec[10]; ex, ey = rec[18], rec[19]; m00, m01, m11, r2 = rec[20:24]
    dx = max(rx0 - mx, mx - rx1, 0.0); dy = max(ry0 - my, my - ry1, 0.0)
    if dx * dx + dy * dy <= r2: return "disc"
    if m00 == 0: return "unbounded"
    X0, X1, Y0, Y1 = rx0 - ex, rx1 - ex, ry0 - ey, ry1 - ey
    if X0 <= 0 <= X1 and Y0 <= 0 <= Y1: return "center"
    ky, kx = -m01 / m11, -m01 / m00
    gs = []
    for Xe in (X0, X1):
        Y = min(max(ky * Xe, Y0), Y1); gs.append(m00 * Xe * Xe + (2 * m01 * Xe + m11 * Y) * Y)
    for Ye in (Y0, Y1):
        X = min(max(kx * Ye, X0), X1); gs.append(m11 * Ye * Ye + (2 * m01 * Ye + m00 * X) * X)
    return ("ellipse" if min(gs) <= 1.02 else None), min(gs)

seen = set()
for (py, px) in diff[:40]:
    t = (py // 16) * gx + px // 16
    r0, r1 = ranges[t]
    bx0 = (px // 8) * 8; by0 = (py // 4) * 4
    for i in range(r0, r1):
        gid = plist[i]; rec = recs[gid].astype(np.float64)
        Tu, Tv, Tw = rec[0:3], rec[3:6], rec[6:9]
        k = px * Tw - Tu; l = py * Tw - Tv; p = np.cross(k, l)
        if p[2] == 0: continue
        s = p[:2] / p[2]; rho3 = s @ s; d = rec[9:11] - np.array([px, py]); rho2 = 2 * (d @ d)
        depth = s[0] * Tw[0] + s[1] * Tw[1] + Tw[2]
        alpha = min(0.99, rec[11] * math.exp(-0.5 * min(rho3, rho2)))
        if depth < 0.2 or alpha < 1 / 255: continue
        res = block_test(rec, bx0 - 0.5, bx0 + 7.5, by0 - 0.5, by0 + 3.5)
        hit = res if isinstance(res, str) else res[0]
        if hit is None and gid not in seen:
            seen.add(gid)
            print(f"px ({px},{py}) gid {gid} alpha {alpha:.4f} rho3 {rho3:.3f} rho2 {rho2:.3f} depth {depth:.3f} gmin {res[1]:.4f}")
            print("   rec T:", rec[:9].tolist()); print("   mean", rec[9:11].tolist(), "opac", rec[11], "foot", rec[18:24].tolist())
            print("   xyz", sc["means3D"][gid].tolist(), "scale", sc["scales"][gid].tolist(), "rot", sc["rotations"][gid].tolist())
    if len(seen) >= 4: break
