"""CPU model of the blend kernels' sub-tile cull test (csrc/tile_pipeline.cuh: block_may_contribute, with the footprint
of csrc/preprocess.cu: cull_footprint restated in numpy) on the ORACLE's forward state.

It checks on the CPU what the GPU exactness tests check end to end: the test is conservative -- every (8x4 pixel block,
instance) pair in which some pixel passes the reference's alpha >= 1/255 test (forward.cu:386) is kept -- and it
records how many pairs it keeps per tile instance.  This model is how round 2 found that testing the block inflated by
half a pixel (round 1) instead of the rectangle spanned by the pixel CENTRES produced half of the non-contributing
evaluations (2M street scene: 1.63 -> 1.46 evaluated pairs per instance, 1.30 contributing)."""
import numpy as np
import torch

from streetunveiler_b200 import synthetic as syn


def _footprint(T, mx, my, op, W, H):
    """preprocess.cu cull_footprint in float64: 8-px bounding box units and the ellipse (e, M) of {rho3d <= tau}."""
    tau = 2.0 * np.log(np.maximum(255.0 * op, 1e-30)) * 1.002 + 0.01
    Tu, Tv, Tw = T[:, 0:3], T[:, 3:6], T[:, 6:9]
    rl = np.sqrt(0.5 * np.maximum(tau, 0)) + 0.5
    lx, hx, ly, hy = mx - rl, mx + rl, my - rl, my + rl
    tz2 = Tw[:, 2] ** 2
    dd = tau * (Tw[:, 0] ** 2 + Tw[:, 1] ** 2) - tz2
    ok = dd < -1e-3 * tz2
    inv = 1.0 / np.where(ok, dd, 1.0)
    f0, f2 = tau * inv, -inv
    cx = f0 * (Tu[:, 0] * Tw[:, 0] + Tu[:, 1] * Tw[:, 1]) + f2 * Tu[:, 2] * Tw[:, 2]
    cy = f0 * (Tv[:, 0] * Tw[:, 0] + Tv[:, 1] * Tw[:, 1]) + f2 * Tv[:, 2] * Tw[:, 2]
    qx = f0 * (Tu[:, 0] ** 2 + Tu[:, 1] ** 2) + f2 * Tu[:, 2] ** 2
    qy = f0 * (Tv[:, 0] ** 2 + Tv[:, 1] ** 2) + f2 * Tv[:, 2] ** 2
    ex = np.sqrt(np.maximum(cx * cx - qx, 0)) * 1.002 + 0.75
    ey = np.sqrt(np.maximum(cy * cy - qy, 0)) * 1.002 + 0.75
    lx = np.where(ok, np.minimum(lx, cx - ex), -1e9)
    hx = np.where(ok, np.maximum(hx, cx + ex), 1e9)
    ly = np.where(ok, np.minimum(ly, cy - ey), -1e9)
    hy = np.where(ok, np.maximum(hy, cy + ey), 1e9)
    box = (np.maximum(np.floor(lx), 0) // 8, np.minimum(np.ceil(hx), W - 1) // 8,
           np.maximum(np.floor(ly), 0) // 8, np.minimum(np.ceil(hy), H - 1) // 8)
    Tup, Tvp = Tu - mx[:, None] * Tw, Tv - my[:, None] * Tw
    a0, a1, a2 = np.cross(Tvp, Tw), np.cross(Tw, Tup), np.cross(Tup, Tvp)
    q = lambda a, b: a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] - tau * a[:, 2] * b[:, 2]   # noqa: E731
    q00, q01, q11, q02, q12, q22 = q(a0, a0), q(a0, a1), q(a1, a1), q(a0, a2), q(a1, a2), q(a2, a2)
    det = q00 * q11 - q01 * q01
    good = ok & (q00 > 0) & (q11 > 0) & (det > 1e-7 * q00 * q11)
    invd = 1.0 / np.where(good, det, 1.0)
    sx, sy = (q01 * q12 - q11 * q02) * invd, (q01 * q02 - q00 * q12) * invd
    fmin = q22 + q02 * sx + q12 * sy
    good &= fmin < 0
    sc = -1.0 / np.where(good, fmin, -1.0)
    return box, good, mx + sx, my + sy, q00 * sc, q01 * sc, q11 * sc


def _rect_min(ex, ey, m00, m01, m11, rx0, rx1, ry0, ry1):
    """tile_pipeline.cuh: minimum of the convex quadratic over the rectangle (0 if the centre is inside)."""
    X0, X1, Y0, Y1 = rx0 - ex, rx1 - ex, ry0 - ey, ry1 - ey
    inside = (X0 <= 0) & (X1 >= 0) & (Y0 <= 0) & (Y1 >= 0)
    ky, kx = -m01 / m11, -m01 / m00
    Ya, Yb = np.clip(ky * X0, Y0, Y1), np.clip(ky * X1, Y0, Y1)
    Xa, Xb = np.clip(kx * Y0, X0, X1), np.clip(kx * Y1, X0, X1)
    g = np.minimum(np.minimum(m00 * X0 * X0 + (2 * m01 * X0 + m11 * Ya) * Ya, m00 * X1 * X1 + (2 * m01 * X1 + m11 * Yb) * Yb),
                   np.minimum(m11 * Y0 * Y0 + (2 * m01 * Y0 + m00 * Xa) * Xa, m11 * Y1 * Y1 + (2 * m01 * Y1 + m00 * Xb) * Xb))
    return np.where(inside, 0.0, g)


def _pairs(fwd, W, H, inflate):
    """-> (contributing pairs, pairs kept by the cull test, missed pairs, tile instances) over every tile."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pl, ranges, tm, m2, no, ncon = fwd.point_list(), fwd.ranges(), fwd.transmat(), fwd.means2d(), fwd.normal_opacity(), fwd.n_contrib()
    true_n = kept_n = missed = inst = 0
    for t in range(gx * gy):
        a, b = int(ranges[t, 0]), int(ranges[t, 1])
        n = b - a
        if n == 0:
            continue
        ids = pl[a:b]
        ty, tx = divmod(t, gx)
        xs, ys = (tx * 16 + np.arange(16)).astype(np.float32), (ty * 16 + np.arange(16)).astype(np.float32)
        X, Y = np.meshgrid(xs, ys)
        X, Y = X[None], Y[None]
        T = tm[ids].astype(np.float32)
        Tu, Tv, Tw = T[:, 0:3], T[:, 3:6], T[:, 6:9]
        k = [X * Tw[:, i, None, None] - Tu[:, i, None, None] for i in range(3)]
        l = [Y * Tw[:, i, None, None] - Tv[:, i, None, None] for i in range(3)]
        px, py, pz = k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]
        with np.errstate(all="ignore"):
            sx, sy = px / pz, py / pz
            rho = np.minimum(sx * sx + sy * sy, 2 * ((m2[ids, 0, None, None] - X) ** 2 + (m2[ids, 1, None, None] - Y) ** 2))
            depth = sx * Tw[:, 0, None, None] + sy * Tw[:, 1, None, None] + Tw[:, 2, None, None]
            alpha = np.minimum(0.99, no[ids, 3, None, None] * np.exp(-0.5 * rho))
        contrib = (pz != 0) & ~(depth < 0.2) & ~(alpha < 1 / 255)
        contrib = np.pad(contrib, ((0, 0), (0, 0), (0, 0)))
        hh, ww = min(16, H - ty * 16), min(16, W - tx * 16)
        valid = np.zeros((16, 16), bool)
        valid[:hh, :ww] = True
        contrib &= valid[None]
        blocks = contrib.reshape(n, 4, 4, 2, 8).any(axis=(2, 4))            # [n, 4 block rows, 2 block columns]
        box, good, ex, ey, m00, m01, m11 = _footprint(T.astype(np.float64), m2[ids, 0].astype(np.float64),
                                                       m2[ids, 1].astype(np.float64), no[ids, 3].astype(np.float64), W, H)
        kept = np.zeros_like(blocks)
        mx, my = m2[ids, 0].astype(np.float64), m2[ids, 1].astype(np.float64)
        for yb in range(4):
            for xb in range(2):
                X0, Y0 = tx * 16 + xb * 8, ty * 16 + yb * 4
                inb = ((X0 >> 3) >= box[0]) & ((X0 >> 3) <= box[1]) & ((Y0 >> 3) >= box[2]) & ((Y0 >> 3) <= box[3])
                rx0, rx1, ry0, ry1 = X0 - inflate, X0 + 7 + inflate, Y0 - inflate, Y0 + 3 + inflate
                dx = np.maximum(np.maximum(rx0 - mx, mx - rx1), 0)
                dy = np.maximum(np.maximum(ry0 - my, my - ry1), 0)
                disc = dx * dx + dy * dy <= 5.6
                g = _rect_min(ex, ey, np.where(good, m00, 1.0), np.where(good, m01, 0.0), np.where(good, m11, 1.0), rx0, rx1, ry0, ry1)
                kept[:, yb, xb] = inb & (~good | disc | (g <= 1.02))
        true_n += int(blocks.sum())
        kept_n += int(kept.sum())
        missed += int((blocks & ~kept).sum())
        inst += n
    return true_n, kept_n, missed, inst


def test_cull_rectangle_of_pixel_centres_is_conservative_and_tighter():
    from oracle import oracle
    cam = syn.make_camera(320, 208, 342.5, 342.5)
    scene = syn.street_scene(12_000, 6, 0)
    scene["scales"] = (scene["scales"] * 3).contiguous()          # footprints of a few pixels, like the 2M scene at 1920x1280
    fwd = oracle.rasterize_forward(torch.zeros(3), scene["means3D"], None, scene["opacities"], scene["scales"],
                                   scene["rotations"], 1.0, None, cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy,
                                   cam.height, cam.width, scene["shs"], 0, cam.campos)
    assert fwd.num_rendered > 5_000
    true_n, hull_n, hull_missed, inst = _pairs(fwd, cam.width, cam.height, inflate=0.01)     # the kernel's rectangle
    _, infl_n, infl_missed, _ = _pairs(fwd, cam.width, cam.height, inflate=0.5)              # round 1's rectangle
    assert hull_missed == 0 and infl_missed == 0            # conservative: no contributing pair is culled
    assert true_n <= hull_n < infl_n                         # and strictly tighter than the inflated block
    assert (infl_n - hull_n) >= 0.25 * (infl_n - true_n)    # it removes a good part of the non-contributing evaluations
    print(f"pairs per tile instance: contributing {true_n / inst:.3f}, kept (centre hull) {hull_n / inst:.3f}, "
          f"kept (half-pixel inflated) {infl_n / inst:.3f}")
