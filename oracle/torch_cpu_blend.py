"""Pure-PyTorch CPU alpha-blend of 2D-Gaussian surfels (BASELINE.json configs[0]: "10k random surfels, 256x256,
SH deg 0, 1 cam -- pure-PyTorch CPU alpha-blend reference (no GPU)").

TEST / MEASUREMENT INFRASTRUCTURE ONLY: may be imported from tests/ and from bench.py's cpu-baseline legs, never from
streetunveiler_b200/.  It is a restatement of the reference's FORWARD algorithm with dense torch ops on the host
(every statement cites the reference line it follows; "RAST/" = /root/reference/submodules/diff-surfel-rasterization/)
and lets autograd differentiate it.  Autograd yields the TRUE gradient of the forward; the reference's hand-written
backward deliberately deviates from it in a few terms (no quaternion-normalisation
Jacobian, clamp ignored, ...), so this file pins the forward only (tests/test_torch_cpu_blend.py against the golden
vectors captured from the reference extension) and is the *timing* baseline the north_star names; the gradient
oracle is oracle/surfel_oracle.c.

Formulation: all (tile, Gaussian) instances are sorted by (tile, depth, id) like rasterizer_impl.cu:301-309, padded
to the longest tile list, and blended for all tiles at once as [tiles, list, 256 pixels] tensors; the front-to-back
recurrence becomes an exclusive cumulative product of (1 - alpha) with the reference's two stopping rules applied as
masks (alpha < 1/255 skipped, forward.cu:386; everything from the first splat that would push T below 1e-4 on is
dropped, forward.cu:388-393).
"""
from __future__ import annotations

import math

import torch

TILE = 16
NEAR_N, FAR_N = 0.2, 100.0
FILTER_SIZE, FILTER_INV_SQUARE = 0.707106, 2.0
SH_C0 = 0.28209479177387814


def _quat_to_rot(q):
    """auxiliary.h:213-235 (w, x, y, z), normalised in-kernel.  Returns the three COLUMNS R[:, c]."""
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    c0 = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)], 1)
    c1 = torch.stack([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)], 1)
    c2 = torch.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], 1)
    return c0, c1, c2


def rasterize(means3D, shs, opacities, scales, rotations, viewmatrix, projmatrix, campos, W, H, bg, sh_degree=0,
              scale_modifier=1.0):
    """-> (color [3,H,W], allmap [7,H,W], radii [P] int32, num_rendered).  SH degree 0 only (configs[0])."""
    if sh_degree != 0:
        raise NotImplementedError("the pure-PyTorch CPU baseline covers BASELINE configs[0] (SH degree 0)")
    P = means3D.shape[0]
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    ones = torch.ones(P, 1, dtype=means3D.dtype)
    # ---- preprocess (forward.cu:148-251) ----
    p_view = torch.cat([means3D, ones], 1) @ viewmatrix                     # auxiliary.h:185-210
    c0, c1, c2 = _quat_to_rot(rotations)
    L0, L1 = c0 * (scale_modifier * scales[:, 0:1]), c1 * (scale_modifier * scales[:, 1:2])   # forward.cu:75-90
    zeros = torch.zeros(P, 1, dtype=means3D.dtype)
    rows = [torch.cat([L0, zeros], 1) @ projmatrix, torch.cat([L1, zeros], 1) @ projmatrix,
            torch.cat([means3D, ones], 1) @ projmatrix]                     # h_r = a_r . world2ndc
    # ndc2pix (forward.cu:97-107): x_pix = h.x W/2 + h.w (W-1)/2, y likewise, w = h.w
    Tu = torch.stack([h[:, 0] * (W / 2) + h[:, 3] * ((W - 1) / 2) for h in rows], 1)
    Tv = torch.stack([h[:, 1] * (H / 2) + h[:, 3] * ((H - 1) / 2) for h in rows], 1)
    Tw = torch.stack([h[:, 3] for h in rows], 1)
    normal = c2 @ viewmatrix[:3, :3]                                        # forward.cu:113
    cosv = -(p_view[:, :3] * normal).sum(1)                                 # forward.cu:209-214 (DUAL_VISIABLE)
    normal = normal * torch.where(cosv > 0, 1.0, -1.0).unsqueeze(1)
    # compute_aabb (forward.cu:119-145), cutoff 3
    t = torch.tensor([9.0, 9.0, -1.0], dtype=means3D.dtype)
    d = (t * Tw * Tw).sum(1)
    f = t / d.unsqueeze(1)
    cx, cy = (f * Tu * Tw).sum(1), (f * Tv * Tw).sum(1)
    ex = torch.sqrt(torch.clamp_min(cx * cx - (f * Tu * Tu).sum(1), 1e-4))
    ey = torch.sqrt(torch.clamp_min(cy * cy - (f * Tv * Tv).sum(1), 1e-4))
    radius = torch.ceil(torch.maximum(torch.maximum(ex, ey), torch.tensor(3.0 * FILTER_SIZE)))
    with torch.no_grad():
        ri = radius.to(torch.int32).to(means3D.dtype)
        x0 = ((cx - ri) / TILE).to(torch.int64).clamp(0, gx)                # auxiliary.h:67-77 (truncation)
        y0 = ((cy - ri) / TILE).to(torch.int64).clamp(0, gy)
        x1 = ((cx + ri + TILE - 1) / TILE).to(torch.int64).clamp(0, gx)
        y1 = ((cy + ri + TILE - 1) / TILE).to(torch.int64).clamp(0, gy)
        ok = (p_view[:, 2] > 0.2) & (cosv != 0) & (d != 0) & ((x1 - x0) * (y1 - y0) > 0)
        radii = torch.where(ok, ri.to(torch.int32), torch.zeros(P, dtype=torch.int32))
        # ---- binning (rasterizer_impl.cu:70-138, 278-318): instances sorted by (tile, depth, id) ----
        ids = torch.nonzero(ok).squeeze(1)
        nx, ny = (x1 - x0)[ids], (y1 - y0)[ids]
        cnt = nx * ny
        R = int(cnt.sum())
        inst_g = torch.repeat_interleave(ids, cnt)
        local = torch.arange(R) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
        w_ = torch.repeat_interleave(nx, cnt)
        tile = (torch.repeat_interleave(y0[ids], cnt) + local // w_) * gx + torch.repeat_interleave(x0[ids], cnt) + local % w_
        depth_bits = p_view[inst_g, 2].contiguous().view(torch.int32).to(torch.int64)   # positive floats order like ints
        order = torch.argsort(tile * (1 << 32) + depth_bits, stable=True)
        inst_g, tile = inst_g[order], tile[order]
        counts = torch.bincount(tile, minlength=gx * gy)
        n_max = max(int(counts.max()), 1)
        start = torch.cumsum(counts, 0) - counts
        slot = torch.arange(R) - start[tile]
        pad = torch.full((gx * gy, n_max), -1, dtype=torch.int64)
        pad[tile, slot] = inst_g
        valid = pad >= 0
        gidx = pad.clamp_min(0)                                             # [tiles, n]
    # SH degree 0 colour (forward.cu:20-71): clamp(C0 sh0 + 0.5, 0)
    rgb = torch.clamp_min(SH_C0 * shs[:, 0, :] + 0.5, 0.0)

    # ---- blend (forward.cu:256-448) for all tiles at once: [tiles, n, 256] ----
    tx = (torch.arange(gx * gy) % gx) * TILE
    ty = (torch.arange(gx * gy) // gx) * TILE
    px = (tx[:, None] + (torch.arange(TILE * TILE) % TILE)[None, :]).to(means3D.dtype)[:, None, :]   # [tiles,1,256]
    py = (ty[:, None] + (torch.arange(TILE * TILE) // TILE)[None, :]).to(means3D.dtype)[:, None, :]
    g = lambda a: a[gidx]                                                   # noqa: E731  [tiles, n, ...]
    Tu_, Tv_, Tw_ = g(Tu), g(Tv), g(Tw)
    k = [px * Tw_[..., i:i + 1] - Tu_[..., i:i + 1] for i in range(3)]     # forward.cu:362-366
    l = [py * Tw_[..., i:i + 1] - Tv_[..., i:i + 1] for i in range(3)]
    pz = k[0] * l[1] - k[1] * l[0]
    nz = pz != 0
    pz_safe = torch.where(nz, pz, torch.ones_like(pz))
    sx = (k[1] * l[2] - k[2] * l[1]) / pz_safe
    sy = (k[2] * l[0] - k[0] * l[2]) / pz_safe
    rho3d = sx * sx + sy * sy
    dx, dy = g(cx)[..., None] - px, g(cy)[..., None] - py
    rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy)
    rho = torch.minimum(rho3d, rho2d)                                       # forward.cu:373-376
    depth = sx * Tw_[..., 0:1] + sy * Tw_[..., 1:2] + Tw_[..., 2:3]        # forward.cu:379
    alpha = torch.clamp_max(g(opacities)[..., 0:1] * torch.exp(-0.5 * rho), 0.99)
    live = valid[..., None] & nz & ~(depth < NEAR_N) & ~(alpha < 1.0 / 255.0)   # forward.cu:368,380,386
    a = torch.where(live, alpha, torch.zeros_like(alpha))
    T_after = torch.cumprod(1 - a, dim=1)                                   # T after each list position
    T_before = torch.cat([torch.ones_like(T_after[:, :1]), T_after[:, :-1]], 1)
    stop = live & (T_after < 1e-4)                                          # forward.cu:388-393: this splat ends the pixel
    dead = torch.cumsum(stop.to(torch.int32), 1) > 0                        # ... and is itself not blended
    contrib = live & ~dead
    w = torch.where(contrib, a * T_before, torch.zeros_like(a))             # forward.cu:395
    T_final = torch.where(contrib, 1 - a, torch.ones_like(a)).prod(1)       # [tiles, 256]

    def acc(v):
        return (w * v).sum(1)

    m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / torch.where(contrib, depth, torch.ones_like(depth)))   # forward.cu:399
    A_before = 1 - T_before
    M1_before = torch.cumsum(w * m, 1) - w * m
    M2_before = torch.cumsum(w * m * m, 1) - w * m * m
    distortion = acc(m * m * A_before + M2_before - 2 * m * M1_before)     # forward.cu:400-404
    D = acc(depth)
    Nn = [acc(g(normal)[..., i:i + 1]) for i in range(3)]
    C = [acc(g(rgb)[..., i:i + 1]) for i in range(3)]
    # median depth: depth of the LAST contributor blended while T > 0.5 (forward.cu:406-411)
    med_mask = contrib & (T_before > 0.5)
    pos = torch.arange(a.shape[1])[None, :, None]
    last_med = torch.where(med_mask, pos, torch.full_like(pos, -1)).max(1).values
    median = torch.where(last_med >= 0, torch.gather(depth, 1, last_med.clamp_min(0)[:, None, :]).squeeze(1),
                         torch.zeros_like(T_final))

    def to_img(v):   # [tiles, 256] -> [H, W]
        return v.reshape(gy, gx, TILE, TILE).permute(0, 2, 1, 3).reshape(gy * TILE, gx * TILE)[:H, :W]

    color = torch.stack([to_img(C[i] + T_final * bg[i]) for i in range(3)], 0)
    allmap = torch.stack([to_img(D), to_img(1 - T_final), to_img(Nn[0]), to_img(Nn[1]), to_img(Nn[2]), to_img(median),
                          to_img(distortion)], 0)
    return color, allmap, radii, R


def forward_backward(scene, cam, grads, threads=None):
    """One fwd+bwd of configs[0] on the host cores; returns (outputs dict, seconds)."""
    import time
    if threads:
        torch.set_num_threads(int(threads))
    leaves = {k: scene[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    t0 = time.perf_counter()
    color, allmap, radii, R = rasterize(leaves["means3D"], leaves["shs"], leaves["opacities"], leaves["scales"],
                                        leaves["rotations"], cam.viewmatrix, cam.projmatrix, cam.campos, cam.width,
                                        cam.height, torch.zeros(3), int(scene["sh_degree"]))
    torch.autograd.backward([color, allmap], [grads[0], grads[1]])
    dt = time.perf_counter() - t0
    out = {"color": color.detach().numpy(), "allmap": allmap.detach().numpy(), "radii": radii.numpy(), "num_rendered": R}
    out.update({"g_" + k: v.grad.numpy() for k, v in leaves.items()})
    return out, dt
