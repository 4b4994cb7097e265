"""world_size-2 gloo test of the multi-GPU choreography (streetunveiler_b200/sharded.py) on CPU.

The CUDA stages are replaced by a tiny differentiable torch "rasterizer" with the same backend
interface, so what is exercised is the host logic: the global tile histogram, the cost-balanced
partition of the row-major tile order, routing of every record to the ranks whose tile range its
rect touches (stable, index order kept), the all-to-all with unequal splits in both directions,
the image all-reduce, accumulation of the returned gradient rows and gradient routing to the
owners.  The sharded result must equal the unsharded one."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streetunveiler_b200 import sharded

W, H = 48, 40  # 3 x 3 tiles (last row partial)
GX, GY = 3, 3
RADIUS = 10


def _rects(rec, radii):
    """Tile rect [x0, x1) x [y0, y1) of every record, the reference's formula (auxiliary.h:67-77)."""
    r = radii.to(torch.float32)
    x0 = ((rec[:, 0] - r) / 16).to(torch.int64).clamp(0, GX)
    y0 = ((rec[:, 1] - r) / 16).to(torch.int64).clamp(0, GY)
    x1 = ((rec[:, 0] + r + 15) / 16).to(torch.int64).clamp(0, GX)
    y1 = ((rec[:, 1] + r + 15) / 16).to(torch.int64).clamp(0, GY)
    return x0, y0, x1, y1


def _tile_cover(rec, radii):
    """[P, tiles] bool: tile inside the record's rect (and the record visible)."""
    x0, y0, x1, y1 = _rects(rec, radii)
    t = torch.arange(GX * GY)
    tx, ty = (t % GX)[None, :], (t // GX)[None, :]
    return (radii > 0)[:, None] & (tx >= x0[:, None]) & (tx < x1[:, None]) & (ty >= y0[:, None]) & (ty < y1[:, None])


class ToyBackend:
    """records: [x, y, opacity, r, g, b, 0...]; 'blend' = sum of blobs truncated to their tile rect (order-independent toy)."""

    def shard_preprocess(self, s, means3D, shs, opacities, scales, rotations):
        P = means3D.shape[0]
        rec = torch.zeros(P, sharded.REC_FLOATS)
        rec[:, 0:2] = means3D[:, 0:2]
        rec[:, 2] = opacities[:, 0]
        rec[:, 3:6] = shs[:, 0, :]
        radii = torch.where(means3D[:, 2] < 4.5, RADIUS, 0).to(torch.int32)   # a few Gaussians are "culled": never exchanged
        keys = means3D[:, 2].contiguous().view(torch.int32).clone()
        return radii, rec, keys, torch.zeros(P, dtype=torch.uint8)

    def tile_hist(self, s, rec, radii):
        return _tile_cover(rec, radii).sum(0).to(torch.int32)

    def partition(self, s, hist, world, shares=None):
        cost = hist.to(torch.int64) + sharded.COST_BASE
        excl = torch.cumsum(cost, 0) - cost
        shares = [1.0 / world] * world if shares is None else shares
        target = torch.cumsum(torch.tensor([0.0] + list(shares[:-1]), dtype=torch.float64), 0) / sum(shares)   # [G]
        rank_of = (excl[:, None].double() / float(cost.sum()) >= target[None, 1:]).sum(1)
        cuts = torch.tensor([int((rank_of < k).sum()) for k in range(world)] + [GX * GY], dtype=torch.int32)
        window_R = torch.tensor([int(hist[int(cuts[k]):int(cuts[k + 1])].sum()) for k in range(world)], dtype=torch.int64)
        return cuts, window_R

    def route_count(self, s, rec, radii, cuts, world, extra=0):
        cover = _tile_cover(rec, radii)
        mask = torch.stack([cover[:, int(cuts[d]):int(cuts[d + 1])].any(1) for d in range(world)], 1)   # [P, G]
        return mask, torch.cat([mask.sum(0).to(torch.int32), torch.tensor([extra], dtype=torch.int32)])

    def route_scatter(self, rec, radii, keys, mask, send_counts, n_send, world):
        rows, src = [], []
        for d in range(world):
            idx = torch.nonzero(mask[:, d]).squeeze(1)                      # index order kept (stable)
            row = torch.zeros(idx.numel(), sharded.XROW_FLOATS)
            row[:, :sharded.REC_FLOATS] = rec[idx]
            row[:, 24] = keys[idx].view(torch.float32)
            row[:, 25] = radii[idx].view(torch.float32)
            rows.append(row)
            src.append(idx.to(torch.int32))
        rows, src = torch.cat(rows, 0), torch.cat(src, 0)
        assert rows.shape[0] == n_send == int(send_counts[:world].sum())
        return rows, src

    def unpack(self, rows):
        return (rows[:, :sharded.REC_FLOATS].contiguous(), rows[:, 25].contiguous().view(torch.int32),
                rows[:, 24].contiguous().view(torch.int32))

    @staticmethod
    def _render(rec, radii, tile_lo, tile_hi):
        ys = torch.arange(H, dtype=torch.float32)[:, None, None]
        xs = torch.arange(W, dtype=torch.float32)[None, :, None]
        tile_of_pixel = (torch.arange(H)[:, None] // 16) * GX + (torch.arange(W)[None, :] // 16)   # [H, W]
        cover = _tile_cover(rec.detach(), radii)                            # [P, tiles]
        live = cover[:, tile_of_pixel.reshape(-1)].t().reshape(H, W, -1)    # pixel's tile inside the rect
        live = live & ((tile_of_pixel >= tile_lo) & (tile_of_pixel < tile_hi))[:, :, None]
        w = live * rec[:, 2] * torch.exp(-((xs - rec[:, 0]) ** 2 + (ys - rec[:, 1]) ** 2) / 60.0)   # [H, W, P]
        return torch.einsum("ywp,pc->cyw", w, rec[:, 3:6])

    def window_forward(self, s, rec_w, radii_w, keys_w, tile_lo, tile_hi, R):
        assert R == int(_tile_cover(rec_w, radii_w)[:, tile_lo:tile_hi].sum())   # the partition's count is exact
        planes = torch.zeros(10, H, W)
        img = self._render(rec_w, radii_w, tile_lo, tile_hi)
        planes[:3] = img
        planes[4] = img.sum(0)
        return planes, (radii_w, tile_lo, tile_hi)

    def window_backward(self, s, rec_w, state, dL_dcolor, dL_dothers):
        radii_w, tile_lo, tile_hi = state
        grows = torch.zeros(rec_w.shape[0], sharded.GREC_FLOATS)
        if rec_w.shape[0]:
            with torch.enable_grad():
                r = rec_w.detach().clone().requires_grad_(True)
                img = self._render(r, radii_w, tile_lo, tile_hi)
                loss = (img * dL_dcolor).sum() + (img.sum(0) * dL_dothers[1]).sum()
                (grad,) = torch.autograd.grad(loss, r)
            grows[:, :6] = grad[:, :6]
        return grows

    def grad_accumulate(self, P, rows, send_src):
        gacc = torch.zeros(P, sharded.GREC_FLOATS)
        gacc.index_add_(0, send_src.long(), rows)
        return gacc

    def shard_backward(self, s, means3D, shs, scales, rotations, radii, rec, clamped, grec, slot):
        assert slot is None
        P, M = means3D.shape[0], shs.shape[1]
        g = {"means3D": torch.zeros(P, 3), "means2D": torch.zeros(P, 3), "opacities": grec[:, 2:3].clone(),
             "shs": torch.zeros(P, M, 3), "scales": torch.zeros(P, 2), "rotations": torch.zeros(P, 4)}
        g["means3D"][:, 0:2] = grec[:, 0:2]
        g["shs"][:, 0, :] = grec[:, 3:6]
        return g


def _scene(P):
    g = torch.Generator().manual_seed(5)
    return {"means3D": torch.stack([torch.rand(P, generator=g) * W, torch.rand(P, generator=g) * H,
                                    torch.rand(P, generator=g) * 5 + 1], 1),
            "shs": torch.randn(P, 4, 3, generator=g), "opacities": torch.rand(P, 1, generator=g),
            "scales": torch.rand(P, 2, generator=g), "rotations": torch.randn(P, 4, generator=g)}


def _run(rast, sc, lo, hi):
    p = {k: v[lo:hi].clone().requires_grad_(True) for k, v in sc.items()}
    m2 = torch.zeros(hi - lo, 3, requires_grad=True)
    st = SimpleNamespace(image_width=W, image_height=H, sh_degree=1, debug=False)
    color, radii, allmap = rast(p["means3D"], m2, p["opacities"], p["shs"], p["scales"], p["rotations"], st)
    g = torch.Generator().manual_seed(9)
    dc, da = torch.randn(3, H, W, generator=g), torch.randn(7, H, W, generator=g)
    torch.autograd.backward([color, allmap], [dc, da])
    return color.detach(), allmap.detach(), {k: v.grad for k, v in p.items()}


def _worker(rank, world, port, splits, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = _scene(splits[-1])
        rast = sharded.ShardedRasterizer(backend=ToyBackend())
        color, allmap, grads = _run(rast, sc, splits[rank], splits[rank + 1])
        torch.save({"color": color, "allmap": allmap, "grads": grads, "info": dict(sharded.LAST_INFO)},
                   os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


import pytest


@pytest.mark.parametrize("splits", [[0, 23, 40], [0, 9, 9, 40]])      # unequal shards; an EMPTY shard in the 3-rank case
def test_sharded_equals_unsharded(tmp_path, splits):
    world = len(splits) - 1
    mp.spawn(_worker, args=(world, _free_port(), splits, str(tmp_path)), nprocs=world, join=True)
    # unsharded reference: same toy backend, world of one
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        sc = _scene(splits[-1])
        ref_color, ref_allmap, ref_grads = _run(sharded.ShardedRasterizer(backend=ToyBackend()), sc, 0, splits[-1])
    finally:
        dist.destroy_process_group()
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for r in res:                            # full image on every rank, equal to the single-process one
        assert torch.allclose(r["color"], ref_color, atol=1e-5)
        assert torch.allclose(r["allmap"], ref_allmap, atol=1e-5)
    assert float(ref_color.abs().max()) > 0
    info = [r["info"] for r in res]
    assert all(i["cuts"] == info[0]["cuts"] for i in info) and info[0]["cuts"][0] == 0 and info[0]["cuts"][-1] == GX * GY
    assert sum(sum(i["send"]) for i in info) == sum(sum(i["recv"]) for i in info)
    assert sum(sum(i["send"]) for i in info) > int((_scene(splits[-1])["means3D"][:, 2] < 4.5).sum())   # some records go to 2 ranks
    for k in ("means3D", "shs", "opacities"):
        got = torch.cat([r["grads"][k] for r in res], 0)
        assert got.shape == ref_grads[k].shape
        assert torch.allclose(got, ref_grads[k], atol=1e-4), k
        assert float(ref_grads[k].abs().max()) > 0


def _worker_mismatch(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = _scene(20)
        rast = sharded.ShardedRasterizer(backend=ToyBackend(), check_replicated=True)
        p = {k: v[rank * 10:(rank + 1) * 10].clone().requires_grad_(True) for k, v in sc.items()}
        m2 = torch.zeros(10, 3, requires_grad=True)
        st = SimpleNamespace(image_width=W, image_height=H, sh_degree=1, debug=False)
        color, radii, allmap = rast(p["means3D"], m2, p["opacities"], p["shs"], p["scales"], p["rotations"], st)
        msg = "no error"
        try:   # every rank scales the loss differently: the replicated-loss contract is violated
            torch.autograd.backward([color, allmap], [torch.ones(3, H, W) * (rank + 1), torch.zeros(7, H, W)])
        except RuntimeError as e:
            msg = str(e)
        open(os.path.join(out, f"m{rank}.txt"), "w").write(msg)
    finally:
        dist.destroy_process_group()


def test_replicated_loss_contract_is_checked(tmp_path):
    mp.spawn(_worker_mismatch, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert "replicated-loss contract" in open(os.path.join(tmp_path, f"m{r}.txt")).read()


def test_pad_rows_and_toy_partition():
    x = torch.arange(6.0).reshape(3, 2)
    y = sharded.pad_rows(x, 5, -1)
    assert y.shape == (5, 2) and torch.equal(y[:3], x) and bool((y[3:] == -1).all())
    assert sharded.pad_rows(x, 3) is not None
    hist = torch.tensor([0, 0, 0, 50, 0, 0, 0, 0, 50], dtype=torch.int32)
    cuts, wr = ToyBackend().partition(None, hist, 2)
    assert cuts.tolist()[0] == 0 and cuts.tolist()[-1] == 9 and 3 < cuts.tolist()[1] <= 8 and wr.tolist() == [50, 50]
