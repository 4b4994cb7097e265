"""world_size-2 gloo test of the multi-GPU choreography (streetunveiler_b200/sharded.py) on CPU.

The CUDA stages are replaced by a tiny differentiable torch "rasterizer" with the same backend
interface, so what is exercised is the host logic: padding of unequal shards, all-gather order,
tile-row windows (interleaved rows), compaction of the culled rows, image all-reduce, reduce-scatter of gradient records back to the
owners, gradient routing.  The sharded result must equal the unsharded one."""
import os
import socket
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streetunveiler_b200 import sharded

W, H = 48, 40  # 3 x 3 tiles (last row partial)


class ToyBackend:
    """records: [x, y, opacity, r, g, b, 0...]; 'blend' = sum of isotropic blobs (order-independent toy)."""

    def shard_preprocess(self, s, means3D, shs, opacities, scales, rotations):
        P = means3D.shape[0]
        rec = torch.zeros(P, sharded.REC_FLOATS)
        rec[:, 0:2] = means3D[:, 0:2]
        rec[:, 2] = opacities[:, 0]
        rec[:, 3:6] = shs[:, 0, :]
        radii = (means3D[:, 2] < 4.5).to(torch.int32)        # a few Gaussians are "culled": never exchanged
        keys = means3D[:, 2].contiguous().view(torch.int32).clone()
        return radii, rec, keys, torch.zeros(P, dtype=torch.uint8)

    def shard_compact(self, radii, rec, keys):
        P = radii.shape[0]
        vis = radii > 0
        n = int(vis.sum())
        rec_c, radii_c, keys_c = torch.full_like(rec, float("nan")), torch.zeros_like(radii), torch.full_like(keys, -1)
        rec_c[:n], radii_c[:n], keys_c[:n] = rec[vis], radii[vis], keys[vis]
        slot = torch.full((P,), -1, dtype=torch.int32)
        slot[vis] = torch.arange(n, dtype=torch.int32)
        return rec_c, radii_c, keys_c, slot, torch.tensor([n], dtype=torch.int32)

    @staticmethod
    def _render(rec, radii, rows):
        ys = torch.tensor(rows, dtype=torch.float32)[:, None, None]
        xs = torch.arange(W, dtype=torch.float32)[None, :, None]
        live = (radii > 0)[None, None, :]
        rec = torch.where((radii > 0)[:, None], rec, torch.zeros_like(rec))   # padded rows hold garbage
        w = live * rec[:, 2] * torch.exp(-((xs - rec[:, 0]) ** 2 + (ys - rec[:, 1]) ** 2) / 60.0)   # [rows, W, P]
        return torch.einsum("ywp,pc->cyw", w, rec[:, 3:6])

    @staticmethod
    def _rows(row_offset, row_stride):
        gy = (H + 15) // 16
        return [y for ty in range(row_offset, gy, row_stride) for y in range(ty * 16, min(ty * 16 + 16, H))]

    def window_forward(self, s, rec_all, radii_all, keys_all, row_offset, row_stride):
        rows = self._rows(row_offset, row_stride)
        color = torch.zeros(3, H, W)
        others = torch.zeros(7, H, W)
        if rows:
            img = self._render(rec_all, radii_all, rows)
            color[:, rows, :] = img
            others[1, rows, :] = img.sum(0)
        return color, others, (radii_all, rows)

    def window_backward(self, s, rec_all, state, dL_dcolor, dL_dothers):
        radii_all, rows = state
        grec = torch.zeros(rec_all.shape[0], sharded.GREC_FLOATS)
        if rows:
            with torch.enable_grad():
                r = rec_all.detach().clone().requires_grad_(True)
                img = self._render(r, radii_all, rows)
                loss = (img * dL_dcolor[:, rows, :]).sum() + (img.sum(0) * dL_dothers[1, rows, :]).sum()
                (grad,) = torch.autograd.grad(loss, r)
            grec[:, :6] = grad[:, :6]
        return grec

    def shard_backward(self, s, means3D, shs, scales, rotations, radii, rec, clamped, grec, slot):
        P, M = means3D.shape[0], shs.shape[1]
        vis = radii > 0
        dense = torch.zeros(P, sharded.GREC_FLOATS)
        dense[vis] = grec[slot[vis].long()]
        grec = dense
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
        torch.save({"color": color, "allmap": allmap, "grads": grads}, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_sharded_equals_unsharded_world2(tmp_path):
    splits = [0, 7, 11]                      # unequal shards: 7 and 4 Gaussians -> padding path
    mp.spawn(_worker, args=(2, _free_port(), splits, str(tmp_path)), nprocs=2, join=True)
    # unsharded reference: same toy backend, world of one
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        sc = _scene(11)
        ref_color, ref_allmap, ref_grads = _run(sharded.ShardedRasterizer(backend=ToyBackend()), sc, 0, 11)
    finally:
        dist.destroy_process_group()
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2)]
    for r in res:                            # full image on every rank, equal to the single-process one
        assert torch.allclose(r["color"], ref_color, atol=1e-5)
        assert torch.allclose(r["allmap"], ref_allmap, atol=1e-5)
    assert float(ref_color.abs().max()) > 0
    for k in ("means3D", "shs", "opacities"):
        got = torch.cat([res[0]["grads"][k], res[1]["grads"][k]], 0)
        assert got.shape == ref_grads[k].shape
        assert torch.allclose(got, ref_grads[k], atol=1e-4), k
        assert float(ref_grads[k].abs().max()) > 0


def test_pad_rows_and_window_rows():
    x = torch.arange(6.0).reshape(3, 2)
    y = sharded.pad_rows(x, 5, -1)
    assert y.shape == (5, 2) and torch.equal(y[:3], x) and bool((y[3:] == -1).all())
    assert sharded.pad_rows(x, 3) is not None
    rows0, rows1 = ToyBackend._rows(0, 2), ToyBackend._rows(1, 2)
    assert sorted(rows0 + rows1) == list(range(H)) and not set(rows0) & set(rows1)
