"""Multi-GPU surfel rasterization: Gaussians sharded by index, one process per GPU (torch.distributed).

Why not "every rank blends its own Gaussians and the images are summed": front-to-back alpha blending
is an ordered product, so a sum of per-shard images is not the reference image.  The
exact scheme used here keeps parameters, gradients and optimiser state sharded by Gaussian index and
exchanges only *projected* data:

  forward   1. every rank projects its own shard            (surfel_shard_preprocess, 96-B records) and keeps
               the rows that survived culling, in index order (surfel_shard_compact) -- typically a
               fifth of a street scene is in view, so the exchanges below shrink by that factor
            2. all-gather of the compact records / radii / depth keys, padded to the largest count
               (NCCL over NVLink; one tiny all-gather of the counts first)
            3. every rank bins + blends ALL Gaussians for its own tile rows (rows r, r+G, r+2G, ...;
               interleaved for load balance)               (surfel_window_prepare / _render)
            4. all-reduce(sum) of the ten image planes      (each rank wrote only its rows, the rest is 0)
  backward  5. every rank back-propagates its tile rows into 80-B gradient records of ALL Gaussians
                                                           (surfel_window_backward)
            6. reduce-scatter(sum) of the gradient records to the owners of the Gaussians
            7. every rank turns its records into parameter gradients (surfel_shard_backward; Gaussian i
               reads compact row slot[i])

Every pixel sees exactly the list, order and arithmetic of the single-GPU path, so the forward is
bit-identical to it and gradients differ only by float-add order.

The compute stages sit behind a small backend object so that the collective choreography can be
tested on CPU (gloo, world_size 2) with a toy backend (tests/test_sharded_gloo.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

REC_FLOATS = 24
GREC_FLOATS = 20
ROW_QUANTUM = 4096   # exchanged row counts are rounded up to this (stable buffer shapes from step to step)

# SURFEL_SHARD_TIMING=1: synchronise around every phase and record wall-clock per phase and call (diagnostics only)
import os as _os
import time as _time
_TIMING = _os.environ.get("SURFEL_SHARD_TIMING") == "1"
PHASE_MS = {}


class _phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _TIMING:
            torch.cuda.synchronize()
            self.t = _time.perf_counter()

    def __exit__(self, *a):
        if _TIMING:
            torch.cuda.synchronize()
            PHASE_MS.setdefault(self.name, []).append((_time.perf_counter() - self.t) * 1e3)


# ----------------------------------------------------------------------------------------------------
# collectives with gloo fallbacks (gloo has no reduce_scatter; used only by the CPU tests)
# ----------------------------------------------------------------------------------------------------
def _all_gather_rows(x: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
    try:
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        out = torch.cat(parts, 0)
    return out


def _reduce_scatter_rows(x: torch.Tensor, group) -> torch.Tensor:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0] // world
    out = x.new_empty((n,) + tuple(x.shape[1:]))
    if dist.get_backend(group) == "gloo":
        y = x.clone()
        dist.all_reduce(y, group=group)
        return y[rank * n:(rank + 1) * n].contiguous()
    dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
    return out


def pad_rows(x: torch.Tensor, n: int, value=0) -> torch.Tensor:
    """Pad the first dimension to n rows (shards may differ in size; collectives need equal shapes)."""
    if x.shape[0] == n:
        return x.contiguous()
    pad = x.new_full((n - x.shape[0],) + tuple(x.shape[1:]), value)
    return torch.cat([x, pad], 0).contiguous()


# ----------------------------------------------------------------------------------------------------
# native backend: the CUDA library through its C ABI
# ----------------------------------------------------------------------------------------------------
class NativeBackend:
    def __init__(self):
        from . import _lib
        self._lib = _lib
        self.L = _lib.lib()

    @staticmethod
    def _p(t):
        return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def shard_preprocess(self, s, means3D, shs, opacities, scales, rotations):
        P, dev = means3D.shape[0], means3D.device
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        rec = torch.empty((P, REC_FLOATS), dtype=torch.float32, device=dev)
        keys = torch.empty((P,), dtype=torch.int32, device=dev)
        clamped = torch.empty((P,), dtype=torch.uint8, device=dev)
        M = shs.shape[1]
        p = self._p
        self._lib.check(self.L.surfel_shard_preprocess(
            P, int(s.sh_degree), M, s.image_width, s.image_height, p(means3D), p(shs), None, p(opacities), p(scales),
            float(s.scale_modifier), p(rotations), None, p(s.viewmatrix), p(s.projmatrix), p(s.campos),
            float(s.tanfovx), float(s.tanfovy), int(bool(s.prefiltered)), p(radii), p(rec), p(keys), p(clamped),
            self._stream()), "surfel_shard_preprocess")
        return radii, rec, keys, clamped

    def shard_compact(self, radii, rec, keys):
        """-> (rec_c, radii_c, keys_c [P rows of capacity each], slot [P], count [1] int32 on the device)."""
        P, dev = radii.shape[0], radii.device
        rec_c = torch.empty_like(rec)
        radii_c = torch.empty_like(radii)
        keys_c = torch.empty_like(keys)
        slot = torch.empty((P,), dtype=torch.int32, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        tmp = torch.empty((self._lib.size(self.L.surfel_shard_compact_bytes(P), "surfel_shard_compact_bytes"),),
                          dtype=torch.uint8, device=dev)
        p = self._p
        self._lib.check(self.L.surfel_shard_compact(P, p(radii), p(rec), p(keys), p(rec_c), p(radii_c), p(keys_c),
                                                    p(slot), C.c_void_p(count.data_ptr()), p(tmp), self._stream()),
                        "surfel_shard_compact")
        return rec_c, radii_c, keys_c, slot, count

    def window_forward(self, s, rec_all, radii_all, keys_all, row_offset, row_stride):
        L, p, dev = self.L, self._p, rec_all.device
        Pt, W, H = rec_all.shape[0], s.image_width, s.image_height
        u8 = dict(dtype=torch.uint8, device=dev)
        win = torch.empty((self._lib.size(L.surfel_window_bytes(Pt), "surfel_window_bytes"),), **u8)
        img = torch.empty((self._lib.size(L.surfel_image_bytes(W, H), "surfel_image_bytes"),), **u8)
        R = C.c_int64(0)
        st = self._stream()
        self._lib.check(L.surfel_window_prepare(Pt, W, H, row_offset, row_stride, p(rec_all), p(radii_all), p(keys_all),
                                                p(win), C.byref(R), st, int(bool(s.debug))), "surfel_window_prepare")
        R = int(R.value)
        binb = torch.empty((self._lib.size(L.surfel_binning_bytes(R), "surfel_binning_bytes") if R else 0,), **u8)
        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)   # rows of other ranks stay 0 for the all-reduce
        others = torch.zeros((7, H, W), dtype=torch.float32, device=dev)
        self._lib.check(L.surfel_window_render(Pt, W, H, row_offset, row_stride, R, p(s.bg), p(rec_all), p(radii_all),
                                               p(win), p(binb), p(img), p(color), p(others), st, int(bool(s.debug))),
                        "surfel_window_render")
        return color, others, (R, binb, img, row_offset, row_stride)

    def window_backward(self, s, rec_all, state, dL_dcolor, dL_dothers):
        R, binb, img, row_offset, row_stride = state
        Pt = rec_all.shape[0]
        grec = torch.empty((Pt, GREC_FLOATS), dtype=torch.float32, device=rec_all.device)
        p = self._p
        self._lib.check(self.L.surfel_window_backward(
            Pt, s.image_width, s.image_height, row_offset, row_stride, R, p(s.bg), p(rec_all), p(binb), p(img),
            p(dL_dcolor.contiguous()), p(dL_dothers.contiguous()), p(grec), self._stream(), int(bool(s.debug))),
            "surfel_window_backward")
        return grec

    def shard_backward(self, s, means3D, shs, scales, rotations, radii, rec, clamped, grec, slot):
        P, dev, M = means3D.shape[0], means3D.device, shs.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        g = {"means2D": torch.empty((P, 3), **f32), "opacities": torch.empty((P, 1), **f32),
             "colors": torch.empty((P, 3), **f32), "means3D": torch.empty((P, 3), **f32),
             "transMat": torch.empty((P, 9), **f32), "shs": torch.empty((P, M, 3), **f32),
             "scales": torch.empty((P, 2), **f32), "rotations": torch.empty((P, 4), **f32)}
        p = self._p
        self._lib.check(self.L.surfel_shard_backward(
            P, int(s.sh_degree), M, s.image_width, s.image_height, p(means3D), p(shs), p(scales), p(rotations), None,
            p(s.viewmatrix), p(s.projmatrix), p(s.campos), float(s.tanfovx), float(s.tanfovy), p(radii), p(rec),
            p(clamped), p(grec.contiguous()), p(slot), p(g["means2D"]), None, p(g["opacities"]), p(g["colors"]), p(g["means3D"]),
            p(g["transMat"]), p(g["shs"]), p(g["scales"]), p(g["rotations"]), self._stream()), "surfel_shard_backward")
        return g


# ----------------------------------------------------------------------------------------------------
class _ShardedRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, opacities, scales, rotations, settings, backend, group):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        with _phase("fwd preprocess"):
            radii, rec, keys, clamped = backend.shard_preprocess(settings, means3D.contiguous(), shs.contiguous(),
                                                                 opacities.contiguous(), scales.contiguous(),
                                                                 rotations.contiguous())
            rec_c, radii_c, keys_c, slot, count = backend.shard_compact(radii, rec, keys)
        # rows past a rank's count are culled Gaussians: radius 0, depth key 0xFFFFFFFF (sort last, emit nothing)
        with _phase("fwd counts"):
            counts = _all_gather_rows(count, group)
            c_max = -(-max(int(counts.max().item()), 1) // ROW_QUANTUM) * ROW_QUANTUM   # the one host sync of the exchange
        with _phase("fwd all-gather"):
            keep = min(c_max, rec_c.shape[0])
            rec_all = _all_gather_rows(pad_rows(rec_c[:keep], c_max), group)
            radii_all = _all_gather_rows(pad_rows(radii_c[:keep], c_max), group)
            keys_all = _all_gather_rows(pad_rows(keys_c[:keep], c_max, -1), group)
        with _phase("fwd window (sort+bin+blend)"):
            color, others, state = backend.window_forward(settings, rec_all, radii_all, keys_all, rank, world)
        with _phase("fwd image all-reduce"):
            planes = torch.cat([color, others], 0)
            dist.all_reduce(planes, group=group)               # every rank wrote only its tile rows
        ctx.settings, ctx.backend, ctx.group, ctx.state = settings, backend, group, state
        ctx.num_rendered = state[0] if isinstance(state, tuple) else None
        ctx.save_for_backward(means3D, shs, scales, rotations, radii, rec, clamped, rec_all, slot)
        ctx.mark_non_differentiable(radii)
        return planes[:3].contiguous(), radii, planes[3:].contiguous()

    @staticmethod
    def backward(ctx, g_color, g_radii, g_others):
        means3D, shs, scales, rotations, radii, rec, clamped, rec_all, slot = ctx.saved_tensors
        s, backend, group = ctx.settings, ctx.backend, ctx.group
        with _phase("bwd window blend"):
            grec_all = backend.window_backward(s, rec_all, ctx.state, g_color.contiguous(), g_others.contiguous())
        with _phase("bwd reduce-scatter"):
            grec = _reduce_scatter_rows(grec_all, group)
        with _phase("bwd per-Gaussian"):
            g = backend.shard_backward(s, means3D, shs, scales, rotations, radii, rec, clamped, grec, slot)
        return g["means3D"], g["means2D"], g["shs"], g["opacities"], g["scales"], g["rotations"], None, None, None


class ShardedRasterizer:
    """Callable with the argument order bench.py / tests use: each rank passes ITS shard of the scene.

    ``rasterizer(means3D, means2D, opacities, shs, scales, rotations, settings) -> (color, radii, allmap)``
    where color / allmap are the full images (identical on every rank) and radii belongs to the shard.
    """

    def __init__(self, world: Optional[int] = None, rank: Optional[int] = None, backend=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.backend = backend if backend is not None else NativeBackend()

    def __call__(self, means3D, means2D, opacities, shs, scales, rotations, settings):
        return _ShardedRasterize.apply(means3D, means2D, shs, opacities, scales, rotations, settings, self.backend,
                                       self.group)
