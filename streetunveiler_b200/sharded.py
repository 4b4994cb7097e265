"""Multi-GPU surfel rasterization: Gaussians sharded by index, one process per GPU (torch.distributed).

Why not "every rank blends its own Gaussians and the images are summed": front-to-back alpha blending
is an ordered product, so a sum of per-shard images is not the reference image.  The
exact scheme used here keeps parameters, gradients and optimiser state sharded by Gaussian index,
partitions the SCREEN, and exchanges only *projected* data -- and only with the ranks that need it:

  forward   1. every rank projects its own shard                        (surfel_shard_preprocess, 96-B records)
            2. per-tile instance histogram of the shard, all-reduce(sum) -> the frame's global histogram;
               every rank cuts the row-major tile order into G contiguous ranges of equal cost
               (surfel_shard_tile_hist / surfel_shard_partition; cost = instances + a constant per tile).
               The same kernel yields every range's exact instance count, so no num_rendered read-back
               follows later.
            3. per-destination send counts (a record goes to every rank whose range its tile rect touches:
               1.1-1.3 ranks on average instead of all G), all-gather of the G x G count matrix -- the ONE
               host synchronisation of the forward (sizes of the exchange buffers)
            4. stable per-destination scatter with the exchange fused in: the routing kernel stores every
               record straight into the receive arrays of its destination ranks over NVLink (symmetric
               memory, class PeerRows), at row offsets derived from the count matrix; one barrier.
               Fallback (no peer mappings, CPU/gloo): 112-B rows + NCCL all-to-all + unpack.  Either way
               segments arrive in rank order and keep index order inside, i.e. global index order, so the
               receiver's stable depth sort breaks ties exactly like one GPU does.
            5. every rank depth-sorts, bins and blends what it received for its own tiles
               (surfel_window_unpack / _prepare / _render): same lists, order and arithmetic per pixel as
               on one GPU
            6. image exchange fused into the blend: every rank's kernel stores the ten planes of its tiles
               straight into symmetric buffers of ALL ranks (NVSwitch multicast `multimem.st`, or one store
               per peer over NVLink), then one cross-rank barrier (class PeerImages).  Fallback where no
               peer mapping exists (and on CPU/gloo): all-reduce(sum) of the planes -- each pixel has
               exactly one non-zero summand, so the sum is exact
  backward  7. every rank back-propagates its tiles into 80-B gradient rows of the records it received
                                                                     (surfel_window_backward)
            8. the gradient rows go back along the same routes (pushed straight into the owners' symmetric
               return buffers + barrier, or NCCL all-to-all); the owner sums the rows of a Gaussian that
               went to several ranks (surfel_shard_grad_accumulate)
            9. every rank turns its accumulator into parameter gradients (surfel_shard_backward)

Contract: the upstream gradients (dL/dcolor, dL/dallmap) must be THE SAME on every rank -- each rank
back-propagates only its own tiles and trusts its local copy for them (a loss computed identically on
the all-reduced image satisfies this).  ``ShardedRasterizer(check_replicated=True)`` verifies it with
one all-reduce of a checksum per backward.

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
XROW_FLOATS = 28     # exchanged row: record + depth key + radius + 2 pad words (csrc/common.cuh)
TILE = 16
COST_BASE = 512      # per-tile constant of the partition's cost model, in units of one tile instance (prior; the feedback of _Balancer does the rest)

# SURFEL_SHARD_TIMING=1: synchronise around every phase and record wall-clock per phase and call (diagnostics only)
import os as _os
import time as _time
_TIMING = _os.environ.get("SURFEL_SHARD_TIMING") == "1"
PHASE_MS = {}
LAST_INFO = {}       # diagnostics of the most recent forward on this rank (cuts, counts, routed rows)


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
# collectives (gloo is used only by the CPU tests)
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


def _all_to_all_rows(x: torch.Tensor, in_splits, out_splits, group) -> torch.Tensor:
    """Rows [sum(in_splits), ...] split by destination -> rows [sum(out_splits), ...] concatenated by source."""
    out = x.new_empty((int(sum(out_splits)),) + tuple(x.shape[1:]))
    dist.all_to_all_single(out, x.contiguous(), output_split_sizes=list(out_splits), input_split_sizes=list(in_splits),
                           group=group)
    return out


def pad_rows(x: torch.Tensor, n: int, value=0) -> torch.Tensor:
    """Pad the first dimension to n rows."""
    if x.shape[0] == n:
        return x.contiguous()
    pad = x.new_full((n - x.shape[0],) + tuple(x.shape[1:]), value)
    return torch.cat([x, pad], 0).contiguous()


# ----------------------------------------------------------------------------------------------------
# image exchange fused into the blend kernel: symmetric [10,H,W] buffers that every rank's kernel stores into
# ----------------------------------------------------------------------------------------------------
IMAGE_EXCHANGE = _os.environ.get("SURFEL_IMAGE_EXCHANGE", "auto")   # auto | multicast | peers | allreduce


class PeerImages:
    """Two symmetric-memory image buffers per resolution (torch.distributed._symmetric_memory: CUDA VMM allocations
    mapped into every rank of the node, plus an NVSwitch multicast mapping when the fabric has one).  The forward
    blend kernel of every rank stores the ten planes of ITS tiles into all ranks' buffers -- one ``multimem.st`` per
    value to the multicast address, or one store per peer -- so the image exchange overlaps with the blend and the
    all-reduce of ten mostly-zero planes (98 MB at 1920x1280, ring: 2 x 7/8 of it per rank) disappears.  ``finish``
    = cross-rank barrier (every store has landed) + copy-out into a private tensor.  Buffers alternate between frames:
    a rank that runs ahead never writes into a buffer a slower rank may still be copying out of."""

    def __init__(self, H, W, device, group, mode):
        import torch.distributed._symmetric_memory as symm
        self.H, self.W, self.frame = H, W, 0
        g = group if group is not None else dist.group.WORLD
        self.slots = []
        for _ in range(2):
            t = symm.empty((10 * H * W,), dtype=torch.float32, device=device)
            hdl = symm.rendezvous(t, g)
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if mode == "peers" or (mode == "auto" and not mc):
                ptrs, multicast = [int(x) for x in hdl.buffer_ptrs], 0
            elif mc:
                ptrs, multicast = [mc], 1
            else:
                raise RuntimeError("no NVSwitch multicast mapping for the symmetric buffer")
            self.slots.append((t, hdl, ptrs, multicast))
        self.mode = "multicast" if self.slots[0][3] else "peers"

    def begin(self):
        self.frame += 1
        return self.slots[self.frame & 1]

    def finish(self, slot):
        t, hdl, _, _ = slot
        hdl.barrier(channel=0)
        return t.view(10, self.H, self.W).clone()


ROW_EXCHANGE = _os.environ.get("SURFEL_ROW_EXCHANGE", "auto")       # auto | peers | nccl


class PeerRows:
    """Symmetric receive buffers for the two row exchanges (records forward, gradient rows backward).  With them the
    routing kernel stores every record straight into its destination ranks' arrays over NVLink, and the backward pushes
    gradient rows straight back into their owners' buffers: no send buffers, no NCCL all-to-all, no unpack pass; one
    cross-rank barrier each way.  Capacities grow collectively: every rank knows the whole count matrix after the
    forward's host synchronisation, so all ranks take the same decision."""

    def __init__(self, device, group):
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.cap_in = self.cap_out = 0
        self.buf_in = self.hdl_in = self.buf_out = self.hdl_out = None

    @staticmethod
    def _round(n):
        # generous: re-allocating is a collective rendezvous (tens of ms), and the balancer may double a rank's share
        # of the screen after the first frames; HBM is not the constraint (2M rows = 208 MB)
        n = int(n * 2.5) + 1
        return -(-n // 65536) * 65536

    def ensure(self, need_in, need_out):
        import torch.distributed._symmetric_memory as symm
        if need_in > self.cap_in:
            self.cap_in = self._round(need_in)
            self.buf_in = symm.empty((self.cap_in * (REC_FLOATS + 2),), dtype=torch.float32, device=self.device)
            self.hdl_in = symm.rendezvous(self.buf_in, self.group)
        if need_out > self.cap_out:
            self.cap_out = self._round(need_out)
            self.buf_out = symm.empty((self.cap_out * GREC_FLOATS,), dtype=torch.float32, device=self.device)
            self.hdl_out = symm.rendezvous(self.buf_out, self.group)

    def in_ptrs(self):
        """Per rank: addresses of its records / depth keys / radii receive arrays."""
        base = [int(x) for x in self.hdl_in.buffer_ptrs]
        c = self.cap_in
        return base, [b + c * REC_FLOATS * 4 for b in base], [b + c * (REC_FLOATS + 1) * 4 for b in base]

    def out_ptrs(self):
        return [int(x) for x in self.hdl_out.buffer_ptrs]

    def take_received(self, n):
        """Barrier (every rank's stores have landed), then private copies of the n received rows: the records are
        needed again by the backward, by which time a later forward may have refilled the symmetric buffer."""
        self.hdl_in.barrier(channel=0)
        c = self.cap_in
        rec = self.buf_in[:n * REC_FLOATS].view(n, REC_FLOATS).clone()
        keys = self.buf_in[c * REC_FLOATS:c * REC_FLOATS + n].view(torch.int32).clone()
        radii = self.buf_in[c * (REC_FLOATS + 1):c * (REC_FLOATS + 1) + n].view(torch.int32).clone()
        return rec, radii, keys

    def returned(self, n):
        self.hdl_out.barrier(channel=0)
        return self.buf_out[:n * GREC_FLOATS].view(n, GREC_FLOATS)


# ----------------------------------------------------------------------------------------------------
# native backend: the CUDA library through its C ABI
# ----------------------------------------------------------------------------------------------------
class NativeBackend:
    def __init__(self):
        from . import _lib
        self._lib = _lib
        self.L = _lib.lib()
        self._peer_images = {}
        self.image_exchange = None      # "multicast" | "peers" | "allreduce" once the first frame ran
        self._peer_rows = {}
        self.row_exchange = None        # "peers" | "nccl" once the first frame ran

    def peer_rows(self, device, group, need_in, need_out):
        """Symmetric row buffers with room for this frame, or None -> NCCL all-to-all (SURFEL_ROW_EXCHANGE=nccl, a
        world of one, or a node without peer mappings)."""
        if ROW_EXCHANGE == "nccl" or dist.get_world_size(group) == 1:
            self.row_exchange = "nccl"
            return None
        key = device.index
        if key not in self._peer_rows:
            try:
                pr = PeerRows(device, group)
                pr.ensure(need_in, need_out)
                self._peer_rows[key] = pr
            except Exception as exc:   # all ranks fail alike (same node, same software): a collective decision
                import warnings
                warnings.warn(f"symmetric-memory row exchange unavailable ({type(exc).__name__}: {exc}); "
                              f"using NCCL all-to-all")
                self._peer_rows[key] = None
        pr = self._peer_rows[key]
        if pr is not None:
            pr.ensure(need_in, need_out)
        self.row_exchange = "peers" if pr is not None else "nccl"
        return pr

    def route_scatter_peers(self, rec, radii, keys, route_state, send_counts_dev, n_send, world, pr, row0):
        src = torch.empty((n_send,), dtype=torch.int32, device=rec.device)
        prec, pkeys, pradii = pr.in_ptrs()
        A = C.c_void_p * world
        self._lib.check(self.L.surfel_shard_route_scatter_peers(
            rec.shape[0], world, self._p(rec), self._p(radii), self._p(keys), C.c_void_p(route_state.data_ptr()),
            C.c_void_p(send_counts_dev.data_ptr()), A(*prec), A(*pkeys), A(*pradii), (C.c_int64 * world)(*row0),
            C.c_void_p(src.data_ptr()), self._stream()), "surfel_shard_route_scatter_peers")
        return src

    def push_grad_rows(self, grows, seg_count, world, pr, row0):
        self._lib.check(self.L.surfel_window_push_grad_rows(
            grows.shape[0], self._p(grows), world, (C.c_int64 * world)(*seg_count), (C.c_void_p * world)(*pr.out_ptrs()),
            (C.c_int64 * world)(*row0), self._stream()), "surfel_window_push_grad_rows")

    def peer_images(self, s, device, group):
        """Symmetric image buffers for this resolution, or None -> the caller all-reduces the planes instead
        (SURFEL_IMAGE_EXCHANGE=allreduce, a world of one, or a node without peer mappings)."""
        if IMAGE_EXCHANGE == "allreduce" or dist.get_world_size(group) == 1:
            self.image_exchange = "allreduce"
            return None
        key = (s.image_height, s.image_width, device.index)
        if key not in self._peer_images:
            try:
                self._peer_images[key] = PeerImages(s.image_height, s.image_width, device, group, IMAGE_EXCHANGE)
            except Exception as exc:   # all ranks fail alike (same node, same software): a collective decision
                import warnings
                warnings.warn(f"symmetric-memory image exchange unavailable ({type(exc).__name__}: {exc}); "
                              f"using the NCCL all-reduce of the image planes")
                self._peer_images[key] = None
        pi = self._peer_images[key]
        self.image_exchange = pi.mode if pi is not None else "allreduce"
        return pi

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
        """Round-1 exchange helper, kept for callers of the all-gather scheme.
        -> (rec_c, radii_c, keys_c [P rows of capacity each], slot [P], count [1] int32 on the device)."""
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

    def tile_hist(self, s, rec, radii):
        W, H, dev = s.image_width, s.image_height, rec.device
        tiles = ((W + TILE - 1) // TILE) * ((H + TILE - 1) // TILE)
        hist = torch.empty((tiles,), dtype=torch.int32, device=dev)
        self._lib.check(self.L.surfel_shard_tile_hist(rec.shape[0], W, H, self._p(rec), self._p(radii),
                                                      C.c_void_p(hist.data_ptr()), self._stream()), "surfel_shard_tile_hist")
        return hist

    def partition(self, s, hist, world, shares=None, cost_base=None):
        """-> (cuts int32[G + 1] tile ids, window_R int64[G] instance count of every range), both on the device.
        shares: G positive floats, the fraction of the frame's cost every rank should receive (None: equal)."""
        W, H, dev = s.image_width, s.image_height, hist.device
        tmp = torch.empty((self._lib.size(self.L.surfel_shard_partition_bytes(W, H), "surfel_shard_partition_bytes"),),
                          dtype=torch.uint8, device=dev)
        cuts = torch.empty((world + 1,), dtype=torch.int32, device=dev)
        wr = torch.empty((world,), dtype=torch.int64, device=dev)
        sh = None if shares is None else (C.c_float * world)(*[float(x) for x in shares])
        self._lib.check(self.L.surfel_shard_partition(W, H, world, C.c_void_p(hist.data_ptr()),
                                                      COST_BASE if cost_base is None else int(cost_base), sh,
                                                      C.c_void_p(tmp.data_ptr()), C.c_void_p(cuts.data_ptr()),
                                                      C.c_void_p(wr.data_ptr()), self._stream()), "surfel_shard_partition")
        return cuts, wr

    def route_count(self, s, rec, radii, cuts, world, extra=0):
        """-> (routing state, int32[G + 1] on the device: send counts per destination, then `extra`)."""
        P, dev = rec.shape[0], rec.device
        tmp = torch.empty((self._lib.size(self.L.surfel_shard_route_bytes(P, world), "surfel_shard_route_bytes"),),
                          dtype=torch.uint8, device=dev)
        counts = torch.empty((world + 1,), dtype=torch.int32, device=dev)
        self._lib.check(self.L.surfel_shard_route_count(P, s.image_width, s.image_height, world, self._p(rec),
                                                        self._p(radii), C.c_void_p(cuts.data_ptr()),
                                                        C.c_void_p(tmp.data_ptr()), C.c_void_p(counts.data_ptr()),
                                                        int(extra), self._stream()), "surfel_shard_route_count")
        return tmp, counts

    def route_scatter(self, rec, radii, keys, route_state, send_counts_dev, n_send, world):
        dev = rec.device
        rows = torch.empty((n_send, XROW_FLOATS), dtype=torch.float32, device=dev)
        src = torch.empty((n_send,), dtype=torch.int32, device=dev)
        self._lib.check(self.L.surfel_shard_route_scatter(rec.shape[0], world, self._p(rec), self._p(radii), self._p(keys),
                                                          C.c_void_p(route_state.data_ptr()),
                                                          C.c_void_p(send_counts_dev.data_ptr()), self._p(rows),
                                                          self._p(src), self._stream()), "surfel_shard_route_scatter")
        return rows, src

    def unpack(self, rows):
        n, dev = rows.shape[0], rows.device
        rec = torch.empty((n, REC_FLOATS), dtype=torch.float32, device=dev)
        radii = torch.empty((n,), dtype=torch.int32, device=dev)
        keys = torch.empty((n,), dtype=torch.int32, device=dev)
        self._lib.check(self.L.surfel_window_unpack(n, self._p(rows), self._p(rec), self._p(radii), self._p(keys),
                                                    self._stream()), "surfel_window_unpack")
        return rec, radii, keys

    def window_forward(self, s, rec_w, radii_w, keys_w, tile_lo, tile_hi, R, peer_slot=None):
        """Blend the received Gaussians for the tiles [tile_lo, tile_hi); R = the window's instance count (known from
        the partition).  Without peer_slot: returns [10,H,W] planes, zero outside the window (for the all-reduce).
        With peer_slot (PeerImages.begin()): the kernel stores the window's pixels into every rank's buffer; returns
        None for the planes."""
        L, p, dev = self.L, self._p, rec_w.device
        n, W, H = rec_w.shape[0], s.image_width, s.image_height
        u8 = dict(dtype=torch.uint8, device=dev)
        win = torch.empty((self._lib.size(L.surfel_window_bytes(n), "surfel_window_bytes"),), **u8)
        img = torch.empty((self._lib.size(L.surfel_image_bytes(W, H), "surfel_image_bytes"),), **u8)
        st = self._stream()
        self._lib.check(L.surfel_window_prepare(n, W, H, tile_lo, tile_hi, p(rec_w), p(radii_w), p(keys_w), p(win), None,
                                                st, int(bool(s.debug))), "surfel_window_prepare")
        binb = torch.empty((self._lib.size(L.surfel_binning_bytes(R), "surfel_binning_bytes") if R else 0,), **u8)
        if peer_slot is not None:
            _, _, ptrs, multicast = peer_slot
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            self._lib.check(L.surfel_window_render_peers(n, W, H, tile_lo, tile_hi, R, p(s.bg), p(rec_w), p(radii_w), p(win),
                                                         p(binb), p(img), len(ptrs), arr, int(multicast), st,
                                                         int(bool(s.debug))), "surfel_window_render_peers")
            return None, (R, binb, img, tile_lo, tile_hi)
        planes = torch.zeros((10, H, W), dtype=torch.float32, device=dev)
        color, others = planes[:3], planes[3:]
        self._lib.check(L.surfel_window_render(n, W, H, tile_lo, tile_hi, R, p(s.bg), p(rec_w), p(radii_w), p(win), p(binb),
                                               p(img), p(color), p(others), st, int(bool(s.debug))), "surfel_window_render")
        return planes, (R, binb, img, tile_lo, tile_hi)

    def window_backward(self, s, rec_w, state, dL_dcolor, dL_dothers):
        R, binb, img, tile_lo, tile_hi = state
        n = rec_w.shape[0]
        grows = torch.empty((n, GREC_FLOATS), dtype=torch.float32, device=dL_dcolor.device)
        p = self._p
        if n:
            self._lib.check(self.L.surfel_window_backward(
                n, s.image_width, s.image_height, tile_lo, tile_hi, R, p(s.bg), p(rec_w), p(binb), p(img),
                p(dL_dcolor.contiguous()), p(dL_dothers.contiguous()), p(grows), self._stream(), int(bool(s.debug))),
                "surfel_window_backward")
        return grows

    def grad_accumulate(self, P, rows, send_src):
        gacc = torch.empty((P, GREC_FLOATS), dtype=torch.float32, device=rows.device)
        self._lib.check(self.L.surfel_shard_grad_accumulate(P, rows.shape[0], self._p(rows), self._p(send_src),
                                                            self._p(gacc), self._stream()), "surfel_shard_grad_accumulate")
        return gacc

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
class _Balancer:
    """Feedback for the screen partition.  The blend time of a tile range is not proportional to its instance count
    (sparse regions cost more per instance than saturated ones, and a range cannot finish before its longest tile
    list), so every rank measures its own window time (forward sort+bin+blend and backward blend, CUDA events) and the
    shares of the cost histogram that the ranks receive are corrected multiplicatively:
        share_k <- share_k * (mean time / time_k) ** gain, normalised.
    The times travel with the per-step all-gather of the send counts, so every rank sees the same numbers and derives
    the same shares; a measurement is applied two steps later (its events are then complete on every rank without an
    extra synchronisation) to the shares it was measured under.  Only WHO blends a tile depends on this, never a result.
    """

    def __init__(self, world: int, gain: float = 0.7, floor: float = 0.02):
        self.world, self.gain, self.floor = world, gain, floor
        self.shares = [1.0 / world] * world
        self.last_times_us = None      # the ranks' window times (fwd + bwd, microseconds) of the latest complete measurement
        self.step = 0
        self.records = {}     # step -> {"shares": [...], "ev": [e0, e1, e2, e3]}

    def begin_step(self):
        self.step += 1
        self.records[self.step] = {"shares": list(self.shares), "ev": [None] * 4}
        for k in [k for k in self.records if k < self.step - 4]:
            del self.records[k]
        return self.step

    def mark(self, step, i):
        rec = self.records.get(step)
        if rec is not None and torch.cuda.is_available():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            rec["ev"][i] = e

    def report(self):
        """(step, microseconds) of this rank's measurement that is two steps old (0 us if there is none)."""
        k = self.step - 2
        rec = self.records.get(k)
        if rec is None or rec["ev"][0] is None or rec["ev"][1] is None:
            return k, 0
        try:
            ms = rec["ev"][0].elapsed_time(rec["ev"][1])
            if rec["ev"][2] is not None and rec["ev"][3] is not None:
                ms += rec["ev"][2].elapsed_time(rec["ev"][3])
        except RuntimeError:      # an event that never completed (e.g. backward skipped mid-way)
            return k, 0
        return k, max(int(ms * 1000.0), 1)

    def update(self, step, times_us):
        rec = self.records.get(step)
        if rec is None or any(t <= 0 for t in times_us):
            return
        self.last_times_us = [int(t) for t in times_us]
        mean = sum(times_us) / len(times_us)
        new = [s * (mean / t) ** self.gain for s, t in zip(rec["shares"], times_us)]
        tot = sum(new)
        new = [max(x / tot, self.floor / self.world) for x in new]
        tot = sum(new)
        self.shares = [x / tot for x in new]


class _ShardedRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, opacities, scales, rotations, settings, backend, group, check_replicated,
                balancer):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        P = means3D.shape[0]
        step_id = balancer.begin_step() if balancer is not None else 0
        with _phase("fwd preprocess"):
            radii, rec, keys, clamped = backend.shard_preprocess(settings, means3D.contiguous(), shs.contiguous(),
                                                                 opacities.contiguous(), scales.contiguous(),
                                                                 rotations.contiguous())
        with _phase("fwd partition (hist all-reduce, cuts)"):
            hist = backend.tile_hist(settings, rec, radii)
            dist.all_reduce(hist, group=group)
            cuts, window_R = backend.partition(settings, hist, world, balancer.shares if balancer is not None else None)
        with _phase("fwd route counts + host sync"):
            rep_step, rep_us = balancer.report() if balancer is not None else (0, 0)
            route_state, send_counts = backend.route_count(settings, rec, radii, cuts, world, min(rep_us, 2 ** 31 - 1))
            gathered = _all_gather_rows(send_counts, group).view(world, world + 1)
            host = torch.cat([gathered.flatten().to(torch.int64), cuts.to(torch.int64), window_R.to(torch.int64)]).cpu()
            gmat = host[:world * (world + 1)].view(world, world + 1)
            cmat = gmat[:, :world]
            cuts_h = host[world * (world + 1):world * (world + 1) + world + 1].tolist()
            R = int(host[world * (world + 1) + world + 1 + rank])
            in_splits = cmat[rank].tolist()
            out_splits = cmat[:, rank].tolist()
            if balancer is not None:
                balancer.update(rep_step, gmat[:, world].tolist())
        with _phase("fwd record exchange (scatter into peers + barrier | scatter + all-to-all + unpack)"):
            n_send, n_recv = int(sum(in_splits)), int(sum(out_splits))
            pr = None
            if hasattr(backend, "peer_rows"):
                pr = backend.peer_rows(rec.device, group, int(cmat.sum(0).max()), int(cmat.sum(1).max()))
            if pr is not None:
                row0 = [int(cmat[:rank, d].sum()) for d in range(world)]      # rows of lower ranks bound for d
                send_src = backend.route_scatter_peers(rec, radii, keys, route_state, send_counts, n_send, world, pr, row0)
                rec_w, radii_w, keys_w = pr.take_received(n_recv)
            else:
                send_rows, send_src = backend.route_scatter(rec, radii, keys, route_state, send_counts, n_send, world)
                recv_rows = _all_to_all_rows(send_rows, in_splits, out_splits, group)
                del send_rows
                rec_w, radii_w, keys_w = backend.unpack(recv_rows)
                del recv_rows
        with _phase("fwd window (sort+bin+blend)"):
            if balancer is not None:
                balancer.mark(step_id, 0)
            peers = backend.peer_images(settings, rec.device, group) if hasattr(backend, "peer_images") else None
            if peers is not None:
                slot = peers.begin()
                _, state = backend.window_forward(settings, rec_w, radii_w, keys_w, int(cuts_h[rank]),
                                                  int(cuts_h[rank + 1]), R, slot)
            else:
                planes, state = backend.window_forward(settings, rec_w, radii_w, keys_w, int(cuts_h[rank]),
                                                       int(cuts_h[rank + 1]), R)
            if balancer is not None:
                balancer.mark(step_id, 1)
        with _phase("fwd image exchange (barrier + copy-out | all-reduce)"):
            if peers is not None:
                planes = peers.finish(slot)                    # every rank's kernel stored its tiles into every buffer
            else:
                dist.all_reduce(planes, group=group)           # every rank wrote only its tiles, the rest is zero
        LAST_INFO.update(cuts=cuts_h, send=in_splits, recv=out_splits, num_rendered=R, shard=P,
                         shares=list(balancer.shares) if balancer is not None else None)
        ctx.settings, ctx.backend, ctx.group, ctx.state = settings, backend, group, state
        ctx.splits = (in_splits, out_splits)
        ctx.peer_rows = pr
        ctx.back_row0 = [int(cmat[src, :rank].sum()) for src in range(world)] if pr is not None else None
        ctx.check_replicated = check_replicated
        ctx.balancer, ctx.step_id = balancer, step_id
        ctx.num_rendered = R
        ctx.save_for_backward(means3D, shs, scales, rotations, radii, rec, clamped, rec_w, send_src)
        ctx.mark_non_differentiable(radii)
        return planes[:3], radii, planes[3:]

    @staticmethod
    def backward(ctx, g_color, g_radii, g_others):
        means3D, shs, scales, rotations, radii, rec, clamped, rec_w, send_src = ctx.saved_tensors
        s, backend, group = ctx.settings, ctx.backend, ctx.group
        in_splits, out_splits = ctx.splits
        if ctx.check_replicated:
            chk = torch.stack([g_color.double().sum(), g_others.nan_to_num(0.0).double().sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if not torch.equal(lo, hi):
                raise RuntimeError("ShardedRasterizer: upstream gradients differ between ranks (replicated-loss contract)")
        with _phase("bwd window blend"):
            if ctx.balancer is not None:
                ctx.balancer.mark(ctx.step_id, 2)
            grows = backend.window_backward(s, rec_w, ctx.state, g_color.contiguous(), g_others.contiguous())
            if ctx.balancer is not None:
                ctx.balancer.mark(ctx.step_id, 3)
        with _phase("bwd gradient rows (push into peers + barrier | all-to-all)"):
            if ctx.peer_rows is not None:
                # my rows for owner `src` start where the rows of lower-ranked receivers end in src's send order
                backend.push_grad_rows(grows, out_splits, len(out_splits), ctx.peer_rows, ctx.back_row0)
                back = ctx.peer_rows.returned(int(sum(in_splits)))
            else:
                back = _all_to_all_rows(grows, out_splits, in_splits, group)
        with _phase("bwd per-Gaussian"):
            gacc = backend.grad_accumulate(means3D.shape[0], back, send_src)
            g = backend.shard_backward(s, means3D, shs, scales, rotations, radii, rec, clamped, gacc, None)
        return (g["means3D"], g["means2D"], g["shs"], g["opacities"], g["scales"], g["rotations"], None, None, None, None,
                None)


class ShardedRasterizer:
    """Callable with the argument order bench.py / tests use: each rank passes ITS shard of the scene.

    ``rasterizer(means3D, means2D, opacities, shs, scales, rotations, settings) -> (color, radii, allmap)``
    where color / allmap are the full images (identical on every rank) and radii belongs to the shard.
    The upstream gradients given to ``backward`` must be identical on every rank (module docstring);
    ``check_replicated=True`` verifies that on every backward.  ``balance=True`` (default) adapts the screen
    partition to the ranks' measured blend times (class _Balancer); results do not depend on it.
    """

    def __init__(self, world: Optional[int] = None, rank: Optional[int] = None, backend=None, group=None,
                 check_replicated: bool = False, balance: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        if self.world > 16:
            raise ValueError("at most 16 ranks per scene (csrc/common.cuh MAX_RANKS)")
        self.backend = backend if backend is not None else NativeBackend()
        self.check_replicated = check_replicated
        self.balancer = _Balancer(self.world) if balance else None

    def __call__(self, means3D, means2D, opacities, shs, scales, rotations, settings):
        return _ShardedRasterize.apply(means3D, means2D, shs, opacities, scales, rotations, settings, self.backend,
                                       self.group, self.check_replicated, self.balancer)
