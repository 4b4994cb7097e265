"""GPU tests of the multi-GPU exchange kernels (csrc/exchange.cu) on ONE device.

G "virtual ranks" share the GPU: every CUDA stage of streetunveiler_b200/sharded.py runs per rank through the C ABI
(NativeBackend), and the collectives (all-reduce of the tile histogram, all-gather of the counts, the two all-to-alls,
the image all-reduce) are replaced by the tensor operations they are defined as.  The result must equal the
single-GPU operator: forward bit-identical, gradients within 1e-4 -- for 2, 3 and 8 ranks, uneven shards, an empty
shard.  The torch.distributed side of the same choreography is covered by tests/test_sharded_gloo.py (CPU) and
tests/multigpu_check.py (real ranks)."""
import numpy as np
import pytest
import torch

import harness as hz
from streetunveiler_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


class _LocalPeerRows:
    """PeerRows look-alike whose "peers" are plain tensors on this GPU (one receive / return buffer per virtual rank)."""

    def __init__(self, G, cap_in, cap_out, dev):
        from streetunveiler_b200.sharded import REC_FLOATS, GREC_FLOATS
        self.cap_in, self.cap_out, self.RF, self.GF = cap_in, cap_out, REC_FLOATS, GREC_FLOATS
        self.bufs_in = [torch.full((cap_in * (REC_FLOATS + 2),), float("nan"), device=dev) for _ in range(G)]
        self.bufs_out = [torch.full((cap_out * GREC_FLOATS,), float("nan"), device=dev) for _ in range(G)]

    def in_ptrs(self):
        base = [b.data_ptr() for b in self.bufs_in]
        c = self.cap_in
        return base, [b + c * self.RF * 4 for b in base], [b + c * (self.RF + 1) * 4 for b in base]

    def out_ptrs(self):
        return [b.data_ptr() for b in self.bufs_out]

    def received(self, d, n):
        c, b = self.cap_in, self.bufs_in[d]
        return (b[:n * self.RF].view(n, self.RF).clone(), b[c * (self.RF + 1):c * (self.RF + 1) + n].view(torch.int32).clone(),
                b[c * self.RF:c * self.RF + n].view(torch.int32).clone())

    def returned(self, r, n):
        return self.bufs_out[r][:n * self.GF].view(n, self.GF)


def _virtual_sharded(scene, cam, bounds, grads, bg, use_peers=False):
    from streetunveiler_b200.sharded import NativeBackend
    be = NativeBackend()
    dev = torch.device("cuda")
    G = len(bounds) - 1
    s = hz._settings(hz.ours_module(), cam, bg, int(scene["sh_degree"]), 1.0, dev)
    sh = [{k: v[bounds[r]:bounds[r + 1]].to(dev).contiguous() for k, v in scene.items() if isinstance(v, torch.Tensor)}
          for r in range(G)]
    pre = [be.shard_preprocess(s, p["means3D"], p["shs"], p["opacities"], p["scales"], p["rotations"]) for p in sh]
    hist = sum(be.tile_hist(s, rec, radii).to(torch.int64) for radii, rec, keys, clamped in pre).to(torch.int32)
    cuts, window_R = be.partition(s, hist, G)
    cuts_h, wr_h = cuts.cpu().tolist(), window_R.cpu().tolist()
    routes = [be.route_count(s, rec, radii, cuts, G) for radii, rec, keys, clamped in pre]
    cmat = torch.stack([c[:G] for _, c in routes]).cpu()                     # [src, dst]
    pr = None
    if use_peers:   # the fused exchange: every rank's routing kernel stores into every rank's receive arrays
        pr = _LocalPeerRows(G, int(cmat.sum(0).max()) + 64, int(cmat.sum(1).max()) + 64, dev)
        sends = [(None, be.route_scatter_peers(rec, radii, keys, routes[r][0], routes[r][1], int(cmat[r].sum()), G, pr,
                                               [int(cmat[:r, d].sum()) for d in range(G)]))
                 for r, (radii, rec, keys, clamped) in enumerate(pre)]
    else:
        sends = [be.route_scatter(rec, radii, keys, routes[r][0], routes[r][1], int(cmat[r].sum()), G)
                 for r, (radii, rec, keys, clamped) in enumerate(pre)]

    def seg(r, d):   # rows of rank r's send buffer bound for rank d
        a = int(cmat[r, :d].sum())
        return slice(a, a + int(cmat[r, d]))

    planes, states, recs_w = torch.zeros(10, cam.height, cam.width, device=dev), [], []
    for d in range(G):
        if use_peers:
            rec_w, radii_w, keys_w = pr.received(d, int(cmat[:, d].sum()))
        else:
            rows = torch.cat([sends[r][0][seg(r, d)] for r in range(G)], 0)  # all-to-all: segments in source-rank order
            rec_w, radii_w, keys_w = be.unpack(rows)
        pl, st = be.window_forward(s, rec_w, radii_w, keys_w, cuts_h[d], cuts_h[d + 1], int(wr_h[d]))
        planes += pl                                                          # image all-reduce
        states.append(st)
        recs_w.append(rec_w)
    gc, ga = grads[0].to(dev), grads[1].to(dev)
    grows = [be.window_backward(s, recs_w[d], states[d], gc, ga) for d in range(G)]
    out = {"color": planes[:3].cpu().numpy(), "allmap": planes[3:].cpu().numpy(), "cuts": cuts_h, "window_R": wr_h,
           "cmat": cmat, "radii": torch.cat([p[0] for p in pre]).cpu().numpy()}
    if use_peers:   # every rank pushes the gradient rows of what it received back into the owners' return buffers
        for d in range(G):
            be.push_grad_rows(grows[d], [int(cmat[r, d]) for r in range(G)], G, pr,
                              [int(cmat[r, :d].sum()) for r in range(G)])
    g_all = {}
    for r, (radii, rec, keys, clamped) in enumerate(pre):
        def rseg(d):   # rows of rank d's receive buffer that came from rank r
            a = int(cmat[:r, d].sum())
            return slice(a, a + int(cmat[r, d]))
        back = torch.cat([grows[d][rseg(d)] for d in range(G)], 0)            # all-to-all back
        if use_peers:
            assert torch.equal(pr.returned(r, back.shape[0]), back)           # the push kernel is that all-to-all
            back = pr.returned(r, back.shape[0])
        gacc = be.grad_accumulate(rec.shape[0], back, sends[r][1])
        p = sh[r]
        g = be.shard_backward(s, p["means3D"], p["shs"], p["scales"], p["rotations"], radii, rec, clamped, gacc, None)
        for k, v in g.items():
            g_all.setdefault(k, []).append(v)
    for k, v in g_all.items():
        out["g_" + k] = torch.cat(v, 0).cpu().numpy()
    return out


@pytest.mark.parametrize("G,P,mode,use_peers", [(2, 120_000, "all", False), (3, 120_000, "color_alpha", False),
                                                (8, 400_000, "all", False), (3, 120_000, "all", True),
                                                (8, 400_000, "color_alpha", True)])
def test_virtual_ranks_equal_single_gpu(G, P, mode, use_peers):
    cam = syn.make_camera(960, 640, 1027.5, 1027.5)
    scene = syn.street_scene(P, 4, 3)
    grads = syn.upstream_grads(cam.width, cam.height, mode, seed=11)
    bg = torch.tensor([0.1, 0.3, 0.2])
    # uneven shards, and (G = 3) an empty one
    bounds = [0] + [int(P * (r + 1) / G) - (7 * (r + 1) if r + 1 < G else 0) for r in range(G)]
    if G == 3:
        bounds[2] = bounds[1]
    v = _virtual_sharded(scene, cam, bounds, grads, bg, use_peers)
    ref = hz.run_ours(scene, cam, bg=bg, grads=grads)
    # the image all-reduce adds each rank's planes to zeros: exact; the background term T * bg is part of every tile
    assert np.array_equal(v["color"], ref["color"]) and np.array_equal(v["allmap"], ref["allmap"])
    assert np.array_equal(v["radii"], ref["radii"])
    assert sum(v["window_R"]) == ref["num_rendered"]
    assert v["cuts"][0] == 0 and v["cuts"][-1] == 60 * 40 and all(a <= b for a, b in zip(v["cuts"], v["cuts"][1:]))
    # cost balance: no range holds more than 1.6x its share of the instances (+ the constant per tile)
    cost = [r + 512 * (b - a) for r, a, b in zip(v["window_R"], v["cuts"], v["cuts"][1:])]
    assert max(cost) <= 1.6 * sum(cost) / G, (cost, v["cuts"])
    # a record goes to few ranks: the exchange moves less than 1.7 rows per visible Gaussian (all-gather moves G)
    visible = int((ref["radii"] > 0).sum())
    assert visible <= int(v["cmat"].sum()) <= 1.7 * visible, (int(v["cmat"].sum()), visible)
    for k in ("means3D", "means2D", "shs", "opacities", "scales", "rotations"):
        e = hz.rel_err(v["g_" + k], ref["g_" + k])
        assert e <= 1e-4, (k, e)


def test_partition_matches_host_arithmetic():
    """surfel_shard_partition against the same integer arithmetic in numpy, incl. empty tiles and G > busy tiles."""
    from streetunveiler_b200.sharded import NativeBackend, COST_BASE
    be = NativeBackend()
    cam = syn.make_camera(1920, 1280, 2055.0, 2055.0)
    s = hz._settings(hz.ours_module(), cam, torch.zeros(3), 0, 1.0, torch.device("cuda"))
    n = 120 * 80
    g = torch.Generator().manual_seed(3)
    for case, G, shares in [("random", 8, None), ("spike", 5, None), ("empty", 4, None), ("random", 16, None),
                            ("random", 1, None), ("random", 4, [0.1, 0.4, 0.2, 0.3]), ("random", 3, [0.0, 1.0, 1.0])]:
        if case == "random":
            hist = torch.randint(0, 3000, (n,), generator=g)
        elif case == "spike":
            hist = torch.zeros(n, dtype=torch.int64)
            hist[5000] = 10_000_000
        else:
            hist = torch.zeros(n, dtype=torch.int64)
        cuts, wr = be.partition(s, hist.to(torch.int32).cuda(), G, shares)
        cost = hist.numpy().astype(np.int64) + COST_BASE
        excl = np.cumsum(cost) - cost
        total = int(cost.sum())
        sh = [1.0] * G if shares is None else [np.float32(x).item() for x in shares]
        acc, bounds = 0.0, []
        for k in range(G):                       # the launcher's fixed-point targets and the kernel's ceil(total * t / 2^32)
            t = 0 if k == 0 else int(acc / sum(sh) * 4294967296.0 + 0.5)
            bounds.append((total * t + (1 << 32) - 1) >> 32)
            acc += sh[k]
        rank_of = np.array([sum(1 for k in range(1, G) if int(e) >= bounds[k]) for e in excl])
        want_cuts = [int((rank_of < k).sum()) for k in range(G)] + [n]
        want_R = [int(hist.numpy()[want_cuts[k]:want_cuts[k + 1]].sum()) for k in range(G)]
        assert cuts.cpu().tolist() == want_cuts, (case, G, shares)
        assert wr.cpu().tolist() == want_R, (case, G, shares)


def test_grad_accumulate_sums_rows_of_shared_gaussians():
    from streetunveiler_b200.sharded import NativeBackend, GREC_FLOATS
    be = NativeBackend()
    g = torch.Generator().manual_seed(1)
    P, n = 1000, 1700
    src = torch.randint(0, P, (n,), generator=g)
    rows = torch.randn(n, GREC_FLOATS, generator=g)
    got = be.grad_accumulate(P, rows.cuda(), src.to(torch.int32).cuda()).cpu()
    want = torch.zeros(P, GREC_FLOATS).index_add_(0, src, rows)
    assert torch.allclose(got, want, atol=1e-5)
    assert float(be.grad_accumulate(P, rows[:0].cuda(), src[:0].to(torch.int32).cuda()).abs().max()) == 0.0
