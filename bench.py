#!/usr/bin/env python
"""bench.py -- fwd+bwd throughput of the surfel rasterizer on synthetic street scenes.

Metric (BASELINE.json): M Gaussians/s fwd+bwd @1920x1280 = P / (t_fwd + t_bwd) / 1e6.
A "step" is one forward + one backward of the rasterizer operator over one synthetic scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cpu]

  * N = 1 (default): BASELINE config 3 -- STREET(P=2,000,000, seed 1), CAM-A 1920x1280, SH degree 3,
    colour + alpha upstream gradients.
  * N > 1 (torchrun, one rank per GPU): weak scaling -- every rank owns 2,000,000 surfels (an index
    shard) of one STREET(2,000,000*N) scene (streetunveiler_b200/sharded.py).  The same invocation then
    (a) checks the sharded result against the unsharded operator on rank 0 (`parity_selfcheck`) and
    (b) times BASELINE configs[4] -- ONE 8,000,000-surfel scene split over the N ranks, strong scaling --
    next to the single-GPU time of the same scene measured on rank 0 (`config5_strong`).
  * --impl reference: the UNMODIFIED reference CUDA extension rebuilt for sm_100a (oracle/_ref), same
    inputs and timing protocol, on the same GPU; falls back to the CPU oracle port when that build
    is absent.  --impl reference-cpu forces the CPU oracle port (host cores).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks; inputs (>460 MB) are larger than the 126 MB L2.
`ms_per_step` / `value` are the K timed steps as one bracket (the driver's contract); `ms_per_step_median`
is the median of the K per-step intervals inside that bracket.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from streetunveiler_b200 import synthetic as syn  # noqa: E402

METRIC = "M Gaussians/s fwd+bwd @1920x1280"
UNIT = "MGaussians/s"
P_PER_GPU = 2_000_000
STRONG_TOTAL = 8_000_000   # BASELINE configs[4]


def count_launches(step_fn, device):
    """Kernel launches of ONE step, counted by CUPTI (torch.profiler): -> (ours, other, {name: n}).  `ours` are the
    hand-written kernels of libsurfel_b200.so (namespace surfel::), `other` everything else on the device in that step
    (PyTorch fills / copies of the harness, NCCL)."""
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize(device)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step_fn()
        torch.cuda.synchronize(device)
    names = {}
    for ev in prof.events():
        if getattr(ev, "device_type", None) is not None and "cuda" in str(ev.device_type).lower():
            n = ev.name
            if n.startswith("Memcpy") or n.startswith("Memset"):
                continue
            names[n] = names.get(n, 0) + 1
    mine = lambda n: "surfel::" in n or "publish_u32_kernel" in n   # noqa: E731  (the latter sits in api.cu's anonymous namespace)
    ours = sum(c for n, c in names.items() if mine(n))
    other = sum(c for n, c in names.items() if not mine(n))
    short = {}
    for n, c in names.items():
        k = n.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")[:60]
        short[k] = short.get(k, 0) + c
    return ours, other, short


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            out["samples"] = len(sm)
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                seen = set()
                for r in rows:
                    for n, v in zip(names, r[5:9]):
                        if v.strip().lower().startswith("active"):
                            seen.add(n)
                out["reasons"] = sorted(seen)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


# ------------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes_step(P, R, HW, K=16):
    """ALG_BYTES = (IN + 4 + IN + OUT) P + 36 R + 80 HW with IN = 40 + 12K, OUT = 52 + 12K."""
    IN = 12 + 8 + 16 + 4 + 12 * K
    OUT = 12 + 12 + 4 + 8 + 16 + 12 * K
    return (IN + 4 + IN + OUT) * P + 36 * R + 80 * HW


def alg_bytes_stage(stage, P, P_vis, R, HW, K=16):
    """Compulsory (cache-perfect) bytes of each stage of OUR pipeline; stated in DESIGN.md section 5."""
    IN = 12 + 8 + 16 + 4 + 12 * K
    OUT = 12 + 12 + 4 + 8 + 16 + 12 * K
    return {
        "preprocess_fwd": P * (40 + 4 + 4 + 4 + 4) + P_vis * (12 * K + 96 + 1),   # params in; radii/tiles/key/iota out; SH in + record out
        "depth_order": P * 8 * 2 * 4 + P * 12,                                    # 4 digit passes of 8 B pairs (r+w) + scan
        "tile_binning": P * 12 + R * 8 + R * 16 * 2 + R * 4,                      # emit 8 B/inst, 2 digit passes r+w, ranges read
        "render_fwd": R * 4 + P_vis * 96 + HW * 60,                               # ids + each record once + 15 planes out
        "render_bwd": R * 4 + P_vis * (96 + 72) + HW * 60,                        # ids + records + grad record out + 15 planes in
        "preprocess_bwd": P * 4 + P_vis * (IN + 96 + 72) + P * OUT,               # radii; params+record+grad record in; grads out
    }[stage]


# ------------------------------------------------------------------------------------------------
class Workload:
    """One rank's share of the synthetic scene, resident on its GPU."""

    def __init__(self, P_total, seed, world, rank, device):
        self.cam = syn.cam_a()
        if world == 1:
            scene = syn.street_scene(P_total, seed, 3)
        else:
            # index shard r of a STREET(P_total) scene: the street distribution is i.i.d. per surfel, so each
            # rank draws its own P_total/world surfels (seed + rank) instead of generating all of them
            scene = syn.street_scene(P_total // world, seed * 1000 + rank, 3)
        self.host = scene
        self.crc = syn.scene_crc(scene)
        self.P = scene["means3D"].shape[0]
        self.device = device
        self.dev = {k: v.to(device) for k, v in scene.items() if isinstance(v, torch.Tensor)}
        dc, da = syn.upstream_grads(self.cam.width, self.cam.height, "color_alpha")
        self.grads_host = (dc, da)
        self.grads = (dc.to(device), da.to(device))
        self.bg = torch.zeros(3, device=device)


def make_step(mod, wl: Workload, sharded=None, own_buffers=False):
    import harness as hz
    st = hz._settings(mod, wl.cam, torch.zeros(3), 3, 1.0, wl.device)
    rast = mod.GaussianRasterizer(st)
    src = {k: (v.clone() if own_buffers else v) for k, v in wl.dev.items()}
    leaves = {k: src[k].detach().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
    grads = (wl.grads[0].clone(), wl.grads[1].clone()) if own_buffers else wl.grads
    state = {"grads": grads}

    def step():
        for t in list(leaves.values()) + [m2]:
            t.grad = None
        if sharded is None:
            color, radii, allmap = rast(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"],
                                        shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        else:
            color, radii, allmap = sharded(leaves["means3D"], m2, leaves["opacities"], leaves["shs"], leaves["scales"],
                                           leaves["rotations"], st)
        torch.autograd.backward([color, allmap], [grads[0], grads[1]])
        state["color"], state["allmap"], state["radii"] = color, allmap, radii
        fn = color.grad_fn
        state["R"] = getattr(fn, "num_rendered", None)
        return state

    return step, leaves, m2, state


def timed_loop(fn, steps, warmup, world, device):
    """-> (ms per step over the whole K-step bracket, median of the K per-step intervals); both max over ranks."""
    for _ in range(warmup):
        fn()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record()
    for i in range(steps):
        fn()
        marks[i + 1].record()
    torch.cuda.synchronize(device)
    ms = marks[0].elapsed_time(marks[steps]) / steps
    med = float(np.median([marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]))
    if world > 1:
        t = torch.tensor([ms, med], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, med = float(t[0].item()), float(t[1].item())
        torch.distributed.barrier()
    return ms, med


def _measure_e2e_once(mod, wl, args, world, device, sharded=None):
    """Same metric through the operator with HOST buffers: every step copies all inputs (parameters and
    upstream gradients) from pinned host memory to the device, runs fwd+bwd, and copies every output and
    gradient back to pinned host memory.  Steps are software-pipelined over two device buffer sets and
    three streams (H2D of step i+1 | compute of step i | D2H of step i-1), PCIe being full duplex; the upload
    of step i+1 is queued before step i's kernels so that it does not wait for the host's num_rendered read-back."""
    pin = {k: v.pin_memory() for k, v in wl.host.items() if isinstance(v, torch.Tensor)}
    pin_g = [g.pin_memory() for g in wl.grads_host]
    h2d = sum(v.numel() * v.element_size() for v in pin.values()) + sum(g.numel() * 4 for g in pin_g)
    sets = [make_step(mod, wl, sharded, own_buffers=True) for _ in range(2)]
    # compute stays on the default stream (the reference extension can only launch there); copies use side streams
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    s_comp = torch.cuda.default_stream(device)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_comp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    out_host = [{}, {}]
    trace = [] if os.environ.get("BENCH_E2E_TRACE") else None   # diagnostics: per-phase device timestamps on stderr
    done = []                                                   # one timing event per step: its results have left the device

    def mark(stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def upload(i):
        """H2D of step i's inputs into buffer set i % 2 (waits until step i-2 has consumed that set)."""
        k = i % 2
        _, leaves, _, state = sets[k]
        t = {}
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_comp[k])
            if trace is not None:
                t["in0"] = mark(s_in)
            for name in leaves:
                leaves[name].data.copy_(pin[name], non_blocking=True)
            state["grads"][0].copy_(pin_g[0], non_blocking=True)
            state["grads"][1].copy_(pin_g[1], non_blocking=True)
            ev_in[k].record(s_in)
            if trace is not None:
                t["in1"] = mark(s_in)
        return t

    pending = {}
    held = [None, None]

    def one(i):
        # step i: its inputs were queued by the previous call (or the prologue); queue the NEXT step's upload first, so
        # that the copy engine is busy while the host is inside this step's forward (which waits for num_rendered)
        k = i % 2
        step, leaves, m2, state = sets[k]
        t = pending.pop(i) if i in pending else upload(i)
        pending[i + 1] = upload(i + 1)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in[k])
            s_comp.wait_event(ev_out[k])         # its previous results have left the device
            if trace is not None:
                t["c0"] = mark(s_comp)
            held[k] = None                        # ... so the device memory of those results may be reused from here on
            st = step()
            outs = {"color": st["color"], "allmap": st["allmap"], "radii": st["radii"], "g_means2D": m2.grad}
            outs.update({"g_" + name: v.grad for name, v in leaves.items()})
            # keep this step's results alive until the next step on this buffer set has waited for their download
            # (instead of Tensor.record_stream, which defers the allocator's reuse by an unpredictable number of steps
            # and makes it grow the pool -- multi-ms cudaMallocs -- in the middle of the timed region)
            held[k] = outs
            ev_comp[k].record(s_comp)
            if trace is not None:
                t["c1"] = mark(s_comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[k])
            if trace is not None:
                t["out0"] = mark(s_out)
            for name, v in outs.items():
                if name not in out_host[k]:
                    out_host[k][name] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                out_host[k][name].copy_(v.detach(), non_blocking=True)
            ev_out[k].record(s_out)
            done.append(mark(s_out))
            if trace is not None:
                t["out1"] = mark(s_out)
                trace.append(t)

    steps = max(4, args.steps)
    WARM = 6   # untimed steps: the caching allocator needs a few rounds until three steps' outputs can be in flight
    for i in range(WARM):
        one(i)
    # every timed step uploads exactly one set of inputs (the one for the step after it) and downloads one set of
    # results; the upload queued by the last warm-up step belongs to the first timed step
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(device)
    gc.collect()
    gc.disable()      # a cyclic-GC pause on the host starves this pipeline: every forward waits for the host
    e0.record(cur)
    for s_ in (s_in, s_out):
        s_.wait_stream(cur)
    for i in range(WARM, steps + WARM):
        one(i)
    for s_ in (s_in, s_out):
        cur.wait_stream(s_)
    e1.record(cur)
    torch.cuda.synchronize(device)
    gc.enable()
    ms = e0.elapsed_time(e1) / steps
    # per-step completion intervals on the device timeline: a step far above the median means the run was disturbed
    # (a host stall starves the pipeline, because every forward waits for the host to read num_rendered)
    intervals = [done[j - 1].elapsed_time(done[j]) for j in range(WARM + 1, len(done))]
    stats = {"median_step_ms": round(float(np.median(intervals)), 3), "worst_step_ms": round(float(max(intervals)), 3)}
    if trace:
        for i, t in enumerate(trace[WARM + 1:WARM + 7]):   # early timed steps, milliseconds since the start of the timed region
            try:
                print("e2e trace step", i + 1, {n: round(e0.elapsed_time(ev), 2) for n, ev in t.items()}, file=sys.stderr)
            except Exception as exc:   # diagnostics only
                print("e2e trace unavailable:", exc, file=sys.stderr)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    d2h = sum(v.numel() * v.element_size() for v in out_host[0].values())
    return ms, int(h2d), int(d2h), stats


def measure_e2e(mod, wl, args, world, device, sharded=None):
    """-> (ms per step, h2d bytes, d2h bytes, info).  Like the clock check of the device-resident loop, a disturbed run
    is re-measured ONCE: if some step of the timed region took more than 2x the median step (all ranks agree on the
    decision), the whole measurement is repeated and the faster of the two is reported, with both recorded in `info`."""
    ms, h2d, d2h, stats = _measure_e2e_once(mod, wl, args, world, device, sharded)
    info = dict(stats, remeasured=False)
    disturbed = stats["worst_step_ms"] > 2.0 * stats["median_step_ms"]
    if world > 1:
        f = torch.tensor([1 if disturbed else 0], device=device)
        torch.distributed.all_reduce(f, op=torch.distributed.ReduceOp.MAX)
        disturbed = bool(int(f.item()))
    if disturbed:
        ms2, _, _, stats2 = _measure_e2e_once(mod, wl, args, world, device, sharded)
        info = {"remeasured": True, "first_run": dict(stats, ms_per_step=round(ms, 3)),
                "second_run": dict(stats2, ms_per_step=round(ms2, 3)),
                "median_step_ms": stats2["median_step_ms"] if ms2 <= ms else stats["median_step_ms"],
                "worst_step_ms": stats2["worst_step_ms"] if ms2 <= ms else stats["worst_step_ms"]}
        ms = min(ms, ms2)
    return ms, h2d, d2h, info


def cpu_oracle_baseline(wl_host, cam, grads_host, threads=None):
    """Oracle port timed on the host cores (whole workload, one step)."""
    from oracle import oracle
    if threads:
        oracle.set_num_threads(threads)
    t0 = time.time()
    f = oracle.rasterize_forward(torch.zeros(3), wl_host["means3D"], None, wl_host["opacities"], wl_host["scales"],
                                 wl_host["rotations"], 1.0, None, cam.viewmatrix, cam.projmatrix, cam.tanfovx,
                                 cam.tanfovy, cam.height, cam.width, wl_host["shs"], 3, cam.campos)
    oracle.rasterize_backward(f, grads_host[0], grads_host[1])
    dt = time.time() - t0
    return wl_host["means3D"].shape[0] / dt / 1e6, dt, oracle.num_threads()


def config1_baseline(device=None, mod=None):
    """BASELINE configs[0]: 10k surfels, 256x256, SH degree 0 -- the pure-PyTorch CPU alpha-blend (oracle/torch_cpu_blend.py,
    checker/baseline infrastructure) timed on the host cores, and, when a device is given, this repo's operator on the
    same inputs (median of 20 calls after 3 warm-up calls)."""
    from oracle import torch_cpu_blend as tb
    cam, scene = syn.cam_s(), syn.box_scene(10_000, 3, 0)
    grads = syn.upstream_grads(cam.width, cam.height, "color_alpha")
    tb.forward_backward(scene, cam, grads)                      # warm-up (thread pool, allocator)
    times = [tb.forward_backward(scene, cam, grads)[1] for _ in range(2)]
    dt = float(np.median(times))
    out = {"value": round(10_000 / dt / 1e6, 6), "unit": UNIT, "cores": int(torch.get_num_threads()), "kind": "port",
           "host_cpus": os.cpu_count(),
           "sample": f"BASELINE configs[0] whole workload: BOX(P=10000, seed=3) CAM-S 256x256 SH0, colour+alpha grads, one "
                     f"fwd+bwd of the pure-PyTorch CPU alpha-blend (autograd backward), {dt:.2f} s"}
    if device is not None and mod is not None:
        import harness as hz
        st = hz._settings(mod, cam, torch.zeros(3), 0, 1.0, device)
        rast = mod.GaussianRasterizer(st)
        leaves = {k: scene[k].to(device).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        m2 = torch.zeros_like(leaves["means3D"], requires_grad=True)
        g = (grads[0].to(device), grads[1].to(device))

        def step():
            for t in list(leaves.values()) + [m2]:
                t.grad = None
            color, radii, allmap = rast(means3D=leaves["means3D"], means2D=m2, opacities=leaves["opacities"],
                                        shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
            torch.autograd.backward([color, allmap], [g[0], g[1]])
        _, med = timed_loop(step, 20, 3, 1, device)
        out["this_repo_gpu_ms"] = round(med, 4)
        out["this_repo_gpu_value"] = round(10_000 / (med * 1e-3) / 1e6, 3)
    return out


def workload_name(P_total, seed, world, strong=False):
    base = f"STREET(P={P_total}, seed={seed}) CAM-A 1920x1280 SH3, grads colour+alpha; BASELINE configs[{4 if strong else 2}]"
    if world > 1:
        base += (f"; index shards of {P_total // world} surfels per GPU, screen-band exchange of projected records and "
                 f"gradient records over NCCL/NVLink (streetunveiler_b200/sharded.py)")
    return base


def make_config(P_total, P_rank, seed, world, R, P_vis, crc, strong=False):
    """One function for both arms, so that equal workloads give equal `config` dicts."""
    return {"workload": workload_name(P_total, seed, world, strong), "P_total": P_total, "P_per_gpu": P_rank,
            "parallelism": f"index-shard x{world}" if world > 1 else "single", "num_rendered": R, "visible": P_vis,
            "input_crc32": crc,
            "l2": "per-GPU inputs (232 B/surfel, 464 MB at 2M) larger than the 126 MB L2; no explicit flush"}


def issue_roofline(kernel, ms_per_launch, clocks):
    """Second roofline of the dominant kernel (the blend kernels are instruction-issue bound, not HBM bound): executed
    warp instructions per launch (ncu `smsp__inst_executed.sum`, profiles/issue.json -- a property of kernel + workload,
    not of the clock) / the LIVE duration of this run, against 148 SMs x 4 schedulers x 1 warp instruction per cycle."""
    path = os.path.join(ROOT, "profiles", "issue.json")
    try:
        rec = json.load(open(path)).get(kernel)
    except Exception:
        rec = None
    if not rec:
        return None
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 148 * 4 * mhz * 1e6 / 1e9                      # G warp-instructions / s
    achieved = rec["inst_executed"] / (ms_per_launch * 1e-3) / 1e9
    return {"bound": "issue", "kernel": kernel, "achieved": round(achieved, 1), "peak": round(peak, 1),
            "unit": "G warp-inst/s", "frac": round(achieved / peak, 4), "inst_per_launch": rec["inst_executed"],
            "ncu_issue_active_pct": rec.get("issue_active_pct"), "source": rec.get("source")}


def gather_union(wl, world, device):
    """Every rank's shard, concatenated in rank order = the scene the sharded operator renders (on every rank)."""
    out = {}
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        x = wl.dev[k].contiguous()
        full = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        torch.distributed.all_gather_into_tensor(full, x)
        out[k] = full
    return out


def selfcheck_and_single(mod, wl, sharded, world, rank, device, steps, warmup, time_single):
    """Sharded fwd+bwd on all ranks, then the UNSHARDED operator on the same scene on rank 0: images must be
    bit-identical (CRC32 of colour + allmap), rank 0's shard of every gradient within 1e-4 (max-abs relative).  With
    `time_single` rank 0 also times the single-GPU step on that scene (the N=1 point of the strong-scaling curve)."""
    import zlib
    import harness as hz
    step, leaves, m2, state = make_step(mod, wl, sharded)
    st = step()
    torch.cuda.synchronize(device)
    sh_color, sh_allmap = st["color"].detach().clone(), st["allmap"].detach().clone()
    sh_grads = {k: v.grad.detach().clone() for k, v in leaves.items()}
    sh_grads["means2D"] = m2.grad.detach().clone()
    union = gather_union(wl, world, device)
    res = None
    if rank == 0:
        class _U:   # a Workload-shaped view of the union for make_step
            pass
        u = _U()
        u.cam, u.device, u.dev, u.grads, u.P = wl.cam, device, union, wl.grads, union["means3D"].shape[0]
        ustep, uleaves, um2, ustate = make_step(mod, u)
        ust = ustep()
        torch.cuda.synchronize(device)
        same = bool(torch.equal(ust["color"], sh_color) and torch.equal(ust["allmap"], sh_allmap))
        crc_sh = zlib.crc32(sh_allmap.cpu().numpy().tobytes(), zlib.crc32(sh_color.cpu().numpy().tobytes()))
        crc_un = zlib.crc32(ust["allmap"].detach().cpu().numpy().tobytes(),
                            zlib.crc32(ust["color"].detach().cpu().numpy().tobytes()))
        P = wl.P
        errs = {}
        for k, v in uleaves.items():
            errs[k] = hz.rel_err(sh_grads[k].cpu().numpy(), v.grad[:P].cpu().numpy())
        errs["means2D"] = hz.rel_err(sh_grads["means2D"].cpu().numpy(), um2.grad[:P].cpu().numpy())
        worst = max(errs.values()) if errs else 0.0
        res = {"scene_surfels": int(u.P), "image_bit_identical": same, "image_crc32_sharded": crc_sh,
               "image_crc32_single_gpu": crc_un, "grad_max_rel_err": {k: float(f"{v:.3e}") for k, v in errs.items()},
               "num_rendered_single_gpu": int(ustate["R"] or 0), "tolerance": 1e-4,
               "ok": bool(same and crc_sh == crc_un and worst <= 1e-4)}
        if time_single:
            ms1, med1 = timed_loop(ustep, steps, warmup, 1, device)
            res["single_gpu_ms_per_step"] = round(ms1, 4)
            res["single_gpu_ms_per_step_median"] = round(med1, 4)
        del ustep, uleaves, um2, ustate, ust, u
    del union, sh_color, sh_allmap, sh_grads
    torch.cuda.empty_cache()
    torch.distributed.barrier()
    return res


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        import datetime
        torch.distributed.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(minutes=20))
    from streetunveiler_b200 import _lib
    import harness as hz
    mod = hz.ours_module()

    strong_only = bool(args.total)
    P_total = (args.total // world) * world if args.total else P_PER_GPU * world
    seed = 1 if (world == 1 and not args.total) else 2
    wl = Workload(P_total, seed, world, rank, device)
    sharded = None
    if world > 1:
        from streetunveiler_b200.sharded import ShardedRasterizer
        sharded = ShardedRasterizer(world, rank)
    step, leaves, m2, state = make_step(mod, wl, sharded)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, ms_median = timed_loop(step, args.steps, args.warmup, world, device)
    clocks = sampler.stop() if rank == 0 else None
    if os.environ.get("BENCH_DEBUG"):
        print("debug: R after timed loop", state["R"], file=sys.stderr)

    # ---- kernel launches of one step, counted by CUPTI ----
    try:
        n_ours, n_other, launch_names = count_launches(step, device)
        launch_note = "counted with CUPTI (torch.profiler) over one step after the timed region, x steps"
    except Exception as exc:   # no CUPTI on this box: say so instead of guessing
        n_ours, n_other, launch_names = 0, 0, {}
        launch_note = f"NOT COUNTED: torch.profiler unavailable ({type(exc).__name__}: {exc})"

    # ---- e2e: host (pinned) buffers in, results out, every step ----
    ms_e2e, h2d, d2h, e2e_info = measure_e2e(mod, wl, args, world, device, sharded)
    if os.environ.get("BENCH_DEBUG"):
        print("debug: R after e2e loop", state["R"], file=sys.stderr)

    # ---- per-stage device time (roofline of the dominant kernel) ----
    roof = None
    stages = {}
    if world == 1:
        _lib.set_option("time_stages", 1)
        for _ in range(max(3, args.steps // 2)):
            step()
        torch.cuda.synchronize(device)
        stages = {k: (ms / max(n, 1)) for k, (ms, n) in _lib.stage_times().items()}
        _lib.set_option("time_stages", 0)

    R = int(state["R"] or 0)
    P_vis = int((state["radii"] > 0).sum().item())
    HW = wl.cam.width * wl.cam.height
    peak, peak_src = measured_peaks()
    roof_issue = None
    if stages:
        dom = max(stages, key=stages.get)
        ab = alg_bytes_stage(dom, wl.P, P_vis, R, HW)
        achieved = ab / (stages[dom] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": int(ab), "ms_per_launch": round(stages[dom], 4),
                "note": "the blend kernels are instruction-issue bound (ncu: DRAM < 2 % of peak); `roofline_issue` is the "
                        "bound that explains their duration, this entry is the HBM fraction the contract asks for"}
        roof_issue = issue_roofline(dom, stages[dom], clocks)

    # ---- N > 1: sharded-vs-single parity self-check and BASELINE configs[4] (one 8M scene, strong scaling) ----
    selfcheck, strong = None, None
    phases = None
    if world > 1:
        from streetunveiler_b200 import sharded as _sh
        selfcheck = {"weak_scene": selfcheck_and_single(mod, wl, sharded, world, rank, device, args.steps, args.warmup,
                                                        time_single=strong_only)}
        if not strong_only and not args.no_strong:
            del step, leaves, m2, state
            wl2 = Workload((STRONG_TOTAL // world) * world, 2, world, rank, device)
            step2, leaves2, m22, state2 = make_step(mod, wl2, sharded)
            ms2, med2 = timed_loop(step2, args.steps, args.warmup, world, device)
            chk = selfcheck_and_single(mod, wl2, sharded, world, rank, device, args.steps, args.warmup, time_single=True)
            selfcheck["strong_scene"] = chk
            if rank == 0:
                n1 = chk.get("single_gpu_ms_per_step")
                strong = {"workload": workload_name(wl2.P * world, 2, world, strong=True), "P_total": wl2.P * world,
                          "n_gpus": world, "ms_per_step": round(ms2, 4), "ms_per_step_median": round(med2, 4),
                          "value": round(wl2.P * world / (ms2 * 1e-3) / 1e6, 2), "unit": UNIT, "scaling": "strong",
                          "n1_ms_per_step": n1, "speedup_vs_n1": round(n1 / ms2, 3) if n1 else None,
                          "n1_note": "single-GPU step of the SAME scene (the union of the ranks' shards), timed on rank 0 of "
                                     "this run with the same steps/warmup while the other ranks wait"}
        if _sh.PHASE_MS and rank == 0:
            import statistics
            phases = {k: round(statistics.median(v), 3) for k, v in _sh.PHASE_MS.items()}
            print("shard phases, median ms per call (sync'd):", phases, file=sys.stderr)
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    value = P_total / (ms_step * 1e-3) / 1e6
    step_bytes = alg_bytes_step(wl.P, R, HW)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "ms_per_step_median": round(ms_median, 4),
        "higher_is_better": True,
        "scaling": "strong" if args.total else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(make_config(P_total, wl.P, seed, world, R, P_vis, wl.crc, strong=strong_only),
                       **({"image_exchange": sharded.backend.image_exchange, "row_exchange": sharded.backend.row_exchange}
                          if sharded is not None else {})),
        "clocks": clocks,
        "e2e": {"value": round(P_total / (ms_e2e * 1e-3) / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e, 3), "step_stats": e2e_info},
        "gpu_launches": int(n_ours) * args.steps,
        "gpu_launches_per_step": {"hand_written": int(n_ours), "other": int(n_other), "kernels": launch_names},
        "gpu_launches_note": launch_note,
        "roofline": roof,
        "roofline_issue": roof_issue,
        "step_roofline": {"alg_bytes": int(step_bytes), "achieved_gbs": round(step_bytes / (ms_step * 1e-3) / 1e9, 2),
                          "frac_of_peak": round(step_bytes / (ms_step * 1e-3) / 1e9 / peak, 4),
                          "formula": "712 P + 36 R + 80 HW (SURVEY.md 8d)"},
        "stage_ms": {k: round(v, 4) for k, v in stages.items()},
    }
    if selfcheck is not None:
        line["parity_selfcheck"] = selfcheck
    if strong is not None:
        line["config5_strong"] = strong
    if phases:
        line["shard_phase_ms"] = phases
    if sharded is not None:
        from streetunveiler_b200 import sharded as _shm
        info = _shm.LAST_INFO
        if info.get("send"):
            snd, rcv = info["send"], info["recv"]
            out_rows = sum(snd) - snd[rank]
            back_rows = sum(rcv) - rcv[rank]
            HW_ = wl.cam.width * wl.cam.height
            owned = info["cuts"][rank + 1] - info["cuts"][rank]
            line["shard_exchange_rank0"] = {
                "records_sent_rows": int(sum(snd)), "records_leaving_gpu_rows": int(out_rows),
                "records_leaving_gpu_bytes": int(out_rows * 104), "gradient_rows_leaving_gpu_bytes": int(back_rows * 80),
                "image_bytes_stored_by_this_rank": int(owned * 256 * 40 * (1 if sharded.backend.image_exchange == "multicast" else world)),
                "rows_per_visible_gaussian": round(sum(snd) / max(1, P_vis), 3), "tile_range": [info["cuts"][rank], info["cuts"][rank + 1]],
                "note": "last step of the run on rank 0 (the strong-scaling scene when that leg ran); NVLink payload only, "
                        "records travel as 96 B + 4 B key + 4 B radius, gradient rows as 80 B, image pixels as 40 B"}
    if sharded is not None and sharded.balancer is not None and sharded.balancer.last_times_us:
        t = sharded.balancer.last_times_us
        line["shard_balance"] = {"window_fwd_bwd_us_per_rank": t, "imbalance_max_over_mean": round(max(t) / (sum(t) / len(t)), 3),
                                 "cost_shares": [round(x, 4) for x in sharded.balancer.shares],
                                 "note": "last measured step of the run (the strong-scaling scene when that leg ran)"}
    if world == 1 and not args.total and not args.no_strong:
        # the N = 1 point of BASELINE configs[4]: the 8M scene on one GPU
        del step, leaves, m2, state
        torch.cuda.empty_cache()
        wl8 = Workload(STRONG_TOTAL, 2, 1, 0, device)
        step8, _l8, _m8, state8 = make_step(mod, wl8)
        ms8, med8 = timed_loop(step8, max(5, args.steps // 2), args.warmup, 1, device)
        line["config5_strong"] = {"workload": workload_name(STRONG_TOTAL, 2, 1, strong=True), "P_total": STRONG_TOTAL,
                                  "n_gpus": 1, "ms_per_step": round(ms8, 4), "ms_per_step_median": round(med8, 4),
                                  "value": round(STRONG_TOTAL / (ms8 * 1e-3) / 1e6, 2), "unit": UNIT, "scaling": "strong",
                                  "num_rendered": int(state8["R"] or 0), "input_crc32": wl8.crc}
        del step8, _l8, _m8, state8, wl8
    if world == 1 and not args.no_cpu_baseline:
        v, dt, th = cpu_oracle_baseline(wl.host, wl.cam, wl.grads_host)
        line["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": th, "kind": "port",
                                "sample": f"whole workload (P={wl.P}), 1 step fwd+bwd of the C oracle, {dt:.1f} s"}
        try:
            line["cpu_baseline_config1"] = config1_baseline(device, mod)
        except Exception as exc:
            line["cpu_baseline_config1"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    print(json.dumps(line))


def run_reference(args, force_cpu=False):
    rank, world, local = dist_env()
    if rank != 0:
        return
    import harness as hz
    use_gpu_ref = (not force_cpu) and torch.cuda.is_available() and hz.reference_available()
    cam = syn.cam_a()
    P_vis = None
    if use_gpu_ref:
        device = torch.device("cuda", local)
        torch.cuda.set_device(device)
        wl = Workload(P_PER_GPU, 1, 1, 0, device)
        step, leaves, m2, state = make_step(hz.reference_module(), wl)
        sampler = ClockSampler(local)
        sampler.start()
        ms_step, ms_median = timed_loop(step, args.steps, args.warmup, 1, device)
        clocks = sampler.stop()
        value = wl.P / (ms_step * 1e-3) / 1e6
        ms_e2e, h2d, d2h, e2e_info = measure_e2e(hz.reference_module(), wl, args, 1, device)
        e2e = {"value": round(wl.P / (ms_e2e * 1e-3) / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e, 3), "step_stats": e2e_info,
               "note": "the reference is a CUDA extension too, so its host-buffer call needs the same PCIe copies as ours; "
                       "device-resident throughput is `value`"}
        kind, cores = "reference", 0
        sample = ("UNMODIFIED reference CUDA extension (oracle/_ref, rebuilt for sm_100a) on the same GPU, whole "
                  "workload; the reference has no CPU implementation of this path")
        R = int(state["R"] or 0)
        P_vis = int((state["radii"] > 0).sum().item())
        crc = wl.crc
    else:
        scene = syn.street_scene(P_PER_GPU, 1, 3)
        crc = syn.scene_crc(scene)
        grads = syn.upstream_grads(cam.width, cam.height, "color_alpha")
        times = []
        for _ in range(max(1, min(args.steps, 2))):
            v, dt, cores = cpu_oracle_baseline(scene, cam, grads)
            times.append(dt)
        ms_step = float(np.mean(times)) * 1e3
        ms_median = float(np.median(times)) * 1e3
        value = P_PER_GPU / (ms_step * 1e-3) / 1e6
        kind, clocks, R = "port", None, None
        e2e = {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        sample = "C oracle port of the reference algorithm on all host threads, whole workload"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
        "ms_per_step_median": round(ms_median, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(P_PER_GPU, P_PER_GPU, 1, 1, R, P_vis, crc),
        "clocks": clocks,
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": e2e,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the 8M-surfel configs[4] leg (quick runs)")
    ap.add_argument("--total", type=int, default=0,
                    help="strong scaling: total surfels split over the ranks (BASELINE configs[4]: --total 8000000); "
                         "default 0 = weak scaling with 2,000,000 surfels per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "ours":
        run_ours(args)
    else:
        run_reference(args, force_cpu=(args.impl == "reference-cpu"))


if __name__ == "__main__":
    main()
