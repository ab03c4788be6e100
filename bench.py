#!/usr/bin/env python
"""bench.py — the localization hot path (rasterise + score + normalise + resample) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload global|tracking|grid]

One "step" = one whole ParticleFilter update on one synthetic scan (TopDownRender::takeStep minus ROS,
top_down_render.cpp:505-572): polar rasterise of a 65 536-point semantic scan, score of every particle
(40-shift polar theta search while the heading is unknown), weight normalisation, systematic resample.
Metric (BASELINE.json): particle scores/sec — one score = one getCostForRot evaluation, i.e. one
(x, y, theta) hypothesis (state_particle.cpp:112-155).

Default workload = BASELINE cfg3 "global localization": 1M particles per GPU over a 4000 x 4000 px map
(4 km^2 at 0.5 m/px), 6 classes, heading unknown.  N > 1: particles are sharded (weak scaling, 1M per
GPU), the map / scan are replicated, and ONE all-gather of (weights + states) per step over NCCL feeds a
normalise + order-exact prefix done redundantly on every rank (bit-exact indices independent of N).

Rank 0 prints ONE JSON line (see the driver contract).  `--impl reference` times the CPU restatement of
the reference (oracle/, all host threads) on bounded samples of the same workload instead.  The partial build of
the reference's own sources (oracle/_ref) is a correctness pin only: against stand-in Eigen headers and on one
thread it is 4x slower per thread than the restatement, which therefore is the conservative CPU arm (kind "port").
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_THETA, N_R = 100, 25
P = N_THETA * N_R
ANG_RES = np.float32(2 * math.pi / 100)
SEED = 1234

WORKLOADS = {
    # name: (map side px, classes, particles per GPU, res m/bin, shifts per particle, description)
    "global": dict(side=4000, C=6, n=1_000_000, res=4.0, shifts=40,
                   desc="cfg3 global localization: 1M particles/GPU x 40-shift polar theta search, 4000x4000 px map "
                        "(4 km^2 @0.5 m/px), 6 classes, 65536-pt scan; rasterise+score+normalise+resample"),
    "grid": dict(side=4000, C=6, n=1_000_000, res=4.0, shifts=100, grid=True,
                 desc="cfg4 exhaustive grid: 1000x1000 centres (stride 4 px over 4000x4000 px, 6 classes) x 100 row shifts = "
                      "1e8 (x, y, theta) hypotheses in total, sharded over the GPUs; rasterise + all costs + weight all-gather "
                      "+ arg-min"),
    "refine": dict(side=4000, C=6, n=0, res=0.5, shifts=0, refine=True, scans=10_000, chunk_scans=64, distinct_chunks=8,
                   desc="cfg5 refine_map batch: 10k recorded scans (65536 pts each, 6.6e8 points) binned with refine_map's rule "
                        "into a 4000x4000 px x 6-class count map + per-class distance-field rebuild"),
    "tracking": dict(side=2000, C=6, n=10_000, res=0.5, shifts=1,
                     desc="cfg2 tracking: 10k particles, 2000x2000 px map, 6 classes, 65536-pt scan; "
                          "rasterise+score+normalise+resample"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def b_score(C):
    """algorithmic bytes per scored particle, SURVEY.md section 8d: P*(4C+1) + 28 B state + 4 B weight"""
    return P * (4 * C + 1) + 32


# ------------------------------------------------------------------------------------------------
# synthetic world
# ------------------------------------------------------------------------------------------------
def make_inputs(wl, rank=0):
    from top_down_renderer_b200 import synth
    side, C, n = wl["side"], wl["C"], wl["n"]
    cm = synth.make_class_map(side, side, C, seed=SEED)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C)
    pose, heading = synth.default_pose(cm, seed=SEED)
    pts = synth.make_scan(cm, pose, heading, seed=SEED)
    if wl["shifts"] == 1:
        st, ld = synth.particles_tracking(n, pose, heading, seed=SEED + 101 * rank)
    else:
        st, ld = synth.particles_global(n, cm, seed=SEED + 101 * rank)
    return dict(cm=cm, img=img, lut=lut, pose=pose, heading=heading, pts=pts, st=st, ld=ld)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference (cpu_baseline / --impl reference)
# ------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, wl, inp):
        from oracle import oracle as orc
        self.orc, self.wl, self.inp = orc, wl, inp
        orc.build()
        C = wl["C"]
        self.cores = os.cpu_count() or 1
        bl = orc.class_image_to_layers(inp["img"], inp["lut"], C, 1.0)
        self.layers, self.mask = orc.compute_dists(bl, 1.0)        # set-up, not timed (map precompute)
        self.tab = orc.polar_table(N_THETA, N_R, ANG_RES, 1.0)
        self.thetas, self.shifts = orc.search_list(N_THETA)
        self.fp = orc.make_params(C, regularization=0.7, map_width=wl["side"], map_height=wl["side"])
        self.u = orc.uniform_draw(SEED)
        self._bin_layers, self._geo = bl, None

    def geo_layers(self):
        """distance fields of the two geometric layers (getGeoRasterMap): what the reference gathers per particle and
        never uses (state_particle.cpp:189) — only the "literal" CPU figure pays for it"""
        if self._geo is None:
            self._geo, _ = self.orc.compute_dists(self.orc.geo_raster(self._bin_layers), 1.0)
        return self._geo

    def step(self, lo, n, literal=False):
        """one reference update on particles [lo, lo+n): render + score (all host threads) + normalise + resample.
        literal: with the unused per-particle geo gather of state_particle.cpp:189 kept"""
        orc, wl, inp = self.orc, self.wl, self.inp
        if wl.get("grid"):
            from top_down_renderer_b200 import synth
            if not hasattr(self, "centers"):
                self.centers = synth.grid_centers(wl["side"], wl["side"], 4)
            t0 = time.perf_counter()
            scan = orc.render_polar(inp["pts"], wl["res"], ANG_RES, N_THETA, N_R, inp["lut"], wl["C"])
            costs = orc.cost_grid(self.centers[lo:lo + n], 2.0, self.fp, self.layers, self.mask, 1.0, self.tab, N_THETA, N_R,
                                  scan, wl["res"], np.arange(N_THETA, dtype=np.int32), n_threads=self.cores)
            np.nanargmin(costs)
            return time.perf_counter() - t0
        st = inp["st"][lo:lo + n].copy()
        ld = inp["ld"][lo:lo + n]
        t0 = time.perf_counter()
        scan = orc.render_polar(inp["pts"], wl["res"], ANG_RES, N_THETA, N_R, inp["lut"], wl["C"])
        w = orc.score_all(st, self.fp, self.layers, self.mask, 1.0, self.tab, N_THETA, N_R, scan, wl["res"],
                          self.thetas, self.shifts, geo_layers=self.geo_layers() if literal else None, n_threads=self.cores)
        wn, _, _ = orc.normalize(w, ld)
        orc.resample_fast(wn, self.u, n)
        return time.perf_counter() - t0


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp = make_inputs(dict(wl, n=16, shifts=1) if wl.get("grid") else wl)
    arm = CpuArm(wl, inp)
    n_s = min(wl["n"], (2048 if wl.get("grid") else 8192) if wl["shifts"] > 1 else wl["n"])
    for i in range(args.warmup):
        arm.step((i * n_s) % max(1, wl["n"] - n_s), n_s)
    ts = []
    for i in range(args.steps):
        ts.append(arm.step(((i + args.warmup) * n_s) % max(1, wl["n"] - n_s), n_s))
    total = float(np.sum(ts))
    value = n_s * wl["shifts"] * args.steps / total
    sample = (f"{n_s} of the workload's {wl['n']} particles per step (x{wl['shifts']} shifts), full scan render + "
              f"normalise + O(N) resample; oracle/tdr_oracle.cpp g++ -O2, {arm.cores} std::threads")
    out = {"impl": "reference", "metric": "particle_scores_per_sec", "value": value, "unit": "scores/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
           "higher_is_better": True, "scaling": "strong" if wl.get("grid") else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": {"workload": wl["desc"], "particles_per_step_sampled": n_s},
           "cpu_baseline": {"value": value, "unit": "scores/s", "cores": arm.cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def setup_ctx(wl, inp, device):
    from top_down_renderer_b200 import hostmath
    from top_down_renderer_b200.core import Context
    ctx = Context(device)
    C = wl["C"]
    ctx.map_set_class_image(inp["img"], inp["lut"], C, 1.0)
    ctx.map_set_polar_table(hostmath.polar_table(N_THETA, N_R, ANG_RES, 1.0), N_THETA, N_R)
    ctx.scan_set_lut(inp["lut"], C)
    ctx.pf_set_params(C, regularization=0.7)
    th, sh = hostmath.search_list(N_THETA)
    ctx.pf_set_search(th, sh)
    ctx.pf_set_states(inp["st"], inp["ld"])
    ctx.pf_checkpoint()
    return ctx


def init_dist():
    """(rank, world, local) of this process; joins the NCCL group once when launched under torchrun"""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def states_digest(states):
    """order-sensitive digest of a resampled particle set (the 28-byte State records, padding zeroed)"""
    import hashlib
    a = np.ascontiguousarray(states).copy()
    a["pad"] = 0
    return hashlib.blake2b(a.tobytes(), digest_size=8).hexdigest()


def verify_update(ctx, wl, inp, arm, u, n_sample=4096):
    """Outside every timed region: one update of the benchmarked configuration, stage by stage through the C ABI,
    against the oracle — class images bit-exact, a sample of raw weights within 1e-5 (a heading that differs must be a
    tie to within 1e-5), normalised weights and ALL resampled indices bit-exact given the device's own raw weights."""
    orc = arm.orc
    n, C, res = wl["n"], wl["C"], wl["res"]
    out = {"sample_particles": int(min(n_sample, n))}
    ctx.pf_restore()
    scan = ctx.scan_render_polar(res, float(ANG_RES), N_THETA, N_R, want=True)
    scan_o = orc.render_polar(inp["pts"], res, ANG_RES, N_THETA, N_R, inp["lut"], C)
    out["class_images_bit_exact"] = bool(np.array_equal(scan, scan_o))
    w = ctx.pf_score(res)
    st_g = ctx.pf_get_states()
    pick = np.sort(np.random.default_rng(SEED).choice(n, out["sample_particles"], replace=False))
    st_o = inp["st"][pick].copy()
    w_o = orc.score_all(st_o, arm.fp, arm.layers, arm.mask, 1.0, arm.tab, N_THETA, N_R, scan_o, res, arm.thetas, arm.shifts,
                        n_threads=arm.cores)
    both_nan = np.isnan(w[pick]) & np.isnan(w_o)
    err = np.abs(w[pick].astype(np.float64) - w_o) / np.maximum(np.abs(w_o.astype(np.float64)), 1e-300)
    err[both_nan] = 0
    err[np.isnan(w[pick]) ^ np.isnan(w_o)] = np.inf
    out["weights_max_rel_err"] = float(err.max())
    diff = np.flatnonzero(st_g["theta"][pick] != st_o["theta"])
    out["headings_differing"] = int(diff.size)
    gap = 0.0
    if diff.size:
        src = inp["st"][pick][diff]
        cen = np.stack([(src["dx_m"] * src["scale"]).astype(np.float32) + src["init_x_px"],
                        (src["dy_m"] * src["scale"]).astype(np.float32) + src["init_y_px"]], axis=1).astype(np.float32)
        costs = orc.cost_grid(cen, 2.0, arm.fp, arm.layers, arm.mask, 1.0, arm.tab, N_THETA, N_R, scan_o, res, arm.shifts,
                              n_threads=arm.cores)
        th = np.asarray(arm.thetas, dtype=np.float32)
        kg = [int(np.flatnonzero(th == t)[0]) for t in st_g["theta"][pick][diff]]
        ko = [int(np.flatnonzero(th == t)[0]) for t in st_o["theta"][diff]]
        cg = costs[np.arange(diff.size), kg].astype(np.float64)
        co = costs[np.arange(diff.size), ko].astype(np.float64)
        gap = float(np.max(np.abs(cg - co) / np.maximum(np.abs(co), 1e-300)))
    out["heading_ties_max_gap"] = gap
    ctx.pf_normalize()
    wn = ctx.pf_get_weights(n)
    wn_o, _, _ = orc.normalize(w, inp["ld"])
    out["normalised_bit_exact"] = bool(np.array_equal(wn.view(np.uint32), wn_o.view(np.uint32)))
    idx = ctx.pf_resample(u, n)
    out["indices_bit_exact"] = bool(np.array_equal(idx, orc.resample_fast(wn, u, n)))
    ctx.pf_restore()
    out["verified"] = bool(out["class_images_bit_exact"] and out["weights_max_rel_err"] <= 1e-5 and gap <= 1e-5 and
                           out["normalised_bit_exact"] and out["indices_bit_exact"])
    return out


def more_warmup(t0, world, local, need_s=1.5):
    """True while the clock sampler has not yet seen need_s seconds of load (decided collectively: every rank
    must run the same number of steps)."""
    import torch
    torch.cuda.synchronize()
    more = 1.0 if time.perf_counter() - t0 < need_s else 0.0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([more], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        more = float(t.item())
    return more > 0


def measure_particles(args, wl, workload_name, rank, world, local, steps, warmup, with_cpu, with_verify):
    """one particle-filter workload (cfg3 global / cfg2 tracking) on `world` ranks; returns rank 0's record"""
    import torch
    import torch.distributed as dist
    from top_down_renderer_b200 import sharded

    class _A:
        pass
    a_ = _A()
    a_.steps, a_.warmup, a_.workload, a_.no_cpu = steps, warmup, workload_name, not with_cpu
    a_.shard_impl = getattr(args, "shard_impl", "library")
    args = a_
    inp = make_inputs(wl, rank)
    ctx = setup_ctx(wl, inp, local)
    ctx.profile_enable(True)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    shard_impl = getattr(args, "shard_impl", "library")
    n, C, res = wl["n"], wl["C"], wl["res"]
    M = n                                   # resample back to the same particle count (per GPU)
    flt = None
    if world > 1:
        # "library": NCCL + peer-mapped state slots below the C ABI (csrc/shard.cu); "torch": the round-1 harness over
        # torch.distributed with the states all-gathered to every rank
        flt = (sharded.LibraryShardedFilter(ctx, rank, world, n) if shard_impl == "library"
               else sharded.ShardedFilter(ctx, stream, rank, world))
    u = float(np.random.default_rng(SEED).random(dtype=np.float32))
    pts_pinned = torch.from_numpy(inp["pts"]).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    arm, verification = None, None
    if rank == 0 and (with_verify or (with_cpu and world == 1)):
        arm = CpuArm(wl, inp)
    if with_verify:
        ctx.scan_set_points_ptr(pts_pinned.data_ptr(), 32, 16, pts_pinned.shape[0])
        if rank == 0:
            verification = verify_update(ctx, wl, inp, arm, u)
        if world > 1:
            dist.barrier()

    def one_step():
        if flt is not None:
            flt.step(res, float(ANG_RES), N_THETA, N_R, u, M * world)
        else:
            ctx.step(res, float(ANG_RES), N_THETA, N_R, u, M)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident run: scan points already in HBM
    ctx.scan_set_points_ptr(pts_pinned.data_ptr(), 32, 16, pts_pinned.shape[0])
    ctx.sync()
    launches0 = None
    evs, stage = [], []
    sampler = ClockSampler(local)
    sampler.start()                          # nvidia-smi needs ~1 s to produce its first sample: start it before the
    t_load0 = time.perf_counter()            # warm-up and keep warming up (same load, untimed) until it has samples
    n_warm = args.warmup
    total_steps = args.warmup + args.steps
    i = -1
    while True:
        i += 1
        if i == n_warm and more_warmup(t_load0, world, local) and n_warm < args.warmup + 5000:
            n_warm += 1
            total_steps += 1
        if i >= total_steps:
            break
        if i == n_warm:
            barrier()
            launches0 = ctx.launch_count()
            t_wall0 = time.perf_counter()
        with torch.cuda.stream(stream):
            ctx.pf_restore()                 # same prior every step (outside the timed events)
            flush.zero_()                    # L2 flush: 256 MiB write > 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one_step()
            e1.record(stream)
        if i >= n_warm:
            evs.append((e0, e1))
            stage.append(ctx.profile_stage_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(step_ms.sum())
    stage = np.array(stage)
    if world > 1:
        t = torch.tensor([total_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    scores_per_step = n * wl["shifts"] * world
    value = scores_per_step * args.steps / (total_ms * 1e-3)

    # ---- end to end: host scan in pinned memory -> H2D -> step -> pose D2H, wall clock around the public calls
    e2e_t = []
    for i in range(args.warmup + args.steps):
        with torch.cuda.stream(stream):
            ctx.pf_restore()
            flush.zero_()
        barrier()
        t0 = time.perf_counter()
        ctx.scan_set_points_ptr(pts_pinned.data_ptr(), 32, 16, pts_pinned.shape[0])     # H2D 2 MiB
        with torch.cuda.stream(stream):
            one_step()
        if flt is not None:
            mean, cov, _, _ = flt.pose(want_ml=False)                                    # 2nd all-gather + D2H pose
        else:
            mean, cov, _, _ = ctx.pf_pose(want_ml=False)                                 # D2H pose (synchronises)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_t.append(dt)
    e2e_total = float(np.sum(e2e_t))
    if world > 1:
        t = torch.tensor([e2e_total], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = scores_per_step * args.steps / e2e_total

    # ---- N > 1: the sharded update reproduces the single-GPU states bit for bit (the design's claim) — one more
    # update from the checkpoint on every rank, digests of the resampled shards against rank 0 running ALL particles
    shard_check = None
    if world > 1 and with_verify:
        with torch.cuda.stream(stream):
            ctx.pf_restore()
            one_step()
        torch.cuda.synchronize()
        mine = states_digest(ctx.pf_get_states())
        digests = [None] * world
        dist.all_gather_object(digests, mine)
        if rank == 0:
            from top_down_renderer_b200 import synth
            parts = [synth.particles_global(n, inp["cm"], seed=SEED + 101 * r) if wl["shifts"] > 1 else
                     synth.particles_tracking(n, inp["pose"], inp["heading"], seed=SEED + 101 * r) for r in range(world)]
            big = dict(inp, st=np.concatenate([p_[0] for p_ in parts]), ld=np.concatenate([p_[1] for p_ in parts]))
            ref = setup_ctx(dict(wl, n=n * world), big, local)
            ref.scan_set_points(inp["pts"])
            ref.step(res, float(ANG_RES), N_THETA, N_R, u, M * world)
            ref.sync()
            st_ref = ref.pf_get_states()
            ref.close()
            want = [states_digest(st_ref[r * M:(r + 1) * M]) for r in range(world)]
            shard_check = {"ranks": world, "states_equal_single_gpu": bool(want == digests), "digests": digests}
        dist.barrier()

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        score_ms = float(stage[:, 1].mean())
        achieved = n * b_score(C) / (score_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json")) as f:
                t_ = json.load(f).get(args.workload, {})
                traffic = int(t_["captured_dram_bytes"] * n / t_["captured_particles"])      # per launch of n particles
                traffic_src = t_.get("source", "profiles/ ncu --set full capture of this kernel, scaled by particles; NOT measured in this run")
        except Exception:
            pass
        out = {"metric": "particle_scores_per_sec", "value": value, "unit": "scores/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": wl["desc"], "particles_per_gpu": n, "shifts_per_particle": wl["shifts"],
                          "map_px": [wl["side"], wl["side"]], "classes": C, "polar_image": [N_THETA, N_R],
                          "res_m_per_bin": res, "parallelism": f"particle shards x{world}, map replicated"
                          + ((", ncclAllGather of 8 B/particle inside libtdr_b200 + resampled states read from the owner's peer-mapped slot"
                              if args.shard_impl == "library" else ", all-gather(weights) + all-gather(states) via torch.distributed") if world > 1 else ""),
                          "l2": "flushed (256 MiB write) and particle set rolled back between steps, outside the timed events"},
               "clocks": clocks,
               "p50_update_ms": float(np.median(step_ms)),
               "stage_ms": {"render": float(stage[:, 0].mean()), "score": score_ms,
                            "normalize": float(stage[:, 2].mean()), "resample": float(stage[:, 3].mean())},
               "e2e": {"value": e2e_value, "unit": "scores/s", "h2d_bytes_per_step": int(inp["pts"].nbytes),
                       "d2h_bytes_per_step": 20 * 4, "ms_per_step": 1e3 * e2e_total / args.steps,
                       "p50_ms": 1e3 * float(np.median(e2e_t)), "timer": "host wall clock around set_points+step+pose"},
               "gpu_launches": int(launches),
               "wall_s_timed_region": t_wall,
               "roofline": {"bound": "hbm", "kernel": "k_score_mma_i8 (tcgen05 kind::i8 gather-GEMM on 16-byte records, operands in tensor memory)" if wl["shifts"] > 1 else "k_score_track",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": n * b_score(C), "kernel_ms": score_ms,
                            "note": ("frac > 1 is expected here: the algorithmic bytes are SURVEY 8d's figure in the reference's fp32 planar "
                                     "format, the kernel moves 16 B per visited cell and those come out of L2 (see traffic); what bounds it is "
                                     "the L1 data pipe, 86 % busy in profiles/r02_score_i8_ncu.csv") if wl["shifts"] > 1 else None}}
        if verification is not None:
            out["verified"] = verification["verified"]
            out["verification"] = verification
        if shard_check is not None:
            out["multi_gpu_check"] = shard_check
    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = min(n, 32768 if wl["shifts"] > 1 else n)
        arm.step(0, min(n_s, 2048))
        reps = 1 if wl["shifts"] > 1 else 20
        tt = sum(arm.step(0, n_s) for _ in range(reps))
        out["cpu_baseline"] = {"value": n_s * wl["shifts"] * reps / tt, "unit": "scores/s", "cores": arm.cores,
                               "kind": "port",
                               "sample": f"{reps} update(s) of {n_s} of the {n} particles (x{wl['shifts']} shifts) incl. scan render, "
                                         f"normalise, O(N) resample; oracle/tdr_oracle.cpp (g++ -O2, no -march), "
                                         f"{arm.cores} std::threads"}
        # BASELINE.md section 5's two figures: "tidy" is the one above (the unused geo gather removed, O(N) resampler);
        # "literal" keeps the geo gather per particle and the reference's O(N*M) resampler, the latter timed at 1e4 x 1e4
        # and extrapolated quadratically to this workload's N = M (labelled as such)
        arm.geo_layers()
        tl = sum(arm.step(0, n_s, literal=True) for _ in range(reps))
        m_lit = min(n, 10_000)
        w_lit = np.full(m_lit, 1.0 / m_lit, dtype=np.float32)
        t0 = time.perf_counter()
        arm.orc.resample_literal(w_lit, arm.u, m_lit)
        t_res = time.perf_counter() - t0
        out["cpu_baseline"]["tidy_value"] = out["cpu_baseline"]["value"]
        out["cpu_baseline"]["literal_value"] = n_s * wl["shifts"] * reps / tl
        out["cpu_baseline"]["literal_resample"] = {"measured_s": t_res, "at": [m_lit, m_lit],
                                                   "extrapolated_s_at_workload": t_res * (n / m_lit) ** 2,
                                                   "note": "particle_filter.cpp:174-185 is O(N*M) on one thread; extrapolated, not run"}
    elif rank == 0:
        out["cpu_baseline"] = None
    if flt is not None and hasattr(flt, "close"):
        flt.close()
    ctx.close()
    if world > 1:
        dist.barrier()
    return out


def grid_traffic(n_local):
    """DRAM bytes of the grid score kernel per launch, from the committed ncu capture (scaled to this rank's share)"""
    try:
        with open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json")) as f:
            g = json.load(f)["grid"]
        return int(g["captured_dram_bytes"] * n_local / g["captured_particles"])
    except Exception:
        return None


def measure_grid(wl, rank, world, local, steps, warmup, grid_collective, with_verify):
    """cfg4: exhaustive (x, y, theta) grid, STRONG scaling: the lattice of centres is split over the ranks; one
    all-gather of the costs ("weight all-gather") per step, then the arg-min on every rank.  world = 1 with rank 0 of a
    larger job = the single-GPU reference of the same run.  Returns rank 0's record."""
    import torch
    import torch.distributed as dist
    from top_down_renderer_b200 import hostmath, sharded, synth
    from top_down_renderer_b200.core import Context

    class _A:
        pass
    args = _A()
    args.steps, args.warmup, args.grid_collective = steps, warmup, grid_collective
    side, C, res = wl["side"], wl["C"], wl["res"]
    inp = make_inputs(dict(wl, n=16, shifts=1), rank)
    centers_all = synth.grid_centers(side, side, 4)                       # 1000 x 1000 centres
    n_total = centers_all.shape[0]
    assert n_total % world == 0
    lo, hi = sharded.shard_range(n_total, rank, world)
    centers = np.ascontiguousarray(centers_all[lo:hi])
    n_local = hi - lo
    shifts = np.arange(N_THETA, dtype=np.int32)
    S = len(shifts)
    ctx = Context(local)
    ctx.map_set_class_image(inp["img"], inp["lut"], C, 1.0)
    ctx.map_set_polar_table(hostmath.polar_table(N_THETA, N_R, ANG_RES, 1.0), N_THETA, N_R)
    ctx.scan_set_lut(inp["lut"], C)
    ctx.pf_set_params(C, regularization=0.7)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    dev = f"cuda:{local}"
    fused = world > 1 and args.grid_collective in ("fused", "fused-nccl")
    tiny = torch.zeros(1, device=dev)
    if fused:
        gather = sharded.FusedGridGather(ctx, rank, world, n_total, S)     # peers map each other's full arrays
        full_ptr, full_numel = gather.full_ptr, gather.numel()
    else:
        send = torch.empty(n_local * S, dtype=torch.float32, device=dev)
        recv = torch.empty(world * n_local * S, dtype=torch.float32, device=dev) if world > 1 else send
        ctx.grid_set_costs_buffer(send.data_ptr(), send.numel())
        full_ptr, full_numel = recv.data_ptr(), recv.numel()
    pts_pinned = torch.from_numpy(inp["pts"]).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ctx.scan_set_points_ptr(pts_pinned.data_ptr(), 32, 16, pts_pinned.shape[0])
    ctx.scan_render_polar(res, float(ANG_RES), N_THETA, N_R, want=False)
    ctx.grid_costs(centers, 2.0, res, shifts, want=False)                  # uploads the centres once
    ctx.sync()
    ctx.profile_enable(True)

    key_t = gather.key_tensor(torch, dev) if fused else None
    checked = [False]

    def one_step():
        with torch.cuda.stream(stream):
            ctx.scan_render_polar(res, float(ANG_RES), N_THETA, N_R, want=False)
            ctx.grid_run_resident(n_local, 2.0, res, shifts)
            if fused and args.grid_collective == "fused":
                # the cross-rank step inside the library: MIN of the packed (cost, global flat index) key + barrier, by
                # system-scope atomics on a mailbox behind every rank's peer-mapped array; after it every rank's array
                # holds every cost (peer stores of all ranks done).  Returns the reduced key (synchronises).
                out = ctx.grid_key_decode(ctx.grid_peer_exchange())
            elif fused:
                # "fused-nccl": the same with ONE NCCL collective, a MIN all-reduce of the key, as barrier + reduction
                dist.all_reduce(key_t, op=dist.ReduceOp.MIN)
                out = ctx.grid_key_decode(int(key_t.item()))               # D2H of the key: synchronises
            elif world > 1:
                dist.all_gather_into_tensor(recv, send)
                out = ctx.grid_best_dev(full_ptr, full_numel)              # D2H of (cost, index): synchronises
            else:
                out = ctx.grid_key_decode(ctx.grid_best_key())             # folded by the score kernel; D2H synchronises
            if not checked[0] and not os.environ.get("TDR_GRID_SELF_ONLY"):   # first (warm-up) step: same answer as a re-scan
                checked[0] = True
                ref = ctx.grid_best_dev(full_ptr, full_numel)
                assert ref == out, (ref, out)
            return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    evs, kern_ms, best = [], [], None
    sampler = ClockSampler(local)
    sampler.start()
    t_load0 = time.perf_counter()
    n_warm, total_steps, i = args.warmup, args.warmup + args.steps, -1
    while True:
        i += 1
        if i == n_warm and more_warmup(t_load0, world, local) and n_warm < args.warmup + 5000:
            n_warm += 1
            total_steps += 1
        if i >= total_steps:
            break
        if i == n_warm:
            barrier()
            launches0 = ctx.launch_count()
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        best = one_step()
        e1.record(stream)
        if i >= n_warm:
            evs.append((e0, e1))
            kern_ms.append(ctx.profile_stage_ms()[1])
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    total_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    e2e_t = []
    for i in range(args.warmup + args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        barrier()
        t0 = time.perf_counter()
        ctx.scan_set_points_ptr(pts_pinned.data_ptr(), 32, 16, pts_pinned.shape[0])
        one_step()
        if i >= args.warmup:
            e2e_t.append(time.perf_counter() - t0)
    e2e_total = float(np.sum(e2e_t))
    if world > 1:
        t = torch.tensor([total_ms, e2e_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0].item()), float(t[1].item())
    verification = None
    if with_verify and rank == 0:
        # a sample of this rank's cost rows against the oracle (outside the timed region)
        arm = CpuArm(wl, inp)
        rows = np.sort(np.random.default_rng(SEED).choice(n_local, 1024, replace=False))
        got = np.stack([ctx.copy_from_device(full_ptr + 4 * (lo + int(r)) * S, S) for r in rows])
        scan_o = arm.orc.render_polar(inp["pts"], res, ANG_RES, N_THETA, N_R, inp["lut"], C)
        want = arm.orc.cost_grid(centers[rows], 2.0, arm.fp, arm.layers, arm.mask, 1.0, arm.tab, N_THETA, N_R, scan_o, res, shifts,
                                 n_threads=arm.cores)
        both = np.isnan(got) & np.isnan(want)
        err = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want.astype(np.float64)), 1e-300)
        err[both] = 0
        err[np.isnan(got) ^ np.isnan(want)] = np.inf
        verification = {"sample_centres": int(len(rows)), "costs_max_rel_err": float(err.max()),
                        "verified": bool(err.max() <= 1e-5)}
    out = None
    if rank == 0:
        peak, peak_src = peaks()
        scores = n_total * S
        k_ms = float(np.mean(kern_ms))
        achieved = n_local * b_score(C) / (k_ms * 1e-3) / 1e9
        out = {"metric": "particle_scores_per_sec", "value": scores * args.steps / (total_ms * 1e-3), "unit": "scores/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": wl["desc"], "centres_total": n_total, "centres_per_gpu": n_local, "shifts": S,
                          "map_px": [side, side], "classes": C, "polar_image": [N_THETA, N_R], "res_m_per_bin": res,
                          "parallelism": f"centre shards x{world}, map replicated"
                          + (((", weight all-gather fused into the score kernel (NVLink peer stores) + arg-min/barrier by "
                               "system-scope atomics on peer mailboxes (tdr_grid_peer_exchange), no NCCL in the step"
                               if args.grid_collective == "fused" else
                               ", weight all-gather fused into the score kernel (NVLink peer stores) + 1 NCCL MIN all-reduce/step")
                              if fused else ", 1 NCCL all-gather(costs)/step") if world > 1 else ""),
                          "l2": "flushed (256 MiB write) between steps, outside the timed events"},
               "clocks": clocks, "best": {"cost": best[0], "flat_index": best[1]},
               "stage_ms": {"score": k_ms},
               "e2e": {"value": scores * args.steps / e2e_total, "unit": "scores/s",
                       "h2d_bytes_per_step": int(inp["pts"].nbytes), "d2h_bytes_per_step": 12,
                       "ms_per_step": 1e3 * e2e_total / args.steps,
                       "timer": "host wall clock around set_points+render+grid(+all-gather)+arg-min"},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": "k_score_mma (tcgen05 gather-GEMM, sliding scan ring, operands in tensor memory)",
                            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": grid_traffic(n_local), "traffic_source": "profiles/ ncu --set full capture, scaled; NOT measured in this run",
                            "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": n_local * b_score(C), "kernel_ms": k_ms},
               "cpu_baseline": None}
        if verification is not None:
            out["verified"] = verification["verified"]
            out["verification"] = verification
    if fused:
        dist.barrier()
        gather.close()
    ctx.close()
    if world > 1:
        dist.barrier()
    return out


_REAL_STDOUT = None


def class_api_record():
    """the same update through the reference's OWN class interface (ParticleFilter::propagate / update / pose members on the
    adapters, tools/class_api_bench.py) — a separate process: the adapters own their device context"""
    so = os.path.join(ROOT, "oracle", "_ref", "libtdr_adapters_gpu.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libtdr_adapters_gpu.so is not built (needs /root/reference at build time)"}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "class_api_bench.py"), "1000000", "5", "both"],
                           capture_output=True, text=True, timeout=300)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not line:
            return {"unavailable": "tools/class_api_bench.py failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200]}
        return json.loads(line[-1])
    except Exception as e:                                      # a missing adapter build must not cost the headline line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_refine(args, wl):
    """cfg5 (a composition): refine_map's binning rule over a batch of recorded scans, then the
    distance fields of the binned class maps.  One step = the WHOLE batch: zero the counters, bin every chunk, rebuild
    the map.  value: chunks resident in HBM (tdr_refine_add_dev); e2e: every chunk copied from pinned host memory
    inside the timed region (tdr_refine_add) + a read of the rebuilt map at one point.  Single GPU."""
    import torch
    from top_down_renderer_b200 import hostmath, synth
    from top_down_renderer_b200.core import Context
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.cuda.set_device(0)
    side, C, res = wl["side"], wl["C"], wl["res"]
    inp = make_inputs(dict(wl, n=16, shifts=1), 0)
    rng = np.random.default_rng(SEED + 5)
    road = synth.road_pixels(inp["cm"])
    local_xy = inp["pts"][:, 0:2].astype(np.float32)                       # metres, sensor frame
    local_cls = inp["pts"][:, 4].astype(np.int32)                          # raw class id (255 = unknown -> dropped)
    keep = ~((local_xy[:, 0] == 0) & (local_xy[:, 1] == 0))
    n_pts = local_xy.shape[0]
    per_chunk = wl["chunk_scans"] * n_pts
    chunks = []
    for _ in range(wl["distinct_chunks"]):                                 # recorded scans: random road poses, world frame (metres)
        xy = np.empty((wl["chunk_scans"], n_pts, 2), dtype=np.float32)
        cl = np.empty((wl["chunk_scans"], n_pts), dtype=np.int32)
        for k in range(wl["chunk_scans"]):
            p = int(road[rng.integers(0, road.size)])
            px, py, th = (p % side) * res, (p // side) * res, float(rng.uniform(-math.pi, math.pi))
            c, s_ = np.float32(math.cos(th)), np.float32(math.sin(th))
            xy[k, :, 0] = c * local_xy[:, 0] - s_ * local_xy[:, 1] + np.float32(px)
            xy[k, :, 1] = s_ * local_xy[:, 0] + c * local_xy[:, 1] + np.float32(py)
            cl[k] = np.where(keep, local_cls, -1)
        chunks.append((torch.from_numpy(xy.reshape(-1, 2)).pin_memory(), torch.from_numpy(cl.reshape(-1)).pin_memory()))
    dev_chunks = [(a.cuda(), b.cuda()) for a, b in chunks]
    n_chunks = (wl["scans"] + wl["chunk_scans"] - 1) // wl["chunk_scans"]
    total_pts = n_chunks * per_chunk
    ctx = Context(0)
    ctx.map_set_polar_table(hostmath.polar_table(N_THETA, N_R, ANG_RES, 1.0), N_THETA, N_R)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    probe = np.array([[side / 2, side / 2]], dtype=np.float32)

    def step(resident):
        ctx.refine_begin(res, 0.0, 0.0, side, side, C)
        for k in range(n_chunks):
            if resident:
                a, b = dev_chunks[k % len(dev_chunks)]
                ctx.refine_add_dev(a.data_ptr(), b.data_ptr(), per_chunk)
            else:
                a, b = chunks[k % len(chunks)]
                ctx.refine_add_ptr(a.data_ptr(), b.data_ptr(), per_chunk)
        ctx.refine_rebuild_map(1.0)

    sampler = ClockSampler(0)
    sampler.start()
    ms, e2e = [], []
    l0 = None
    for i in range(args.warmup + args.steps):
        flush.fill_(i & 0xff)                                               # L2 flush, outside the events
        torch.cuda.synchronize()
        if i == args.warmup:
            l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            step(True)
            e1.record(stream)
        torch.cuda.synchronize()
        if i >= args.warmup:
            ms.append(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    for i in range(max(2, args.steps // 4)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(False)
        ctx.map_local_polar(probe, 2.0, 4.0)                                # D2H read of the rebuilt map: synchronises
        e2e.append(time.perf_counter() - t0)
    clocks = sampler.stop()
    counts = ctx.refine_counts()
    ctx.close()
    t = float(np.mean(ms))
    peak, peak_src = peaks()
    alg_bytes = total_pts * 12 + side * side * (4 * C + 2)                  # 12 B per point in + the EDT's map bytes
    # CPU baseline: the oracle's binning on one chunk (1 thread, like refine_map) + cv2's EDT on the full map, extrapolated
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as orc
        a, b = chunks[0]
        t0 = time.perf_counter()
        orc.refine_bin(a.numpy(), b.numpy(), res, 0.0, 0.0, side, side, C)
        t_bin = time.perf_counter() - t0
        layers = np.ascontiguousarray((counts == 0).astype(np.float32).transpose(0, 2, 1))
        t0 = time.perf_counter()
        orc.compute_dists(layers, 1.0)
        t_edt = time.perf_counter() - t0
        cpu = {"value": total_pts / (t_bin * n_chunks + t_edt), "unit": "points/s", "cores": 1, "kind": "port",
               "sample": f"oracle binning of 1 of {n_chunks} chunks ({per_chunk} points, {t_bin * 1e3:.0f} ms, extrapolated) + oracle "
                         f"distance fields of the full {side}x{side}x{C} map ({t_edt * 1e3:.0f} ms); oracle/tdr_oracle.cpp, 1 thread"}
    out = {"metric": "refine_points_per_sec", "value": total_pts / (t * 1e-3), "unit": "points/s", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32",
           "data": "synthetic",
           "config": {"workload": wl["desc"], "scans": n_chunks * wl["chunk_scans"], "points": total_pts, "map_px": [side, side],
                      "classes": C, "res_m_per_px": res, "chunks": n_chunks,
                      "distinct_scans": wl["distinct_chunks"] * wl["chunk_scans"],
                      "l2": "flushed (256 MiB write) between steps, outside the timed events"},
           "clocks": clocks,
           "e2e": {"value": total_pts / float(np.mean(e2e)), "unit": "points/s", "h2d_bytes_per_step": int(total_pts * 12),
                   "d2h_bytes_per_step": int(N_THETA * N_R * (4 * C + 1)), "ms_per_step": 1e3 * float(np.mean(e2e)),
                   "timer": "host wall clock around begin + all chunks from pinned host memory + rebuild + a read of the map"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "k_refine_bin (warp-aggregated integer red) + EDT rebuild", "achieved": alg_bytes / (t * 1e-3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": alg_bytes / (t * 1e-3) / 1e9 / peak, "traffic": None,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": t},
           "cpu_baseline": cpu}
    emit(out)


def claim_stdout():
    """Rank 0 must print exactly ONE line on stdout, but NCCL / torch write banners to fd 1: point fd 1 at stderr
    for the whole run and keep a private handle on the real stdout for the JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    claim_stdout()
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="global", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--grid-collective", default="fused", choices=["fused", "fused-nccl", "nccl"],
                    help="grid workload, N > 1: all-gather fused into the score kernel over peer memory, or NCCL")
    ap.add_argument("--shard-impl", default="library", choices=["library", "torch"],
                    help="particle workloads, N > 1: the sharded filter inside the library (tdr_shard_*) or the torch.distributed harness")
    ap.add_argument("--no-sub", action="store_true", help="default workload only: skip the cfg2 / cfg4 sub-records")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the benchmarked configuration")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = dict(WORKLOADS[args.workload])
    if args.particles:
        wl["n"] = args.particles
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if wl.get("refine"):
        run_refine(args, wl)
        return
    import torch.distributed as dist
    rank, world, local = init_dist()
    verify = not args.no_verify
    if wl.get("grid"):
        out = measure_grid(wl, rank, world, local, args.steps, args.warmup, args.grid_collective, verify)
    else:
        out = measure_particles(args, wl, args.workload, rank, world, local, args.steps, args.warmup, not args.no_cpu, verify)
        if args.workload == "global" and not args.no_sub and not args.particles:
            # the other two configurations BASELINE.json's metric names, as sub-records of the same line:
            # cfg2 = p50 update latency of 10k tracked particles on ONE GPU (rank 0; the others wait),
            # cfg4 = the 1e8-hypothesis exhaustive grid, strong scaling over all ranks (+ rank 0 alone as the 1-GPU reference)
            trk = None
            if rank == 0:
                t = measure_particles(args, dict(WORKLOADS["tracking"]), "tracking", 0, 1, local, 200, 10, False, verify)
                trk = {"workload": t["config"]["workload"], "p50_update_ms": t["p50_update_ms"], "ms_per_step": t["ms_per_step"],
                       "stage_ms": t["stage_ms"], "e2e_p50_ms": t["e2e"]["p50_ms"], "value": t["value"], "unit": t["unit"],
                       "steps": 200, "verified": t.get("verified"), "verification": t.get("verification"),
                       "roofline_frac": t["roofline"]["frac"]}
            if world > 1:
                dist.barrier()
            g = measure_grid(dict(WORKLOADS["grid"]), rank, world, local, 10, 3, args.grid_collective, verify)
            g1 = None
            if world > 1:
                if rank == 0:
                    g1 = measure_grid(dict(WORKLOADS["grid"]), 0, 1, local, 10, 3, args.grid_collective, False)
                dist.barrier()
            if rank == 0:
                out["tracking"] = trk
                out["grid"] = {"workload": g["config"]["workload"], "value": g["value"], "unit": g["unit"], "ms_per_step": g["ms_per_step"],
                               "scaling": "strong", "n_gpus": world, "steps": 10, "score_kernel_ms": g["stage_ms"]["score"],
                               "e2e_ms_per_step": g["e2e"]["ms_per_step"], "best": g["best"], "verified": g.get("verified"),
                               "verification": g.get("verification"), "roofline_frac": g["roofline"]["frac"],
                               "vs_1gpu": (g["value"] / g1["value"]) if g1 else 1.0,
                               "ms_per_step_1gpu": g1["ms_per_step"] if g1 else g["ms_per_step"],
                               "collective": g["config"]["parallelism"]}
    if rank == 0 and out is not None and args.workload == "global" and not args.no_sub and not args.particles:
        out["e2e_class_api"] = class_api_record()
    if rank == 0 and out is not None:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
