"""e2e through the reference's OWN class interface: ParticleFilter::propagate + ParticleFilter::update + the pose members,
as TopDownRender::updateFilter calls them (top_down_render.cpp:413-431), on the adapters (adapters/*.cpp over the C ABI,
built against the reference's unchanged headers with stand-in Eigen / PCL / OpenCV types) with libtdr_b200 behind them.
BASELINE cfg3 shape: N particles (default 1e6) on a 4000 x 4000 px, 6-class map, 40-shift theta search on the first update.

    python tools/class_api_bench.py [N] [steps]        prints one JSON object
Host wall clock per step; the host mirror of the particle set is NOT read inside the loop (the node's visualize would:
28 B per particle and step).  Two propagate variants: the reference's RNG stream (4 host variates per particle from the
shared std::mt19937, then 16 B / particle H2D) and TDR_ADAPTER_DEVICE_RNG=1 (no host loop over the particles)."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, steps, device_rng):
    os.environ["TDR_ADAPTER_DEVICE_RNG"] = "1" if device_rng else "0"
    import bench
    from oracle import oracle as orc
    from oracle import refbuild as ref
    wl = dict(bench.WORKLOADS["global"], n=16)
    inp = bench.make_inputs(wl, 0)
    C_ = wl["C"]
    scan = orc.render_polar(inp["pts"], wl["res"], bench.ANG_RES, bench.N_THETA, bench.N_R, inp["lut"], C_)   # input of update()
    with ref.using_adapters("gpu"):
        side = wl["side"]
        m = ref.Map.from_class_image(inp["img"], inp["lut"], C_, 1.0, center=(side // 2, side // 2))
        m.polar_table(bench.N_THETA, bench.N_R, bench.ANG_RES)
        t0 = time.perf_counter()
        f = ref.Filter(m, n, 1234, regularization=0.7, pos_cov=0.15, theta_cov=0.004, fixed_scale=2.0)
        t_init = time.perf_counter() - t0
        ts = {"propagate": [], "update": [], "pose": []}
        for i in range(steps + 2):
            t0 = time.perf_counter()
            f.propagate(0.4, 0.05, 0.01)
            t1 = time.perf_counter()
            f.update(scan, wl["res"])
            t2 = time.perf_counter()
            mean, cov, ml, cov_ml = f.pose()                 # meanLikelihood, computeMeanCov, maxLikelihood, computeCov: D2H, synchronises
            t3 = time.perf_counter()
            if i >= 2:
                ts["propagate"].append(t1 - t0); ts["update"].append(t2 - t1); ts["pose"].append(t3 - t2)
        n_now = f.num_particles()
    tot = np.array(ts["propagate"]) + np.array(ts["update"]) + np.array(ts["pose"])
    return {"particles": n, "particles_after": n_now, "steps": steps, "initialize_s": t_init,
            "ms_per_step": 1e3 * float(np.mean(tot)), "p50_ms": 1e3 * float(np.median(tot)),
            "propagate_ms": 1e3 * float(np.mean(ts["propagate"])), "update_ms": 1e3 * float(np.mean(ts["update"])),
            "pose_ms": 1e3 * float(np.mean(ts["pose"])), "mean_pose": [float(v) for v in mean]}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    out = {"interface": "ParticleFilter::propagate / update / meanLikelihood / computeMeanCov / maxLikelihood / computeCov "
                        "(adapters over libtdr_b200, reference headers unchanged)",
           "note": "steps after the first run the tracking kernel (headings found by the first update's 40-shift search); "
                   "the adaptive particle count (particle_filter.cpp:151-158) shrinks the set step by step"}
    if which == "both":
        # TDR_ADAPTER_DEVICE_RNG is read once per process: one child per variant
        import subprocess
        for key, arg in (("reference_rng_stream", "host"), ("device_rng", "device")):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(n), str(steps), arg], capture_output=True, text=True)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            out[key] = json.loads(line[-1])[key] if r.returncode == 0 and line else {"failed": (r.stderr.strip().splitlines() or ["?"])[-1][:200]}
    elif which == "host":
        out["reference_rng_stream"] = run(n, steps, False)
    else:
        out["device_rng"] = run(n, steps, True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
