import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = dict(bench.WORKLOADS["global"], n=1_000_000)
inp = bench.make_inputs(wl, 0)
ctx = bench.setup_ctx(wl, inp, 0)
ctx.profile_enable(True)
ctx.scan_set_points(inp["pts"])
n = wl["n"]
ctx.step(4.0, float(bench.ANG_RES), 100, 25, 0.37, n); ctx.sync()
print("search step", ctx.profile_stage_ms())
for i in range(3):
    t0 = time.perf_counter()
    ctx.step(4.0, float(bench.ANG_RES), 100, 25, 0.37, n); ctx.sync()
    print("tracking step", ctx.profile_stage_ms(), "wall %.2f ms" % (1e3 * (time.perf_counter() - t0)))
for i in range(2):
    t0 = time.perf_counter(); ctx.pf_score(4.0, want=False); ctx.sync(); t1 = time.perf_counter()
    ctx.pf_normalize(); t2 = time.perf_counter()
    ctx.pf_resample(0.3, n, want=False); ctx.sync(); t3 = time.perf_counter()
    ctx.pf_pose(); t4 = time.perf_counter()
    print("score %.2f normalize %.2f resample %.2f pose %.2f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)))
