"""One GPU running ONE rank's share of cfg4 (the 1e6-centre x 100-heading grid cut into `--ranks` shards): per-step device
time of render + score for the shard, to see how the per-rank kernel of the 8-GPU run behaves without any collective.
    python tools/grid_shard_probe.py --ranks 8 [--steps 20]          (run it under ncu for the launch list)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from top_down_renderer_b200 import hostmath, sharded, synth
    from top_down_renderer_b200.core import Context
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    wl = dict(bench.WORKLOADS["grid"])
    side, C, res = wl["side"], wl["C"], wl["res"]
    inp = bench.make_inputs(dict(wl, n=16, shifts=1), 0)
    centers_all = synth.grid_centers(side, side, 4)
    lo, hi = sharded.shard_range(centers_all.shape[0], a.rank, a.ranks)
    centers = np.ascontiguousarray(centers_all[lo:hi])
    shifts = np.arange(bench.N_THETA, dtype=np.int32)
    ctx = Context(0)
    ctx.map_set_class_image(inp["img"], inp["lut"], C, 1.0)
    ctx.map_set_polar_table(hostmath.polar_table(bench.N_THETA, bench.N_R, bench.ANG_RES, 1.0), bench.N_THETA, bench.N_R)
    ctx.scan_set_lut(inp["lut"], C)
    ctx.pf_set_params(C, regularization=0.7)
    ctx.scan_set_points(inp["pts"])
    ctx.scan_render_polar(res, float(bench.ANG_RES), bench.N_THETA, bench.N_R, want=False)
    ctx.grid_costs(centers, 2.0, res, shifts, want=False)
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    ts = []
    for i in range(a.steps + 3):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.scan_render_polar(res, float(bench.ANG_RES), bench.N_THETA, bench.N_R, want=False)
            ctx.grid_run_resident(hi - lo, 2.0, res, shifts)
            e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    print(f"ranks {a.ranks}: {hi - lo} centres, {np.mean(ts):.4f} ms/step (min {np.min(ts):.4f}); ideal from 1 rank = t1/{a.ranks}")
    ctx.close()


if __name__ == "__main__":
    main()
