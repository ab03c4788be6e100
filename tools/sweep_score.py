"""Sweep tuning knobs of the cfg3 score stage on one GPU: one context per setting (the knobs are read at tdr_create),
same inputs, stage timers.   python tools/sweep_score.py KNOB=v1,v2,... [KNOB2=...] [--particles N]"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    knobs, n = [], 1_000_000
    for a in sys.argv[1:]:
        if a.startswith("--particles="):
            n = int(a.split("=")[1])
        else:
            k, v = a.split("=")
            knobs.append((k, v.split(",")))
    wl = dict(bench.WORKLOADS["global"], n=n)
    inp = bench.make_inputs(wl, 0)
    u = 0.37
    for combo in itertools.product(*[v for _, v in knobs]):
        for (k, _), v in zip(knobs, combo):
            os.environ[k] = v
        ctx = bench.setup_ctx(wl, inp, 0)
        ctx.profile_enable(True)
        ctx.scan_set_points(inp["pts"])
        ms = []
        for i in range(6):
            ctx.pf_restore()
            ctx.step(wl["res"], float(bench.ANG_RES), bench.N_THETA, bench.N_R, u, n)
            ctx.sync()
            if i >= 2:
                ms.append(ctx.profile_stage_ms())
        ms = np.array(ms).mean(0)
        print(" ".join(f"{k}={v}" for (k, _), v in zip(knobs, combo)), "score %.3f ms  (render %.3f norm %.3f resample %.3f)" % (ms[1], ms[0], ms[2], ms[3]), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
