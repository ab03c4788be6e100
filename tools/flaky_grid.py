import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.common import make_ctx, make_world
from top_down_renderer_b200 import synth
wd = make_world(h=2000, w=2000, C=6, seed=21)
centers = synth.grid_centers(wd.h, wd.w, 4)
per_row = len(np.arange(2, wd.w, 4))
centers = np.ascontiguousarray(centers[per_row * 20: per_row * 20 + 102_400])
shifts = np.arange(100, dtype=np.int32)
for rep in range(6):
    c = make_ctx(wd)
    c.set_score_impl(2)
    c.scan_set_polar_images(wd.scan)
    a = c.grid_costs(centers, 2.0, 4.0, shifts)
    a2 = c.grid_costs(centers, 2.0, 4.0, shifts)
    os.environ["TDR_GRID_PHASE_LOG2"] = "0"
    b = c.grid_costs(centers, 2.0, 4.0, shifts)
    b2 = c.grid_costs(centers, 2.0, 4.0, shifts)
    del os.environ["TDR_GRID_PHASE_LOG2"]
    c.close()
    def diff(x, y):
        d = np.argwhere(x.view(np.uint32) != y.view(np.uint32))
        return len(d), (d[:4].tolist() if len(d) else [])
    print(rep, "split vs split", diff(a, a2), "plain vs plain", diff(b, b2), "split vs plain", diff(a, b), flush=True)
    if diff(a, b)[0]:
        d = np.argwhere(a.view(np.uint32) != b.view(np.uint32))
        rows = np.unique(d[:, 0])
        print("   rows", len(rows), rows[:10], "cols", np.unique(d[:, 1])[:20], "vals", a[d[0][0], d[0][1]], b[d[0][0], d[0][1]])
from oracle import oracle as orc
from tests.common import N_THETA, N_R
c = make_ctx(wd)
c.set_score_impl(2)
c.scan_set_polar_images(wd.scan)
runs = [c.grid_costs(centers, 2.0, 4.0, shifts) for _ in range(4)]
c.close()
for k in range(1, 4):
    d = np.argwhere(runs[0].view(np.uint32) != runs[k].view(np.uint32))
    print("run", k, "differs from run 0 in", len(d), "entries,", len(np.unique(d[:, 0])) if len(d) else 0, "rows")
d = np.argwhere(runs[0].view(np.uint32) != runs[1].view(np.uint32))
if len(d):
    rows = np.unique(d[:, 0])[:6]
    want = orc.cost_grid(centers[rows], 2.0, wd.fp, wd.layers, wd.mask, 1.0, wd.tab, N_THETA, N_R, wd.scan, 4.0, shifts)
    for j, r in enumerate(rows):
        cols = d[d[:, 0] == r][:, 1]
        print("row", r, "centre", centers[r], "n cols", len(cols), "first cols", cols[:6])
        for cc in cols[:3]:
            print("    col", cc, "run0 %.9g run1 %.9g oracle %.9g" % (runs[0][r, cc], runs[1][r, cc], want[j, cc]))
    print("rows mod 128:", (np.unique(d[:, 0]) % 128)[:40])
    print("row // 128 :", np.unique(np.unique(d[:, 0]) // 128)[:40])
