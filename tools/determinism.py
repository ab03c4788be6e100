"""Repeat-call determinism of the tensor-core score kernels (a race shows up as run-to-run differences):
grid costs under every ring-kernel configuration, and the particle theta search under each kernel."""
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
def ndiff(x, y):
    return int(np.count_nonzero(x.view(np.uint32) != y.view(np.uint32)))
for cfg in sys.argv[1:] or ["413", "412", "112", "114", "12", "22", "14"]:
    os.environ["TDR_MMA_RING_CFG"] = cfg
    c = make_ctx(wd)
    c.set_score_impl(2)
    c.scan_set_polar_images(wd.scan)
    runs = [c.grid_costs(centers, 2.0, 4.0, shifts) for _ in range(5)]
    c.close()
    print("grid ring cfg", cfg, "diffs vs run 0:", [ndiff(runs[0], r) for r in runs[1:]], flush=True)
del os.environ["TDR_MMA_RING_CFG"]
st, ld = synth.particles_global(200_000, wd.class_map, seed=5)
for kern in ["1", "2", "3"]:
    os.environ["TDR_MMA_KERNEL"] = kern
    c = make_ctx(wd)
    c.set_score_impl(2)
    c.scan_set_polar_images(wd.scan)
    out = []
    for _ in range(5):
        c.pf_set_states(st, ld)
        out.append(c.pf_score(4.0))
    c.close()
    print("search kernel", kern, "diffs vs run 0:", [ndiff(out[0], r) for r in out[1:]], flush=True)
