"""Small launches of every hand-rolled mbarrier / TMEM pipeline for compute-sanitizer (tools/run_sanitizer.sh):
the streamed-operand theta-search kernel and the ring kernel with SEVERAL batches per CTA (grid capped at 2 CTAs), the
ring kernel on a lattice of centres (phase-split map, asynchronous bulk-store epilogue), k_score_track, the fused
small update and the tiled normalise / resample.  No oracle here: parity is pytest's job, this only has to execute."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from top_down_renderer_b200 import hostmath, synth          # noqa: E402
from top_down_renderer_b200.core import Context             # noqa: E402


def main():
    C, H, W = 6, 320, 320
    cm = synth.make_class_map(H, W, C, seed=3)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C)
    ang = np.float32(2 * math.pi / 100)
    pose, heading = synth.default_pose(cm, seed=3)
    pts = synth.make_scan(cm, pose, heading, seed=3)[::4]
    tab = hostmath.polar_table(100, 25, ang, 1.0)
    thetas, shifts = hostmath.search_list(100)
    os.environ["TDR_MMA_GRID_CAP"] = "2"
    for kernel in ("1", "2"):
        os.environ["TDR_MMA_KERNEL"] = kernel
        ctx = Context(0)
        ctx.map_set_class_image(img, lut, C, 1.0)
        ctx.map_set_polar_table(tab, 100, 25)
        ctx.scan_set_lut(lut, C)
        ctx.pf_set_params(C, regularization=0.7)
        ctx.pf_set_search(thetas, shifts)
        ctx.scan_set_points(pts)
        ctx.scan_render_polar(4.0, ang, 100, 25)
        ctx.set_score_impl(2)
        st, ld = synth.particles_global(1500, cm, seed=4)
        st["have_init"][::5] = 1
        ctx.pf_set_states(st, ld)
        ctx.pf_update(4.0, 0.37, len(st))                  # search (6 batches per CTA) + track + fused small update
        ctx.sync()
        if kernel == "2":
            centers = synth.grid_centers(H, W, 4)[:900]
            ctx.grid_costs(centers, 2.0, 4.0, np.arange(100, dtype=np.int32))
            ctx.grid_best_key()
            st, ld = synth.particles_tracking(40000, pose, heading, seed=5)
            ctx.pf_set_states(st, ld)
            ctx.pf_update(4.0, 0.61, len(st))              # k_score_track + tiled normalise / prefix / resample
            ctx.pf_pose()
        ctx.close()
    print("sanitize_case done")


if __name__ == "__main__":
    main()
