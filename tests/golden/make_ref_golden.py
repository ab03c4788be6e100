"""Generates tests/golden/ref_mini.npz FROM THE REFERENCE'S OWN SOURCE: every array in it was computed by the reference's
classes as built in oracle/_ref (seven translation units compiled unmodified from /root/reference/src against the
stand-in headers of oracle/ref_shim/, see its README.md), none by the oracle.  Needs /root/reference, i.e. runs only in
the build container; the fixture it writes is committed and read by tests/test_oracle.py (oracle vs reference) and
tests/test_gpu_parity.py (CUDA vs reference) wherever they run.

    python tests/golden/make_ref_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc            # noqa: E402  (only for the polar offset table, an INPUT, and the engine bookkeeping)
from oracle import refbuild as ref          # noqa: E402
from top_down_renderer_b200 import synth    # noqa: E402

SEED, N, C_, H, W, RES = 2718, 160, 4, 160, 200, 2.0
ANG = np.float32(2 * math.pi / 100)
KW = dict(fixed_scale=2.0, init_pos_px_cov=7.0, init_pos_deg_cov=5.0)
FILTER = dict(regularization=0.7, pos_cov=0.15, theta_cov=0.004)
MOTION = (0.4, 0.05, 0.01)


def main():
    cm = synth.make_class_map(H, W, C_, seed=SEED)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C_)
    pose, heading = synth.default_pose(cm, seed=SEED)
    pts = synth.make_scan(cm, pose, heading, seed=SEED, n_rings=16, n_az=256)
    m = ref.Map.from_class_image(img, lut, C_, 1.0, center=(W // 2, H // 2))
    layers, mask = m.get()
    tab = m.polar_table(100, 25, ANG)                                   # the reference's samplePtsPolar (libm cos / sin)
    scan = ref.render_polar(pts, RES, ANG, 100, 25, lut, C_)
    cart = ref.render_cart(pts, RES, 40, 56, lut, C_)
    centres = np.float32([[pose[0], pose[1]], [3.5, 150.25], [-30.0, 40.0]])
    local = [m.local_map_polar(float(c[0]), float(c[1]), 2.0, RES) for c in centres]
    kw = dict(KW, init_pos_px=(float(pose[0]), float(pose[1])), init_pos_deg_theta=math.degrees(heading))
    f = ref.Filter(m, N, SEED, **FILTER, **kw)
    st0, _, _ = f.get()
    peek0 = f.engine_peek()
    f.propagate(*MOTION)
    st1, ld1, _ = f.get()
    peek1 = f.engine_peek()
    gmm_samples, _, covs = f.gmm()
    f.update(scan, RES)
    scored, ld_s, raw = f.get(scored_set=True)
    wn = f.weights()
    cur, _, _ = f.get()
    peek2 = f.engine_peek()
    mean, cov, ml, cov_ml = f.pose()
    # a second filter: no heading -> the 40-candidate theta search, a few particles moved onto unknown ground / off the map
    g = ref.Filter(m, 96, SEED + 1, regularization=0.7, fixed_scale=2.0, force_on_map=True)
    s0, _, _ = g.get()
    s0["init_x_px"][::9] = -120.0
    s0["dx_m"][1::11] = 1e4
    s0["init_x_px"][2::13] = 1.0                                          # a map corner: three quarters of the footprint unknown -> NaN
    s0["init_y_px"][2::13] = 1.0
    s0["have_init"][2::26] = 1                                            # with a heading the weight is NaN; while searching, every
    s0["theta"][2::26] = 0.3                                              # candidate is NaN and the weight becomes 1 / (FLT_MAX + reg)
    ld = np.random.default_rng(SEED).uniform(0, 0.4, len(s0)).astype(np.float32)
    g.set(s0, ld)
    g.update(scan, RES)
    s_scored, _, s_raw = g.get(scored_set=True)
    s_wn = g.weights()
    np.savez_compressed(os.path.join(HERE, "ref_mini.npz"), img=img, lut=lut, num_classes=C_, pts=pts, res=np.float32(RES), ang_res=ANG,
                        layers=layers, mask=mask, tab=tab, scan=scan, cart=cart, centres=centres,
                        local_d=np.stack([d for d, _ in local]), local_m=np.stack([k for _, k in local]),
                        init_px=np.float32(kw["init_pos_px"]), init_theta_deg=np.float32(kw["init_pos_deg_theta"]),
                        init_states=st0, propagated=st1, last_dist=ld1, engine_peek=np.uint32([peek0, peek1, peek2]),
                        gmm_samples=gmm_samples, gmm_cov=covs[0], scored=scored, raw=raw, weights_norm=wn, resampled=cur,
                        mean=mean, cov=cov, ml=ml, cov_ml=cov_ml,
                        search_in=s0, search_last_dist=ld, search_scored=s_scored, search_raw=s_raw, search_weights_norm=s_wn)
    print("ref_mini.npz", os.path.getsize(os.path.join(HERE, "ref_mini.npz")), "bytes;", len(cur), "particles after the update;",
          int(np.isnan(s_raw).sum()), "NaN and", int((s_raw == 0).sum()), "gated weights in the search set")


if __name__ == "__main__":
    main()
