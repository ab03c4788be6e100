"""Generates the committed golden fixtures (run in the build container; outputs are small .npz files).

    python tests/golden/make_golden.py

* edt_cv2.npz      — OpenCV's own answer (cv2.distanceTransform DIST_L2 / DIST_MASK_PRECISE, IPP off, *= resolution,
                     THRESH_TRUNC 50, masked) for a small multi-class map: the REAL third-party routine the
                     reference calls at top_down_map.cpp:312-317.  The reference itself has no tests or vectors
                      and cannot be built here (F1), so this is the only externally pinned stage.
* kat_small.npz    — hand-checkable known-answer cases (tiny map / 4x2 polar image / 5 particles), values derived
                     in the comments of tests/test_oracle.py.
* widen_mini.npz   — oracle outputs for the rows added beyond the first path: propagate with the reference's
                     shared mt19937 + libstdc++ normal_distribution (states, last_dist and the standard variates behind the
                     draws), and the vector-map path (polygons -> binary class layers); regression pins read by the CPU
                     and the GPU tests.
* cfg1_mini.npz    — oracle outputs on a seeded cfg1-style world (regression pin; the GPU tests compare the CUDA
                     path with the same file, so the fixture travels to the GPU box where /root/reference and
                     cv2 need not exist).
"""
import math
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402
from top_down_renderer_b200 import synth  # noqa: E402


def edt_cv2():
    import cv2
    # OpenCV's OWN exact transform (trueDistTrans): IPP is switched off because pip's cv2 routes images under
    # 16384 px to ippiTrueDistanceTransform, whose sqrt is up to 1 ulp off; distro / ROS OpenCV builds (what the
    # reference links, CMakeLists.txt:29) carry no IPP, and maps of realistic size never take that path anyway.
    cv2.ipp.setUseIPP(False)
    H, W, Cn = 72, 96, 3
    cm = synth.make_class_map(H, W, Cn + 1, seed=42)       # classes 0..3, 255 unknown
    cm[cm == 3] = 2
    img, lut = synth.to_cv_image(cm), synth.identity_lut(Cn)
    out = {}
    for resolution in (1.0, 0.5, 2.0):
        bl = orc.class_image_to_layers(img, lut, Cn, resolution)       # (C, cols, rows) of 0/1
        msum = bl.astype(np.uint8).sum(axis=0)
        mask = (msum > Cn - 1)
        layers = np.empty_like(bl)
        for c in range(Cn):
            b8 = bl[c].astype(np.uint8)                                 # the cv::Mat(cols, rows) view of the layer
            d = cv2.distanceTransform(b8, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
            d = d * np.float32(resolution)
            _, d = cv2.threshold(d, 50, 0, cv2.THRESH_TRUNC)
            d[mask] = 0
            layers[c] = d
        out[f"layers_{resolution}"] = layers.astype(np.float32)
        out[f"mask_{resolution}"] = mask.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "edt_cv2.npz"), img=img, lut=lut, num_classes=Cn, **out)


def cfg1_mini():
    Cn, H, W, n = 4, 240, 320, 96
    cm = synth.make_class_map(H, W, Cn, seed=77)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(Cn)
    ang = np.float32(2 * math.pi / 100)
    pose, heading = synth.default_pose(cm, seed=77)
    pts = synth.make_scan(cm, pose, heading, seed=77, n_rings=16, n_az=256)
    st, ld = synth.particles_tracking(n, pose, heading, seed=77)
    st["have_init"][::2] = 0
    tab = orc.polar_table(100, 25, ang, 1.0)
    thetas, shifts = orc.search_list(100)
    u = orc.uniform_draw(77)
    layers, mask = orc.compute_dists(orc.class_image_to_layers(img, lut, Cn, 1.0), 1.0)
    scan = orc.render_polar(pts, 2.0, ang, 100, 25, lut, Cn)
    fp = orc.make_params(Cn, regularization=0.7, map_width=W, map_height=H)
    st_o = st.copy()
    w = orc.score_all(st_o, fp, layers, mask, 1.0, tab, 100, 25, scan, 2.0, thetas, shifts)
    wn, arg, stats = orc.normalize(w, ld)
    idx = orc.resample_fast(wn, u, n)
    new = st_o[idx]
    mean, cov = orc.mean_cov(new)
    ml, _ = orc.ml_cov(st_o, arg)
    np.savez_compressed(os.path.join(HERE, "cfg1_mini.npz"), img=img, lut=lut, num_classes=Cn, pts=pts, states=st,
                        last_dist=ld, tab=tab, thetas=thetas, shifts=shifts, u=np.float32(u), res=np.float32(2.0),
                        ang_res=ang, scan=scan, layers_crc=np.uint32(zlib.crc32(layers.tobytes())),
                        mask_crc=np.uint32(zlib.crc32(mask.tobytes())), weights=w, theta_out=st_o["theta"], weights_norm=wn, argmax=arg,
                        idx=idx, mean=mean, cov=cov, ml=ml)


def widen_mini():
    rng = np.random.default_rng(5)
    st, _ = synth.particles_tracking(256, (120.0, 90.0), 0.4, seed=5)
    out = {}
    for freeze in (0, 1):
        a, last, z = orc.propagate(st, 0.7, -0.2, 0.03, bool(freeze), 0.3, 0.1, 1234)
        out[f"prop_states_{freeze}"], out[f"prop_last_{freeze}"], out[f"prop_z_{freeze}"] = a, last, z
    polys, cls = [], []
    for _ in range(24):
        k = int(rng.integers(3, 9))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(4, 30) * rng.uniform(0.4, 1.0, k)
        cx, cy = rng.uniform(0, 96), rng.uniform(0, 72)
        polys.append(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1).astype(np.float32))
        cls.append(int(rng.integers(0, 3)))
    start = np.zeros(len(polys) + 1, dtype=np.int32)
    start[1:] = np.cumsum([len(p) for p in polys])
    excl = np.array([0, 0, 0, 0, 1], dtype=np.int32)
    layers = orc.raster_polygons(polys, cls, 96, 72, 0.0, 1.0, 3, excl)
    np.savez_compressed(os.path.join(HERE, "widen_mini.npz"), prop_in=st, prop_args=np.float32([0.7, -0.2, 0.03, 0.3, 0.1]),
                        poly_verts=np.concatenate(polys), poly_start=start, poly_class=np.int32(cls), poly_excl=excl,
                        poly_layers=layers.astype(np.uint8), **out)


INIT_CASES = {"free": dict(fixed_scale=-1.0),
              "pixel": dict(fixed_scale=2.0, init_pos_px=(101.5, 88.25), init_pos_px_cov=9.0, init_pos_deg_theta=75.0, init_pos_deg_cov=8.0),
              "metric": dict(fixed_scale=2.0, init_pos_m=(-9.25, 4.125), init_pos_px_cov=5.0)}


def init_mini():
    """particle initialisation on the shared engine (particle_filter.cpp:19-84): states per case + engine outputs used"""
    Cn, H, W = 4, 180, 240
    cm = synth.make_class_map(H, W, Cn, seed=31)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(Cn)
    layers, _ = orc.compute_dists(orc.class_image_to_layers(img, lut, Cn, 1.0), 1.0)
    out = {}
    road = np.argwhere(cm == synth.ROAD)                                   # (y, x) rows of the class map
    y0, x0 = (int(v) for v in road[np.argmin(np.hypot(road[:, 0] - H * 0.4, road[:, 1] - W * 0.6))])
    cases = {k: dict(v) for k, v in INIT_CASES.items()}
    cases["metric"]["init_pos_m"] = ((x0 - W // 2) / 2.0, (y0 - H // 2) / 2.0)
    cases["pixel"]["init_pos_px"] = (x0 + 0.5, y0 + 0.25)
    out["metric_m"], out["pixel_px_in"] = np.float32(cases["metric"]["init_pos_m"]), np.float32(cases["pixel"]["init_pos_px"])
    for name, kw in cases.items():
        st, frozen, px, used = orc.init_particles(2024, layers, 1.0, (W // 2, H // 2), 70, **kw)
        assert len(st) == 70, name
        out[f"{name}_states"], out[f"{name}_used"], out[f"{name}_px"] = st, np.int64(used), np.float32(px)
        out[f"{name}_u"] = np.float32(orc.uniform_draw(2024, discard=used))
    np.savez_compressed(os.path.join(HERE, "init_mini.npz"), img=img, lut=lut, num_classes=Cn, **out)


if __name__ == "__main__":
    edt_cv2()
    cfg1_mini()
    widen_mini()
    init_mini()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
