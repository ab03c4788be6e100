"""Generates the committed golden fixtures (run in the build container; outputs are small .npz files).

    python tests/golden/make_golden.py

* edt_cv2.npz      — OpenCV's own answer (cv2.distanceTransform DIST_L2 / DIST_MASK_PRECISE, IPP off, *= resolution,
                     THRESH_TRUNC 50, masked) for a small multi-class map: the REAL third-party routine the
                     reference calls at top_down_map.cpp:312-317.  The reference itself has no tests or vectors
                      and cannot be built here (F1), so this is the only externally pinned stage.
* kat_small.npz    — hand-checkable known-answer cases (tiny map / 4x2 polar image / 5 particles), values derived
                     in the comments of tests/test_oracle.py.
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


if __name__ == "__main__":
    edt_cv2()
    cfg1_mini()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
