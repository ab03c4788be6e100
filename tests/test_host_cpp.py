"""The C++ host mirror of the reference's class interfaces (top_down_renderer_b200/host/tdr_host.hpp):
it compiles and links against the C ABI on the CPU box (and refuses to run there), and on a GPU one scan step driven
through ScanRendererPolar / TopDownMapPolar / ParticleFilter agrees with the oracle.  The same demo linked against a
CPU stand-in of the C ABI that answers every call with the oracle (tests/cpp/tdr_cpu_standin.cpp, test code only) puts
the mirror's HOST logic — RNG streams, map-centre shift, freezeScale, cache files — under the CPU suite and proves the
checks below on a run the oracle produced."""
import math
import os
import subprocess

import numpy as np
import pytest

import top_down_renderer_b200 as tdr
from oracle import oracle as orc
from top_down_renderer_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "cpp", "host_demo")


def build_demo():
    tdr.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_demo.cpp")
    hdrs = [os.path.join(ROOT, "top_down_renderer_b200", "host", f) for f in ("tdr_host.hpp", "png_gray.hpp")]
    if not os.path.exists(DEMO) or os.path.getmtime(DEMO) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", DEMO, src, "-L", os.path.join(ROOT, "top_down_renderer_b200"),
                               "-l:libtdr_b200.so", "-Wl,-rpath," + os.path.join(ROOT, "top_down_renderer_b200"), "-lz"])
    return DEMO


DEMO_CPU = os.path.join(ROOT, "tests", "cpp", "host_demo_cpu")


def build_demo_cpu():
    """host_demo.cpp + the CPU stand-in of the C ABI + liboracle.so: no CUDA anywhere in this binary"""
    so = orc.build()
    srcs = [os.path.join(ROOT, "tests", "cpp", f) for f in ("host_demo.cpp", "tdr_cpu_standin.cpp")]
    deps = srcs + [os.path.join(ROOT, "top_down_renderer_b200", "host", f) for f in ("tdr_host.hpp", "png_gray.hpp")] + [os.path.join(ROOT, "include", "tdr.h"), so]
    if not os.path.exists(DEMO_CPU) or os.path.getmtime(DEMO_CPU) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", DEMO_CPU] + srcs +
                              ["-L", os.path.dirname(so), "-l:" + os.path.basename(so), "-Wl,-rpath," + os.path.dirname(so), "-pthread", "-lz"])
    return DEMO_CPU


def write_inputs(d, N=800, seed=5):
    C, H, W = 4, 400, 480
    cm = synth.make_class_map(H, W, C, seed=seed)
    img = synth.to_cv_image(cm)
    pose, heading = synth.default_pose(cm, seed=seed)
    pts = synth.make_scan(cm, pose, heading, seed=seed, n_rings=32, n_az=512)
    np.array([H, W, C, N, seed, 100, 25], dtype=np.int32).tofile(os.path.join(d, "meta.i32"))
    np.array([2.0, pose[0], pose[1], 6.0, math.degrees(heading), 3.0], dtype=np.float32).tofile(os.path.join(d, "meta.f32"))
    img.tofile(os.path.join(d, "class_image.u8"))
    pts.tofile(os.path.join(d, "points.f32"))
    polys, cls = demo_polygons()
    np.concatenate(polys).astype(np.float32).tofile(os.path.join(d, "polys.f32"))
    np.cumsum([0] + [len(q) for q in polys]).astype(np.int32).tofile(os.path.join(d, "poly_start.i32"))
    np.array(cls, dtype=np.int32).tofile(os.path.join(d, "poly_class.i32"))
    np.array([VEC_W, VEC_H] + VEC_EXCLUSIVE, dtype=np.int32).tofile(os.path.join(d, "vec_meta.i32"))
    return cm, img, pts


VEC_W, VEC_H, VEC_EXCLUSIVE = 230, 170, [0, 1]


def demo_polygons():
    """a small vector map in class order (setVectorMap hands the polygons over class by class): a background, a concave
    road, an axis-aligned building, overlapping vegetation, one polygon partly off the map"""
    rng = np.random.default_rng(12)
    polys = [np.float32([[-5, -5], [240, -5], [240, 180], [-5, 180]]),                                   # class 0 everywhere
             np.float32([[20.3, 30.1], [200.7, 35.2], [205.1, 60.8], [120.5, 55.5], [110.2, 140.9], [80.8, 139.3], [85.4, 52.6], [18.9, 58.2]]),
             np.float32([[150.25, 90.25], [190.75, 90.25], [190.75, 130.75], [150.25, 130.75]]),
             np.float32([[60, 100], [100, 70], [140, 110], [95, 160]]) + rng.uniform(-0.4, 0.4, (4, 2)).astype(np.float32),
             np.float32([[200, 120], [260, 125], [250, 200], [190, 190]])]
    return polys, [0, 1, 2, 3, 3]


def test_host_mirror_compiles_and_refuses_to_run_without_gpu(tmp_path):
    demo = build_demo()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    write_inputs(str(tmp_path), N=16)
    r = subprocess.run([demo, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def rd(d, name, dtype):
    return np.fromfile(os.path.join(d, name), dtype=dtype)


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def check_demo_outputs(d, cm, img, pts, N, seed, device):
    """what host_demo wrote against the oracle.  device=True: the run came from libtdr_b200 (weights within 1e-5,
    propagate bit-exact against the numpy twin that forms cos / sin like the kernel); False: from the CPU stand-in, where
    every number is the oracle's own and everything is bit-exact."""
    from oracle import numpy_twin as twin
    from top_down_renderer_b200 import eigcache
    C, H, W = 4, cm.shape[0], cm.shape[1]
    lut = synth.identity_lut(C)
    ang = np.float32(2 * math.pi / 100)
    fmeta = rd(d, "meta.f32", np.float32)
    # class images: bit-exact
    scan_o = orc.render_polar(pts, 2.0, ang, 100, 25, lut, C)
    scan = np.stack([rd(d, f"scan_{c}.f32", np.float32).reshape(25, 100) for c in range(C)])
    assert np.array_equal(scan, scan_o)
    layers, mask = orc.compute_dists(orc.class_image_to_layers(img, lut, C, 1.0), 1.0)
    tab = orc.polar_table(100, 25, ang, 1.0)
    # the Cartesian twins through the base classes: bit-exact
    cart = np.stack([rd(d, f"cart_scan_{c}.f32", np.float32) for c in range(C)])
    assert np.array_equal(cart, orc.render_cart(pts, 1.5, 48, 64, lut, C).reshape(C, -1)) and cart.sum() > 1000
    want_c, want_cm = orc.local_map_cart(layers, mask, 1.0, float(fmeta[1]), float(fmeta[2]), 0.3, 2.5, 30, 40)
    got_c = np.stack([rd(d, f"cart_local_{c}.f32", np.float32) for c in range(C)])
    assert same_bits(got_c, want_c.reshape(C, -1)) and np.array_equal(rd(d, "cart_mask.u8", np.uint8), want_cm.reshape(-1))

    # ---- initializeParticles + propagate: the reference's own RNG calls on ONE engine (particle_filter.cpp:19-92) ----
    init_kw = dict(init_pos_px=(float(fmeta[1]), float(fmeta[2])), init_pos_px_cov=float(fmeta[3]), init_pos_deg_theta=float(fmeta[4]),
                   init_pos_deg_cov=float(fmeta[5]), fixed_scale=2.0)
    init, frozen, _, used = orc.init_particles(seed, layers, 1.0, (W // 2, H // 2), N, **init_kw)
    assert len(init) == N and frozen and (init["have_init"] == 1).all() and (init["scale"] == 2.0).all()
    want, last_o, z, used_p = orc.propagate(init, 0.4, 0.05, 0.01, True, 0.15, 0.004, seed, discard=used)
    st = rd(d, "states_before.bin", synth.STATE_DTYPE)
    ld = rd(d, "last_dist.f32", np.float32)
    assert len(st) == N
    for k in ("init_x_px", "init_y_px", "scale"):                 # untouched by propagate: the rejection sampler's output
        assert same_bits(st[k], init[k]), k
    assert np.array_equal(st["have_init"], init["have_init"])
    ref, last_r = (twin.propagate_with_z(init, 0.4, 0.05, 0.01, True, 0.15, 0.004, z) if device else (want, last_o))
    for k in ("dx_m", "dy_m", "theta"):
        assert same_bits(st[k], ref[k]), k
    assert same_bits(ld, last_r)
    assert np.abs(st["dx_m"] - want["dx_m"]).max() <= 1e-6 and np.abs(st["dy_m"] - want["dy_m"]).max() <= 1e-6
    # getClassesAtPoint tests the DISTANCE layer (< 1, top_down_map.cpp:166), and unknown pixels have every layer
    # zeroed (:317), so the reference also accepts unknown pixels as "on the road" — mirrored, not fixed
    at = cm[st["init_y_px"].astype(int), st["init_x_px"].astype(int)]
    assert np.isin(at, [synth.ROAD, synth.UNKNOWN]).all() and (at == synth.ROAD).mean() > 0.5
    assert (ld > 0).all()
    # the ONE uniform of update (:172-173) is the engine's next output
    u = float(rd(d, "u.f32", np.float32)[0])
    assert u == orc.uniform_draw(seed, discard=used + used_p)

    # ---- update: weights within 1e-5 of the oracle's on the same particle set, resampled states consistent ----
    thetas, shifts = orc.search_list(100)
    fp = orc.make_params(C, regularization=0.7, map_width=W, map_height=H)
    w = orc.score_all(st.copy(), fp, layers, mask, 1.0, tab, 100, 25, scan_o, 2.0, thetas, shifts)
    wn, arg, _ = orc.normalize(w, ld)
    got = rd(d, "weights_norm.f32", np.float32)
    assert np.max(np.abs(got - wn) / wn) <= 1e-5
    if not device:
        assert same_bits(got, wn)
    after = rd(d, "states_after.bin", synth.STATE_DTYPE)
    idx = orc.resample_fast(got, u, len(st))                 # stage-wise: the run's own normalised weights
    assert np.array_equal(after, st[idx])
    mean = rd(d, "mean.f32", np.float32)
    wm, _ = orc.mean_cov(after)
    assert abs(mean[0] - wm[0]) <= 0.002 and abs(mean[1] - wm[1]) <= 0.002 and abs(mean[2] - wm[2]) <= math.radians(0.01)
    ml = rd(d, "ml.f32", np.float32)
    s_ml = st[int(np.argmax(got))]                           # maxCoeff: the first maximum of the run's own weights
    want_ml = np.float32([s_ml["dx_m"] * s_ml["scale"] + s_ml["init_x_px"], s_ml["dy_m"] * s_ml["scale"] + s_ml["init_y_px"],
                          s_ml["theta"], s_ml["scale"]])
    assert np.abs(ml - want_ml).max() <= 1e-3

    # ---- the map cache the C++ mirror wrote in the reference's .eig format (and then ran the filter from): read back
    # with the independent Python reader — distance fields and mask to the bit, geo layers like the oracle's ----
    assert eigcache.cache_is_valid(d, "demo_map", C, 1.0) and not eigcache.cache_is_valid(d, "demo_map", C + 1, 1.0)
    c_layers, c_geo, c_mask = eigcache.load_cache(d, C)
    assert same_bits(c_layers, layers) and np.array_equal(c_mask, mask)
    geo_o = orc.geo_raster(orc.class_image_to_layers(img, lut, C, 1.0))
    geo_d, _ = orc.compute_dists(geo_o, 1.0)
    assert same_bits(c_geo, geo_d)

    # ---- ActiveLocalizer::getBestRelPos and getLocalGeoMap through the mirror classes ----
    preds = np.float32([[W * 0.3, H * 0.4, 0.2], [W * 0.6, H * 0.5, -1.1], [W * 0.5, H * 0.7, 2.4]])
    rel_o, _ = orc.active_best_rel_pos(layers, mask, 1.0, tab.reshape(-1), 100, 25, preds)
    rel = rd(d, "active_rel.f32", np.float32)
    assert (float(rel[0]), float(rel[1])) == rel_o and rel_o[0] >= 50
    g0 = rd(d, "geo_local0.f32", np.float32)
    want_g, _ = orc.local_map_polar(geo_d, np.zeros_like(mask), 1.0, tab, np.float32(W * 0.5), np.float32(H * 0.5), 1.0, 2.0)
    assert same_bits(g0, want_g[0])

    # ---- the rest of the ParticleFilter interface: pure host logic over the resident states ----
    sc_fixed, sc_free, sc_frozen, n_metric, n_off, n_ad1, n_st1, n_ad2, n_st2 = (float(v) for v in rd(d, "misc.f32", np.float32))
    assert sc_fixed == 2.0 and sc_free == -1.0                                  # scale(): particle_filter.cpp:358-366
    # (a) updateMap: every init position moves by the map-centre delta (:325-333) — (W/2, H/2) from the initial (0, 0)
    # on the first map message, then (+3, -2); float += int
    sh0, _, _, _ = orc.init_particles(seed + 1, layers, 1.0, (W // 2, H // 2), 64, **init_kw)
    b, a = rd(d, "shift_before.bin", synth.STATE_DTYPE), rd(d, "shift_after.bin", synth.STATE_DTYPE)
    f = np.float32
    assert same_bits(b["init_x_px"], sh0["init_x_px"] + f(W // 2)) and same_bits(b["init_y_px"], sh0["init_y_px"] + f(H // 2))
    assert same_bits(a["init_x_px"], b["init_x_px"] + f(3)) and same_bits(a["init_y_px"], b["init_y_px"] + f(-2))
    for k in ("dx_m", "dy_m", "theta", "scale"):
        assert same_bits(a[k], sh0[k]), k
    # (b) free scale: 12 prototypes x 10 scales 10^(0, 0.1, ...), no heading; propagate jitters the scale; freezeScale
    # locks the geometric mean (float accumulator over double pow, :343-357)
    fr0, frozen_f, _, used_f = orc.init_particles(seed + 2, layers, 1.0, (W // 2 + 3, H // 2 - 2), 120, fixed_scale=-1.0)
    fb = rd(d, "free_before.bin", synth.STATE_DTYPE)
    assert not frozen_f and len(fr0) == 120 and np.array_equal(fb, fr0)
    assert (fb["have_init"] == 0).all() and np.allclose(fb["scale"][:10], 10.0 ** (np.arange(10) / 10), rtol=1e-6)
    assert (fb["init_x_px"].reshape(12, 10) == fb["init_x_px"].reshape(12, 10)[:, :1]).all()   # ten scales per prototype
    fw, _, zf, _ = orc.propagate(fr0, 0.3, -0.1, -0.02, False, 0.15, 0.004, seed + 2, discard=used_f)
    fpg = rd(d, "free_propagated.bin", synth.STATE_DTYPE)
    fref = twin.propagate_with_z(fr0, 0.3, -0.1, -0.02, False, 0.15, 0.004, zf)[0] if device else fw
    for k in ("dx_m", "dy_m", "theta", "scale"):
        assert same_bits(fpg[k], fref[k]), k
    assert not np.array_equal(fpg["scale"], fr0["scale"])
    fz, g = orc.freeze_scale(fpg)
    ffz = rd(d, "free_frozen.bin", synth.STATE_DTYPE)
    assert np.array_equal(ffz, fz) and sc_frozen == g and abs(g - np.exp(np.log(fpg["scale"].astype(np.float64)).mean())) < 1e-4 * g
    # (c) a metric initial position relative to the map centre (:27-54); off the map the filter stays empty
    mc = (W // 2 + 3, H // 2 - 2)
    m_xy = ((fmeta[1] - f(mc[0])) / f(2.0), (fmeta[2] - f(mc[1])) / f(2.0))
    kw = dict(init_kw); kw["init_pos_px"] = (-1.0, -1.0)
    ms, _, px, _ = orc.init_particles(seed + 3, layers, 1.0, mc, 48, init_pos_m=(float(m_xy[0]), float(m_xy[1])), **kw)
    got_m = rd(d, "metric_states.bin", synth.STATE_DTYPE)
    assert n_metric == 48 and np.array_equal(got_m, ms) and abs(px[0] - fmeta[1]) < 1e-3 and abs(px[1] - fmeta[2]) < 1e-3
    assert np.hypot(got_m["init_x_px"] - fmeta[1], got_m["init_y_px"] - fmeta[2]).max() < 6 * fmeta[3]
    assert n_off == 0
    # (d) the GMM drives the particle count (:151-158); the EM input matrix (:262-272) off the device
    cov = np.zeros((1, 4, 4), np.float32)
    cov[0, 0, 0], cov[0, 1, 1] = 100.0, 144.0
    assert n_ad1 == n_st1 == orc.adaptive_count(cov, 300, 300) == 235 and n_ad2 == n_st2 == orc.adaptive_count(cov, 235, 300) == 186
    gst = rd(d, "gmm_states.bin", synth.STATE_DTYPE)
    gsm = rd(d, "gmm_samples.f64", np.float64).reshape(-1, 4)
    want_g = orc.gmm_samples(gst, 300)
    assert len(gst) == 300 and gsm.shape == (300, 4) and np.array_equal(gsm[:, :2], want_g[:, :2])
    assert np.abs(gsm[:, 2:] - want_g[:, 2:]).max() <= (0.0 if not device else 1e-5)

    # ---- vector map -> raster cache -> a second map from the PNG files alone (top_down_map.cpp:22-31, :197-224) ----
    from top_down_renderer_b200 import rastercache
    polys, cls = demo_polygons()
    binl = orc.raster_polygons(polys, cls, VEC_W, VEC_H, 0.0, 1.0, C, VEC_EXCLUSIVE)
    assert all((binl[c] == 0).any() and (binl[c] == 1).any() for c in range(C))
    cache = os.path.join(d, "vec_raster_cache")
    assert np.array_equal(rastercache.load_rasterized_maps(cache, C), binl)      # the C++ writer's files, the Python reader
    try:
        import cv2
        top = cv2.imread(os.path.join(cache, "class2.png"), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(top, rastercache.layer_to_image(binl[2]))           # ... and OpenCV's
    except ImportError:
        pass
    vl, vmask = orc.compute_dists(binl, 1.0)
    want_l, want_m = orc.local_map_polar(vl, vmask, 1.0, tab, np.float32(VEC_W * np.float32(0.45)), np.float32(VEC_H * np.float32(0.55)), 1.5, 2.0)
    vec = np.stack([rd(d, f"vec_local{c}.f32", np.float32) for c in range(C)])
    ras = np.stack([rd(d, f"ras_local{c}.f32", np.float32) for c in range(C)])
    assert same_bits(vec, want_l.reshape(C, -1)) and same_bits(ras, vec)
    assert np.array_equal(rd(d, "ras_mask.u8", np.uint8), want_m.reshape(-1))
    at = rd(d, "vec_classes_at.i32", np.int32)
    assert at[-1] == -1 and list(at[:-1]) == orc.classes_at_point(vl, 1.0, int(VEC_W * np.float32(0.45)), int(VEC_H * np.float32(0.55)))


def test_host_mirror_logic_on_the_cpu_standin(tmp_path):
    """every class of the host mirror driven end to end with the oracle answering the C ABI: the mirror's own code (RNG
    call order, state bookkeeping, file formats) must reproduce the oracle's restatement of the reference bit for bit"""
    demo = build_demo_cpu()
    d = str(tmp_path)
    cm, img, pts = write_inputs(d)
    r = subprocess.run([demo, d], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "No map received for input loc" in r.stderr            # the off-map filter said why it stayed empty
    check_demo_outputs(d, cm, img, pts, 800, 5, device=False)
    # ... and the reference's OWN ParticleFilter (oracle/_ref: its sources compiled against stand-in headers) seeded alike:
    # the mirror's initialisation + propagate reproduce it bit for bit
    from oracle import refbuild as ref
    if ref.available():
        fmeta = rd(d, "meta.f32", np.float32)
        H, W = cm.shape
        rmap = ref.Map.from_class_image(img, synth.identity_lut(4), 4, 1.0, center=(W // 2, H // 2))
        rmap.set_polar_table(orc.polar_table(100, 25, np.float32(2 * math.pi / 100), 1.0), 100, 25)
        f = ref.Filter(rmap, 800, 5, regularization=0.7, pos_cov=0.15, theta_cov=0.004, fixed_scale=2.0,
                       init_pos_px=(float(fmeta[1]), float(fmeta[2])), init_pos_px_cov=float(fmeta[3]),
                       init_pos_deg_theta=float(fmeta[4]), init_pos_deg_cov=float(fmeta[5]))
        f.propagate(0.4, 0.05, 0.01)
        r_st, r_ld, _ = f.get()
        assert np.array_equal(r_st, rd(d, "states_before.bin", synth.STATE_DTYPE)) and same_bits(r_ld, rd(d, "last_dist.f32", np.float32))
        # its update draws the same uniform next
        f.update(orc.render_polar(pts, 2.0, np.float32(2 * math.pi / 100), 100, 25, synth.identity_lut(4), 4), 2.0)
        _, _, r_raw = f.get(scored_set=True)
        r_wn = f.weights()
        assert np.allclose(rd(d, "weights_norm.f32", np.float32), r_wn, rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_host_mirror_step_matches_oracle(tmp_path):
    demo = build_demo()
    d = str(tmp_path)
    cm, img, pts = write_inputs(d)
    r = subprocess.run([demo, d], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    check_demo_outputs(d, cm, img, pts, 800, 5, device=True)
