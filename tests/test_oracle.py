"""The oracle (oracle/tdr_oracle.cpp) against everything that can pin it WITHOUT the reference's own tests
(it has none): OpenCV's real routine (live cv2 and the committed cv2 fixture), a naive numpy
twin, hand-computed known answers, and domain properties."""
import math
import os
import sys
import zlib

import numpy as np
import pytest

from oracle import numpy_twin as twin
from oracle import oracle as orc
from top_down_renderer_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ANG = np.float32(2 * math.pi / 100)


# ---- distance fields (a3/a4) ------------------------------------------------------------------------
def test_edt_matches_committed_cv2_fixture():
    g = np.load(os.path.join(GOLD, "edt_cv2.npz"))
    Cn = int(g["num_classes"])
    for resolution in (1.0, 0.5, 2.0):
        bl = orc.class_image_to_layers(g["img"], g["lut"], Cn, resolution)
        layers, mask = orc.compute_dists(bl, resolution)
        assert np.array_equal(mask, g[f"mask_{resolution}"])
        assert np.array_equal(layers.view(np.uint32), g[f"layers_{resolution}"].view(np.uint32)), resolution


@pytest.mark.parametrize("shape,p", [((200, 300), 1e-3), ((257, 131), 0.2), ((64, 64), 0.0)])
def test_edt_matches_live_cv2(shape, p):
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)      # OpenCV's own trueDistTrans (see test_small_image_ipp_path_is_within_2ulp)
    rng = np.random.default_rng(shape[0])
    b = (rng.random(shape) >= p).astype(np.uint8)          # 0 = seed
    d2 = orc.edt_sq(b)
    got = np.sqrt(d2.astype(np.float32), dtype=np.float32)
    want = cv2.distanceTransform(b, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    if p == 0.0:                                           # no seed: OpenCV returns a large sentinel -> truncates to 50
        assert (want > 50).all() and (np.minimum(got, 50) == 50).all()
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_small_image_ipp_path_is_within_2ulp():
    """pip's cv2 sends images under 16384 px to Intel IPP, whose sqrt is not correctly rounded; the oracle
    follows OpenCV's own code path (bit-exact above), and the IPP deviation is characterised here."""
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "ipp"):
        pytest.skip("no IPP switch in this cv2")
    rng = np.random.default_rng(1)
    b = (rng.random((96, 72)) >= 0.01).astype(np.uint8)
    want = np.sqrt(orc.edt_sq(b).astype(np.float32), dtype=np.float32)
    cv2.ipp.setUseIPP(True)
    got = cv2.distanceTransform(b, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    cv2.ipp.setUseIPP(False)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2


def test_compute_dists_equals_bruteforce_twin():
    cm = synth.make_class_map(40, 56, 4, seed=5)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(4)
    for resolution in (1.0, 0.5):
        bl = orc.class_image_to_layers(img, lut, 4, resolution)
        layers, mask = orc.compute_dists(bl, resolution)
        tl, tm = twin.edt_layers_bruteforce(bl, resolution)
        assert np.array_equal(mask, tm) and np.array_equal(layers.view(np.uint32), tl.view(np.uint32))


def test_edt_known_answer_3x3():
    # class 0 only at the centre, class 1 everywhere else -> no unknown pixel.
    bl = np.ones((2, 3, 3), dtype=np.float32)
    bl[0, 1, 1] = 0
    bl[1] = 0
    bl[1, 1, 1] = 1
    layers, mask = orc.compute_dists(bl, 1.0)
    r2 = np.float32(np.sqrt(np.float32(2)))
    assert np.array_equal(layers[0], np.array([[r2, 1, r2], [1, 0, 1], [r2, 1, r2]], dtype=np.float32))
    assert np.array_equal(layers[1], np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=np.float32))
    assert not mask.any()
    # a pixel with no class at all is unknown: every layer is zeroed there and the mask is 1
    bl[1, 0, 0] = 1
    layers, mask = orc.compute_dists(bl, 1.0)
    assert mask[0, 0] == 1 and mask.sum() == 1 and layers[0, 0, 0] == 0 and layers[1, 0, 0] == 0
    # truncation at 50 and the class-absent layer (top_down_map.cpp:315; OpenCV's no-seed sentinel)
    big = np.ones((2, 1, 120), dtype=np.float32)
    big[0, 0, 0] = 0
    big[1] = 1
    big[1, 0, 1:] = 1
    big[0, 0, 1:] = 1
    layers, mask = orc.compute_dists(np.stack([big[0], np.zeros_like(big[0])]), 1.0)
    assert layers[0, 0, 49] == 49 and layers[0, 0, 50] == 50 and layers[0, 0, 119] == 50


def test_edt_commutes_with_transpose():
    rng = np.random.default_rng(3)
    b = (rng.random((90, 70)) > 0.01).astype(np.uint8)
    assert np.array_equal(orc.edt_sq(b).T, orc.edt_sq(np.ascontiguousarray(b.T)))


def test_class_image_flip_and_lut():
    # image row 0 is the TOP of the map (top_down_map.cpp:137): class at image (row 0, col 2) lands in map row H-1
    img = np.full((4, 5), 255, dtype=np.uint8)
    img[0, 2] = 1
    bl = orc.class_image_to_layers(img, synth.identity_lut(2), 2, 1.0)
    assert bl.shape == (2, 5, 4)
    assert bl[1, 2, 3] == 0 and bl[1].sum() == 19 and bl[0].sum() == 20


# ---- rasterisers (a1/a2) ----------------------------------------------------------------------------
def test_polar_render_known_answers():
    lut = synth.identity_lut(2)
    pts = np.zeros((6, 8), dtype=np.float32)
    pts[0, :2] = (0, 4)        # theta = atan2(x=0, y=4) = 0      -> bin 50, r bin 1
    pts[1, :2] = (4, 0)        # +pi/2 -> 25 bins                 -> bin 75
    pts[2, :2] = (-4, 0)       # -pi/2                            -> bin 25
    pts[3, :2] = (0, -4)       # +pi -> bin 100: dropped (scan_renderer_polar.cpp:102)
    pts[4, :2] = (-1e-4, -4)   # just below -pi... rounds to -50  -> bin 0
    pts[5, :2] = (0, 0)        # invalid return, skipped (:95)
    pts[:, 4] = 1
    img = orc.render_polar(pts, 4.0, ANG, 100, 25, lut, 2)
    nz = {(int(c), int(r), int(t)): int(img[c, r, t]) for c, r, t in np.argwhere(img)}
    assert nz == {(1, 1, 50): 1, (1, 1, 75): 1, (1, 1, 25): 1, (1, 1, 0): 1}


def test_render_polar_equals_twin_and_counts_points():
    cm = synth.make_class_map(300, 300, 4, seed=9)
    pose, heading = synth.default_pose(cm, seed=9)
    pts = synth.make_scan(cm, pose, heading, seed=9, n_rings=16, n_az=128)
    lut = synth.identity_lut(4)
    for res in (4.0, 0.7):
        a = orc.render_polar(pts, res, ANG, 100, 25, lut, 4)
        b = twin.render_polar(pts, res, ANG, 100, 25, lut, 4)
        assert np.array_equal(a, b)
    # property: the images sum to the number of valid, in-range, known-class points
    assert a.sum() == b.sum() > 0
    cart = orc.render_cart(pts, 1.0, 64, 48, lut, 4)
    x, y = pts[:, 0], pts[:, 1]
    xi = np.array([twin._libm.roundf(float(v)) for v in x]) + 24
    yi = np.array([twin._libm.roundf(float(v)) for v in y]) + 32
    ok = ~((x == 0) & (y == 0)) & (xi >= 0) & (xi < 48) & (yi >= 0) & (yi < 64) & (pts[:, 4] < 4)
    assert cart.sum() == ok.sum()


# ---- gather + shift-correlation (a6/a7/a10) -----------------------------------------------------------
@pytest.fixture(scope="module")
def small_world():
    cm = synth.make_class_map(260, 300, 4, seed=21)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(4)
    layers, mask = orc.compute_dists(orc.class_image_to_layers(img, lut, 4, 1.0), 1.0)
    pose, heading = synth.default_pose(cm, seed=21)
    pts = synth.make_scan(cm, pose, heading, seed=21, n_rings=32, n_az=256)
    scan = orc.render_polar(pts, 1.5, ANG, 100, 25, lut, 4)
    tab = orc.polar_table(100, 25, ANG, 1.0)
    return dict(cm=cm, layers=layers, mask=mask, pose=pose, heading=heading, scan=scan, tab=tab, lut=lut)


def test_local_map_equals_twin(small_world):
    w = small_world
    for cx, cy in [(150.3, 120.8), (2.0, 3.0), (299.6, 259.4), (-40.0, 100.0)]:
        d, m = orc.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], cx, cy, 2.0, 1.5)
        td, tm = twin.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], cx, cy, 2.0, 1.5)
        assert np.array_equal(m, tm) and np.array_equal(d.view(np.uint32), td.view(np.uint32))
    # fully off the map: everything masked, zeros
    d, m = orc.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], -5000.0, -5000.0, 2.0, 1.5)
    assert m.all() and not d.any()


def test_shift_correlation_is_a_roll(small_world):
    w = small_world
    d, m = orc.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], w["pose"][0], w["pose"][1], 2.0, 1.5)
    known = (1 - m).astype(np.float32)
    cw = np.array([1.0, 0.5, 2.0, 1.25], dtype=np.float32)
    for s in (0, 1, 3, 50, 98, 99):
        a = orc.cost_for_shift(w["scan"], d, known, 100, 25, cw, s)
        b = twin.cost_for_shift(w["scan"], d, known, 100, 25, cw, s)
        assert abs(a - b) <= 1e-5 * abs(b), (s, a, b)


def test_cost_known_answer_4x2():
    # 4 angular x 2 radial bins, one class, one scan count at (theta=1, r=1): cost(s) = 0.01*w*m[(1-s)%4, 1]
    scan = np.zeros((1, 2, 4), dtype=np.float32)
    scan[0, 1, 1] = 3
    m = np.arange(8, dtype=np.float32).reshape(1, 2, 4) + 1          # map value at (r, theta) = 1 + 4r + theta
    known = np.ones(8, dtype=np.float32)
    for s in range(4):
        want = 0.01 * 2.0 * 3 * m[0, 1, (1 - s) % 4] / 3.0
        got = orc.cost_for_shift(scan, m.reshape(1, 8), known, 4, 2, np.array([2.0], dtype=np.float32), s)
        assert abs(got - want) < 1e-6 * want
    known[:5] = 0                                                     # < half known -> NaN (state_particle.cpp:117-120)
    assert math.isnan(orc.cost_for_shift(scan, m.reshape(1, 8), known, 4, 2, np.array([2.0], dtype=np.float32), 0))


def test_search_picks_first_strict_minimum(small_world):
    w = small_world
    thetas, shifts = orc.search_list(100)
    fp = orc.make_params(4, regularization=0.7, map_width=300, map_height=260)
    st = np.zeros(3, dtype=synth.STATE_DTYPE)
    st["init_x_px"], st["init_y_px"], st["scale"] = [w["pose"][0], 40.0, -900.0], [w["pose"][1], 200.0, 50.0], 2.0
    st0 = st.copy()
    wt = orc.score_all(st, fp, w["layers"], w["mask"], 1.0, w["tab"], 100, 25, w["scan"], 1.5, thetas, shifts)
    for i in range(3):
        d, m = orc.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], st0["init_x_px"][i], st0["init_y_px"][i], 2.0, 1.5)
        costs = [twin.cost_for_shift(w["scan"], d, (1 - m).astype(np.float32), 100, 25, np.ones(4), s) for s in shifts]
        if np.all(np.isnan(costs)):
            assert st["theta"][i] == 0 and wt[i] == np.float32(1.0 / (np.float32(3.402823466e38) + np.float32(0.7)))
            assert 0 < wt[i] < 1.2e-38                                  # the denormal weight (Appendix A.5)
        else:
            k = int(np.nanargmin(costs))
            assert abs(wt[i] - 1.0 / (costs[k] + 0.7)) <= 1e-5 * wt[i]
        assert st["have_init"][i] == 1


# ---- normalise / resample / pose (a11-a13) --------------------------------------------------------------
def test_normalize_known_answer():
    w = np.array([1, 2, 3, np.nan], dtype=np.float32)
    wn, arg, stats = orc.normalize(w, np.full(4, 0.2, dtype=np.float32))        # d = min(0.2*5, 1) = 1
    assert np.allclose(wn, np.array([1, 2, 3, 1]) / 7.0, rtol=1e-6) and arg == 2
    assert stats[0] == 6 and stats[1] == 3 and stats[2] == 2 and stats[3] == 1   # sum, num_valid, mean, lower std
    wn, arg, _ = orc.normalize(w, np.zeros(4, dtype=np.float32))                 # d = 0 -> uniform
    assert np.array_equal(wn, np.full(4, 0.25, dtype=np.float32)) and arg == 0
    wn, _, stats = orc.normalize(np.zeros(5, dtype=np.float32), np.ones(5, dtype=np.float32))
    assert stats[5] == 1 and np.array_equal(wn, np.full(5, 0.2, dtype=np.float32))  # all-ones fallback (:129-131)


@pytest.mark.parametrize("n", [1, 3, 4, 7, 8, 9, 64, 1001])
def test_normalize_equals_twin(n):
    rng = np.random.default_rng(n)
    w = (1.0 / (rng.random(n) * 2 + 0.7)).astype(np.float32)
    if n > 3:
        w[rng.integers(0, n, n // 4)] = np.nan
    ld = rng.uniform(0, 0.4, n).astype(np.float32)
    a, arg, _ = orc.normalize(w, ld)
    b, targ = twin.normalize(w, ld)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and arg == targ


def test_resample_known_answer_and_properties():
    w = np.array([0.1, 0.2, 0.3, 0.4], dtype=np.float32)
    assert list(orc.resample_literal(w, 0.5, 4)) == [1, 2, 3, 3]
    assert list(orc.resample_fast(w, 0.5, 4)) == [1, 2, 3, 3]
    rng = np.random.default_rng(8)
    for n, M in [(1, 5), (50, 50), (300, 123), (257, 1024)]:
        w = rng.random(n).astype(np.float32)
        w /= w.sum()
        u = orc.uniform_draw(n)
        assert 0 <= u < 1
        a, b, c = orc.resample_literal(w, u, M), orc.resample_fast(w, u, M), twin.resample_literal(w, u, M)
        assert np.array_equal(a, b) and np.array_equal(a, c)
        assert (np.diff(a) >= 0).all() and a.max() <= n - 1
    # weights that sum to less than the last sample: the index clamps to N-1 (:180)
    w = np.full(10, 0.05, dtype=np.float32)
    assert orc.resample_fast(w, 0.9, 10)[-1] == 9 and orc.resample_literal(w, 0.9, 10)[-1] == 9


def test_uniform_draw_is_libstdcxx_mt19937():
    # mt19937(seed)() first output for seed 5489 is 3499211612; generate_canonical<float,24> = that / 2^32
    assert abs(orc.uniform_draw(5489) - 3499211612 / 2 ** 32) < 1e-7


def test_pose_known_answer():
    st = np.zeros(4, dtype=synth.STATE_DTYPE)
    st["init_x_px"], st["init_y_px"] = [10, 20, 30, 40], [1, 1, 3, 3]
    st["dx_m"], st["scale"] = 1.0, 2.0                      # ml x = dx*scale + init_x
    st["theta"] = [0.1, -0.1, 0.1, -0.1]
    mean, cov = orc.mean_cov(st)
    assert mean[0] == 27 and mean[1] == 2 and abs(mean[2]) < 1e-7 and mean[3] == 2
    assert abs(cov[0, 0] - 500 / 3) < 1e-4 and abs(cov[1, 1] - 4 / 3) < 1e-6 and cov[3, 3] == 0
    ml, _ = orc.ml_cov(st, 2)
    assert list(ml) == [32, 3, np.float32(0.1), 2]


# ---- regression pin ---------------------------------------------------------------------------------
def test_cfg1_mini_fixture_reproduces():
    g = np.load(os.path.join(GOLD, "cfg1_mini.npz"))
    Cn = int(g["num_classes"])
    layers, mask = orc.compute_dists(orc.class_image_to_layers(g["img"], g["lut"], Cn, 1.0), 1.0)
    assert zlib.crc32(layers.tobytes()) == int(g["layers_crc"]) and zlib.crc32(mask.tobytes()) == int(g["mask_crc"])
    scan = orc.render_polar(g["pts"], float(g["res"]), g["ang_res"], 100, 25, g["lut"], Cn)
    assert np.array_equal(scan, g["scan"])
    assert np.array_equal(orc.polar_table(100, 25, g["ang_res"], 1.0), g["tab"])
    st = g["states"].copy()
    fp = orc.make_params(Cn, regularization=0.7, map_width=layers.shape[1], map_height=layers.shape[2])
    w = orc.score_all(st, fp, layers, mask, 1.0, g["tab"], 100, 25, scan, float(g["res"]), g["thetas"], g["shifts"])
    assert np.array_equal(w.view(np.uint32), g["weights"].view(np.uint32))
    wn, arg, _ = orc.normalize(w, g["last_dist"])
    assert np.array_equal(wn.view(np.uint32), g["weights_norm"].view(np.uint32)) and arg == int(g["argmax"])
    assert np.array_equal(orc.resample_fast(wn, float(g["u"]), len(st)), g["idx"])


# ---- propagate (state_particle.cpp:57-78)
def _ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("freeze", [False, True])
def test_propagate_literal_rng_equals_injected_variates(freeze):
    """the literal restatement (one shared mt19937, libstdc++ normal_distribution objects built exactly as the
    reference builds them) and the twin that applies the recorded STANDARD variates with `z * stddev + mean` agree:
    scale bit for bit, positions / heading up to the 1-ulp differences between glibc's cosf / sinf (CPU-dispatched,
    FMA or SSE2 variant) and a correctly rounded cos / sin — the tolerance the device path is held to as well."""
    st, _ = synth.particles_tracking(20000, (150.0, 150.0), 0.6)
    tx, ty, omega = 0.7, -0.2, 0.03
    a, last, z = orc.propagate(st, tx, ty, omega, freeze, 0.3, 0.1, 99)
    b, last_b = twin.propagate_with_z(st, tx, ty, omega, freeze, 0.3, 0.1, z)
    assert np.array_equal(a["scale"].view(np.uint32), b["scale"].view(np.uint32))
    if freeze:
        assert np.array_equal(a["scale"], st["scale"]) and (z[:, 3] == 0).all()
    # 1 ulp of cos / sin (<= 1.2e-7) on each product, carried through two fp32 additions (an ulp of the result each)
    tol = 2.5e-7 * (abs(tx) + abs(ty)) + 2.5e-7 * max(np.abs(a["dx_m"]).max(), np.abs(a["dy_m"]).max(), 1.0)
    assert np.abs(a["dx_m"] - b["dx_m"]).max() <= tol and np.abs(a["dy_m"] - b["dy_m"]).max() <= tol
    assert _ulps(a["theta"], b["theta"]).max() <= 1
    assert np.abs(last - last_b).max() <= 1e-6
    # the variates are standard normal, and the stream is the reference's: re-running reproduces it
    assert abs(z[:, :3].mean()) < 0.02 and abs(z[:, :3].std() - 1) < 0.02
    a2, _, z2 = orc.propagate(st, tx, ty, omega, freeze, 0.3, 0.1, 99)
    assert np.array_equal(z, z2) and np.array_equal(a2, a)


def test_propagate_known_answer_without_noise():
    """pos_cov = theta_cov = 0 and a frozen scale: pure rotation of the odometry step (hand-computed)"""
    st = np.zeros(2, dtype=synth.STATE_DTYPE)
    st["theta"] = [0.0, math.pi / 2]
    st["scale"] = 2.0
    st["dx_m"] = [1.0, 1.0]
    a, last, _ = orc.propagate(st, 3.0, 4.0, 0.25, True, 0.0, 0.0, 1)
    assert np.allclose(a["dx_m"], [4.0, -3.0], atol=1e-6) and np.allclose(a["dy_m"], [4.0, 3.0], atol=1e-6)
    assert np.allclose(a["theta"], [0.25, math.pi / 2 + 0.25], atol=1e-6) and np.allclose(last, [5.0, 5.0], atol=1e-6)
    assert np.array_equal(a["scale"], st["scale"])


# ---- vector map -> class layers (top_down_map.cpp:328-408)
def test_raster_polygons_known_answers():
    """sample point of pixel (row r, col c) is (x, y) = (c + 0.5, r + 0.5) at resolution 1; a layer is 0 inside a polygon of
    its class; overlapping polygons of one class are a union (max over polygons), a polygon traced twice in one path
    is a hole (even-odd); exclusive classes: the lower class yields where a higher one is present"""
    sq = np.array([[2, 2], [6, 2], [6, 5], [2, 5]], np.float32)
    out = orc.raster_polygons([sq], [0], 10, 8, 0.0, 1.0, 2, [])
    lay = out[0].T                                           # rows x cols
    want = np.ones((8, 10), np.float32)
    want[2:5, 2:6] = 0
    assert np.array_equal(lay, want) and (out[1] == 1).all()
    # union of two overlapping squares of the same class, and a second class on top with exclusivity
    sq2 = sq + np.float32([2, 1])
    out = orc.raster_polygons([sq, sq2, sq2], [0, 0, 1], 10, 8, 0.0, 1.0, 2, [0, 1])
    want0 = want.copy(); want0[3:6, 4:8] = 0
    want1 = np.ones((8, 10), np.float32); want1[3:6, 4:8] = 0
    assert np.array_equal(out[1].T, want1)
    assert np.array_equal(out[0].T, np.minimum(want0 + (1 - want1), 1))        # class 0 removed under class 1
    # a ring: outer square and inner square in ONE path -> the inner part is outside (even-odd)
    ring = np.array([[1, 1], [9, 1], [9, 7], [1, 7], [1, 1], [3, 3], [3, 5], [7, 5], [7, 3], [3, 3]], np.float32)
    out = orc.raster_polygons([ring], [0], 10, 8, 0.0, 1.0, 1, [])
    lay = out[0].T
    assert lay[1, 1] == 0 and lay[6, 8] == 0 and lay[3, 3] == 1 and lay[4, 6] == 1 and lay[0, 0] == 1
    # half resolution: 2 px per unit -> sample points at 0.25, 0.75, ...
    out = orc.raster_polygons([sq], [0], 10, 8, 0.0, 0.5, 1, [])
    assert out.shape == (1, 20, 16) and out[0].T[4:10, 4:12].sum() == 0 and out[0].sum() == 20 * 16 - 6 * 8


def test_widen_mini_fixture_reproduces():
    """tests/golden/widen_mini.npz (make_golden.py): propagate and the vector-map path.  The polygon layers involve no
    transcendental and must match to the bit; the propagate stream goes through glibc's logf / cosf / sinf, whose
    CPU-dispatched variants may differ in the last place between machines, hence 2 ulp on the variates."""
    g = np.load(os.path.join(GOLD, "widen_mini.npz"))
    tx, ty, omega, pos_cov, theta_cov = (float(v) for v in g["prop_args"])
    for freeze in (0, 1):
        a, last, z = orc.propagate(g["prop_in"], tx, ty, omega, bool(freeze), pos_cov, theta_cov, 1234)
        assert _ulps(z, g[f"prop_z_{freeze}"]).max() <= 2
        want = g[f"prop_states_{freeze}"]
        for k in ("dx_m", "dy_m", "theta", "scale"):
            assert np.allclose(a[k], want[k], rtol=2e-6, atol=2e-6), k
        assert np.allclose(last, g[f"prop_last_{freeze}"], rtol=1e-5, atol=2e-6)
    st = g["poly_start"]
    polys = [g["poly_verts"][st[k]:st[k + 1]] for k in range(len(st) - 1)]
    lay = orc.raster_polygons(polys, g["poly_class"], 96, 72, 0.0, 1.0, 3, g["poly_excl"])
    assert np.array_equal(lay.astype(np.uint8), g["poly_layers"]) and set(np.unique(lay)) <= {0.0, 1.0}


# ---- adaptive particle count and the GMM sample matrix
def test_adaptive_count_known_answers():
    """num_particles_ = clamp(sum_k floor(sqrt(l0_k) * sqrt(l1_k)), 3 * last / 4 + 10, max) (particle_filter.cpp:151-158)"""
    cov = np.zeros((2, 4, 4), np.float32)
    cov[0, 0, 0], cov[0, 1, 1] = 400.0, 900.0                      # eigenvalues 400, 900 -> 20 * 30 = 600
    cov[1, :2, :2] = [[50.0, 30.0], [30.0, 50.0]]                  # eigenvalues 20, 80 -> sqrt(1600) = 40
    assert orc.adaptive_count(cov, 100, 100000) == 640
    assert orc.adaptive_count(cov, 2000, 100000) == 1510           # lower bound 3 * 2000 / 4 + 10
    assert orc.adaptive_count(cov, 100, 500) == 500                # upper bound
    rot = np.zeros((1, 4, 4), np.float32)
    rot[0, :2, :2] = [[4.0, -9.0], [9.0, 4.0]]                     # complex pair 4 +- 9i: real parts 4 -> 2 * 2
    assert orc.adaptive_count(rot, 0, 100) == 10                   # 4 < 0 * 3 / 4 + 10


def test_gmm_samples_stride_and_encoding():
    st, _ = synth.particles_tracking(2345, (100.0, 80.0), 0.5, seed=4)
    s = orc.gmm_samples(st, 1000)
    idx = np.minimum(len(st) - 1, np.arange(1000) * len(st) // 1000)
    x = st["dx_m"][idx] * st["scale"][idx] + st["init_x_px"][idx]
    assert np.array_equal(s[:, 0], x.astype(np.float64))
    assert np.allclose(s[:, 2], 50 * np.cos(st["theta"][idx].astype(np.float64)), atol=1e-5)
    assert np.allclose(np.hypot(s[:, 2], s[:, 3]), 50.0, atol=1e-4)
    few = orc.gmm_samples(st[:7], 7)
    assert np.array_equal(few[:, 1], (st["dy_m"][:7] * st["scale"][:7] + st["init_y_px"][:7]).astype(np.float64))


def test_raster_polygons_equals_numpy_twin():
    """an independently structured restatement (crossing counts on a 2-D grid, logical or over polygons) gives the same
    layers bit for bit — pins the row / column conventions and the exclusive-class pass of the C++ oracle"""
    rng = np.random.default_rng(23)
    polys, cls = [], []
    for _ in range(40):
        k = int(rng.integers(3, 10))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(3, 40) * rng.uniform(0.3, 1.0, k)
        cx, cy = rng.uniform(-5, 125), rng.uniform(-5, 95)
        p = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1).astype(np.float32)
        polys.append(np.rint(p).astype(np.float32) if rng.random() < 0.4 else p)
        cls.append(int(rng.integers(0, 4)))
    for resolution in (1.0, 0.5):
        for excl in ([], [0, 0, 0, 0, 0, 1, 2], [2, 1, 0]):
            a = orc.raster_polygons(polys, cls, 120, 90, 0.0, resolution, 4, excl)
            b = twin.raster_polygons(polys, cls, 120, 90, resolution, 4, excl)
            assert a.shape == b.shape and np.array_equal(a, b), (resolution, excl)


def test_active_localizer_equals_numpy_twin(small_world):
    """getBestRelPos against the np.roll / float64 twin: same relative position, difference within 1e-5"""
    w = small_world
    rng = np.random.default_rng(3)
    rows, cols = w["layers"].shape[2], w["layers"].shape[1]
    for n in (2, 4):
        preds = np.stack([rng.uniform(0.2, 0.8, n) * cols, rng.uniform(0.2, 0.8, n) * rows, rng.uniform(-3, 3, n)], axis=1).astype(np.float32)
        (d_o, t_o), best_o = orc.active_best_rel_pos(w["layers"], w["mask"], 1.0, w["tab"], 100, 25, preds)
        (d_t, t_t), best_t = twin.active_best_rel_pos(w["layers"], w["mask"], 1.0, w["tab"], 100, 25, preds)
        assert (d_o, t_o) == (d_t, t_t) and abs(best_o - best_t) <= 1e-5 * best_t


# ---- particle initialisation (rejection sampling on the shared engine), freezeScale -------------------
def _states_of(parts):
    st = np.zeros(len(parts), dtype=synth.STATE_DTYPE)
    for i, (x, y, th, sc, hi) in enumerate(parts):
        st[i]["init_x_px"], st[i]["init_y_px"], st[i]["theta"], st[i]["scale"], st[i]["have_init"] = x, y, th, sc, hi
    return st


@pytest.mark.parametrize("case", ["uniform_free_scale", "gaussian_fixed_scale", "metric_position"])
def test_init_particles_equals_libstdcxx_free_twin(small_world, case):
    """the C++ restatement (libstdc++'s own engine and distributions, the reference's calls) against a twin that has no
    libstdc++ in it — numpy's Mersenne twister, generate_canonical and the polar method written out: same states to the
    bit, same number of engine outputs consumed (the rejection loop runs the same number of times)"""
    layers = small_world["layers"]
    cols, rows = layers.shape[1], layers.shape[2]
    road = np.argwhere((layers[1] < 1) & (small_world["mask"] == 0))
    cx, cy = (float(v) + 0.5 for v in road[len(road) // 2])
    mc = (cols // 2, rows // 2)
    kw = {"uniform_free_scale": dict(fixed_scale=-1.0),
          "gaussian_fixed_scale": dict(fixed_scale=2.0, init_pos_px=(cx, cy), init_pos_px_cov=7.5, init_pos_deg_theta=-140.0, init_pos_deg_cov=12.0),
          "metric_position": dict(fixed_scale=2.0, init_pos_m=((cx - mc[0]) / 2.0, (cy - mc[1]) / 2.0), init_pos_px_cov=4.0)}[case]
    n = 60 if case == "uniform_free_scale" else 45
    got, frozen, px, used = orc.init_particles(99, layers, 1.0, mc, n, **kw)
    parts, used_t = twin.init_particles(99, layers, 1.0, mc, n, **kw)
    assert len(got) == n and frozen == (case != "uniform_free_scale")
    assert np.array_equal(got, _states_of(parts)) and used == used_t
    assert used >= 3 * n * (2 if case != "uniform_free_scale" else 3) / (10 if case == "uniform_free_scale" else 1)
    # every particle sits on a pixel the reference accepts: class 1 present in the DISTANCE layer (road or unknown)
    assert all(1 in orc.classes_at_point(layers, 1.0, int(s["init_x_px"]), int(s["init_y_px"])) for s in got)
    if case == "metric_position":
        assert abs(px[0] - cx) < 1e-3 and abs(px[1] - cy) < 1e-3
    if case == "uniform_free_scale":
        assert np.allclose(got["scale"][:10], 10.0 ** (np.arange(10) / 10), rtol=1e-6) and (got["have_init"] == 0).all()
    # the engine resumes where initialisation left it: update's uniform is output `used` of the same stream
    eng = twin.LibstdcxxEngine(99)
    for _ in range(used):
        eng.canonical()
    assert orc.uniform_draw(99, discard=used) == float(eng.canonical())


def test_init_particles_metric_position_off_map_or_off_road_leaves_the_filter_empty(small_world):
    layers = small_world["layers"]
    cols, rows = layers.shape[1], layers.shape[2]
    st, frozen, px, used = orc.init_particles(1, layers, 1.0, (cols // 2, rows // 2), 20, fixed_scale=2.0, init_pos_m=(1e5, 0.0))
    assert len(st) == 0 and frozen and used == 0 and px[0] > cols
    bare = np.ones_like(layers)                       # no class anywhere within 4 px of the requested position
    st, _, _, used = orc.init_particles(1, bare, 1.0, (cols // 2, rows // 2), 20, fixed_scale=2.0, init_pos_m=(0.0, 0.0))
    assert len(st) == 0 and used == 0
    assert twin.init_particles(1, bare, 1.0, (cols // 2, rows // 2), 20, fixed_scale=2.0, init_pos_m=(0.0, 0.0)) == ([], 0)


def test_freeze_scale_known_answer():
    st = np.zeros(4, dtype=synth.STATE_DTYPE)
    st["scale"] = [1.0, 2.0, 4.0, 8.0]
    out, g = orc.freeze_scale(st)
    assert abs(g - 8.0 ** 0.5) < 1e-6 and (out["scale"] == np.float32(g)).all()
    f = np.float32(1)
    for s in st["scale"]:
        f = np.float32(float(f) * math.pow(float(s), 1.0 / 4))      # float accumulator, double pow
    assert np.float32(g) == f


def test_init_mini_fixture_reproduces():
    """tests/golden/init_mini.npz (make_golden.py): three initialisations on the shared engine.  Uniform positions involve
    no transcendental (bit-exact); Gaussian ones go through glibc's logf (2 ulp between machines); the number of engine
    outputs consumed — how often the rejection loop ran — and the uniform that follows are exact."""
    g = np.load(os.path.join(GOLD, "init_mini.npz"))
    Cn = int(g["num_classes"])
    layers, _ = orc.compute_dists(orc.class_image_to_layers(g["img"], g["lut"], Cn, 1.0), 1.0)
    H, W = g["img"].shape
    from tests.golden.make_golden import INIT_CASES
    for name, kw in INIT_CASES.items():
        kw = dict(kw)
        if name == "metric":
            kw["init_pos_m"] = tuple(float(v) for v in g["metric_m"])
        if name == "pixel":
            kw["init_pos_px"] = tuple(float(v) for v in g["pixel_px_in"])
        st, frozen, px, used = orc.init_particles(2024, layers, 1.0, (W // 2, H // 2), 70, **kw)
        want = g[f"{name}_states"]
        assert len(st) == len(want) == 70 and used == int(g[f"{name}_used"]), name
        for k in ("init_x_px", "init_y_px", "theta", "scale"):
            assert _ulps(st[k], want[k]).max() <= (0 if name == "free" and k != "scale" else 2), (name, k)
        assert np.array_equal(st["have_init"], want["have_init"]) and not st["dx_m"].any() and not st["dy_m"].any()
        assert np.allclose(px, g[f"{name}_px"]) and orc.uniform_draw(2024, discard=used) == float(g[f"{name}_u"])


# ---- the committed fixture computed by the REFERENCE'S OWN SOURCE (tests/golden/make_ref_golden.py, oracle/_ref) --------
def test_ref_mini_fixture_from_the_reference_build():
    """every array of tests/golden/ref_mini.npz was produced by the reference's classes compiled from /root/reference/src
    (oracle/ref_shim/README.md).  The oracle reproduces it: bit for bit everywhere except the weights, which pass through
    Eigen reductions that build sums sequentially (1e-6 relative).  Unlike tests/test_ref_build.py this needs neither the
    reference nor its build, so it also runs wherever only the repository travels."""
    g = np.load(os.path.join(GOLD, "ref_mini.npz"))
    Cn, res, ang = int(g["num_classes"]), float(g["res"]), g["ang_res"]
    img, lut, pts = g["img"], g["lut"], g["pts"]
    H, W = img.shape
    bits = lambda a, b: a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))   # noqa: E731
    layers, mask = orc.compute_dists(orc.class_image_to_layers(img, lut, Cn, 1.0), 1.0)
    assert bits(layers, g["layers"]) and np.array_equal(mask, g["mask"])                      # a3, a4
    tab = orc.polar_table(100, 25, ang, 1.0)
    assert bits(tab.reshape(-1), g["tab"].reshape(-1))                                         # a6 (libm cos / sin on both sides)
    scan = orc.render_polar(pts, res, ang, 100, 25, lut, Cn)
    assert np.array_equal(scan, g["scan"])                                                     # a1
    assert np.array_equal(orc.render_cart(pts, res, 40, 56, lut, Cn).reshape(g["cart"].shape), g["cart"])   # a2
    for c, d_ref, m_ref in zip(g["centres"], g["local_d"], g["local_m"]):                      # a7
        d, m = orc.local_map_polar(layers, mask, 1.0, tab, c[0], c[1], 2.0, res)
        assert bits(d.reshape(d_ref.shape), d_ref) and np.array_equal(m.reshape(m_ref.shape), m_ref)
    # initializeParticles -> propagate -> update on one engine
    sys.path.insert(0, GOLD)
    from make_ref_golden import FILTER, KW, MOTION, N, SEED
    kw = dict(KW, init_pos_px=tuple(float(v) for v in g["init_px"]), init_pos_deg_theta=float(g["init_theta_deg"]))
    st0, frozen, _, used = orc.init_particles(SEED, layers, 1.0, (W // 2, H // 2), N, **kw)
    assert np.array_equal(st0, g["init_states"]) and frozen and orc.engine_peek(SEED, used) == int(g["engine_peek"][0])
    st1, ld1, _, used_p = orc.propagate(st0, *MOTION, True, FILTER["pos_cov"], FILTER["theta_cov"], SEED, discard=used)
    assert np.array_equal(st1, g["propagated"]) and bits(ld1, g["last_dist"]) and orc.engine_peek(SEED, used + used_p) == int(g["engine_peek"][1])
    assert np.array_equal(orc.gmm_samples(st1, N), g["gmm_samples"])
    thetas, shifts = orc.search_list(100)
    fp = orc.make_params(Cn, regularization=FILTER["regularization"], map_width=W, map_height=H)
    st_o = st1.copy()
    raw = orc.score_all(st_o, fp, layers, mask, 1.0, tab, 100, 25, scan, res, thetas, shifts)
    assert np.max(np.abs(raw - g["raw"]) / g["raw"]) <= 1e-6 and np.array_equal(st_o, g["scored"])          # a9, a10
    wn, arg, _ = orc.normalize(g["raw"].copy(), ld1)                                           # a11 on the reference's raw weights
    assert np.max(np.abs(wn - g["weights_norm"]) / g["weights_norm"]) <= 1e-6
    cov4 = np.zeros((1, 4, 4), np.float32)
    cov4[0, :3, :3] = g["gmm_cov"]
    M = orc.adaptive_count(cov4, N, N)
    assert M == len(g["resampled"])
    u = orc.uniform_draw(SEED, discard=used + used_p)
    idx = orc.resample_fast(g["weights_norm"], u, M)                                           # a12 on the reference's weights
    assert np.array_equal(g["scored"][idx], g["resampled"]) and np.array_equal(idx, orc.resample_literal(g["weights_norm"], u, M))
    assert orc.engine_peek(SEED, used + used_p + 1) == int(g["engine_peek"][2])
    mean, cov = orc.mean_cov(g["resampled"])                                                   # a13
    assert bits(mean, g["mean"]) and bits(cov.reshape(-1), g["cov"].reshape(-1))
    ml, _ = orc.ml_cov(g["scored"], int(np.argmax(g["weights_norm"])))
    assert bits(ml, g["ml"])
    # the search set: gates (weight 0), a NaN heading-known particle, the all-NaN search (1 / (FLT_MAX + reg), a denormal)
    fp2 = orc.make_params(Cn, regularization=0.7, force_on_map=True, map_width=W, map_height=H)
    s_o = g["search_in"].copy()
    s_raw = orc.score_all(s_o, fp2, layers, mask, 1.0, tab, 100, 25, scan, res, thetas, shifts)
    r = g["search_raw"]
    assert np.array_equal(np.isnan(s_raw), np.isnan(r)) and np.isnan(r).sum() == 4 and np.array_equal(s_raw == 0, r == 0) and (r == 0).sum() == 18
    tiny = r < 1e-30
    assert (tiny & (r > 0)).sum() >= 3 and np.array_equal(s_raw[tiny & (r > 0)], r[tiny & (r > 0)])
    ok = ~np.isnan(r) & ~tiny
    assert np.max(np.abs(s_raw[ok] - r[ok]) / r[ok]) <= 1e-6
    assert np.array_equal(s_o["have_init"], g["search_scored"]["have_init"]) and (s_o["theta"] == g["search_scored"]["theta"]).mean() >= 0.98
    s_wn, _, _ = orc.normalize(r.copy(), g["search_last_dist"])
    assert np.allclose(s_wn, g["search_weights_norm"], rtol=1e-6, atol=1e-12) and not np.isnan(g["search_weights_norm"]).any()


def test_heading_flip_helper_accepts_ties_and_rejects_real_flips():
    """tests.common.assert_heading_flips_are_ties (what the GPU tests use to judge a heading that differs from the
    oracle's): identical headings pass with zero flips, a heading moved five candidates away is rejected."""
    from tests.common import N_R, N_THETA, assert_heading_flips_are_ties, make_world
    from top_down_renderer_b200 import synth
    wd = make_world(h=300, w=300)
    st, _ = synth.particles_global(64, wd.class_map)
    st_o = st.copy()
    orc.score_all(st_o, wd.fp, wd.layers, wd.mask, 1.0, wd.tab, N_THETA, N_R, wd.scan, 4.0, wd.thetas, wd.shifts)
    assert assert_heading_flips_are_ties(wd, st, st_o["theta"], st_o["theta"], 4.0) == 0
    th = np.asarray(wd.thetas, dtype=np.float32)
    bad = st_o["theta"].copy()
    k = int(np.flatnonzero(th == bad[3])[0])
    bad[3] = th[(k + 5) % len(th)]
    with pytest.raises(AssertionError):
        assert_heading_flips_are_ties(wd, st, bad, st_o["theta"], 4.0)
