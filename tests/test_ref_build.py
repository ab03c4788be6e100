"""The oracle against the REFERENCE'S OWN SOURCE.  Six translation units of the reference (scan_renderer.cpp,
scan_renderer_polar.cpp, top_down_map_polar.cpp, state_particle.cpp, particle_filter.cpp, active_localizer.cpp) are
compiled unmodified from /root/reference/src against stand-in headers for the libraries this image lacks
(oracle/ref_shim/, see its README.md) into oracle/_ref/libtdr_ref.so; these tests run the reference's classes beside the
oracle on the same inputs and the same seeded engine.

What comes out bit for bit: class images, polar gathers, particle initialisation, propagate, resampled states, pose —
everything whose arithmetic is the reference's own statements plus libm / libstdc++.  What is held to a tolerance:
values that pass through an Eigen reduction (.sum()), because the stand-in reduces sequentially while Eigen (and the
oracle, which restates Eigen's SSE2 order) associates differently — 1e-6 relative, ten times tighter than the contract.
CPU only; skipped where neither /root/reference nor a prebuilt oracle/_ref exists."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refbuild as ref
from top_down_renderer_b200 import synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="no /root/reference and no prebuilt oracle/_ref")
ANG = np.float32(2 * math.pi / 100)
C_ = 4


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.fixture(scope="module")
def world():
    cm = synth.make_class_map(260, 300, C_, seed=21)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C_)
    seeds = orc.class_image_to_layers(img, lut, C_, 1.0)
    layers, mask = orc.compute_dists(seeds, 1.0)
    geo, _ = orc.compute_dists(orc.geo_raster(seeds), 1.0)
    pose, heading = synth.default_pose(cm, seed=21)
    pts = synth.make_scan(cm, pose, heading, seed=21, n_rings=32, n_az=256)
    tab = orc.polar_table(100, 25, ANG, 1.0)
    H, W = cm.shape
    m = ref.Map(layers, mask, 1.0, tab, 100, 25, geo=geo, center=(W // 2, H // 2))
    thetas, shifts = orc.search_list(100)
    return dict(cm=cm, img=img, seeds=seeds, lut=lut, layers=layers, mask=mask, geo=geo, pose=pose, heading=heading, pts=pts, tab=tab, H=H, W=W, map=m,
                thetas=thetas, shifts=shifts, scan=orc.render_polar(pts, 1.5, ANG, 100, 25, lut, C_))


# ---- a1 / a2: ScanRendererPolar / ScanRenderer::renderSemanticTopDown ----------------------------------------------
@pytest.mark.parametrize("res", [0.5, 1.5, 4.0])
def test_class_images_equal_the_reference_renderers(world, res):
    pts = world["pts"].copy()
    pts[::17, :2] = 0                                           # dropped returns: x == 0 && y == 0 is skipped
    pts[5::23, 4] = 7                                           # a class the lut maps to -1
    lut = np.full(256, -1, dtype=np.int32)                      # every intensity the scan holds must index the lut: the
    lut[:C_] = np.arange(C_)                                    # reference reads flatten_lut_[pt_class] unchecked (:103-104)
    got = ref.render_polar(pts, res, ANG, 100, 25, lut, C_)
    assert np.array_equal(got, orc.render_polar(pts, res, ANG, 100, 25, lut, C_)) and got.sum() > 1000
    cart = ref.render_cart(pts, res, 48, 64, lut, C_)
    assert np.array_equal(cart, orc.render_cart(pts, res, 48, 64, lut, C_).reshape(cart.shape)) and cart.sum() > 100


# ---- a7: TopDownMapPolar::getLocalMap / getLocalGeoMap ----------------------------------------------------------------
@pytest.mark.parametrize("res", [4.0, 1.0, 0.37])
def test_geometric_renderers_equal_the_reference(world, res):
    """f2: ScanRendererPolar / ScanRenderer::renderGeometricTopDown (scan_renderer_polar.cpp:6-81, scan_renderer.cpp:7-53)
    compiled from the reference's own sources against the oracle's restatement: an organised 1024 x 64 cloud with rough
    ground (so that both slope branches, the fill loop and the line drawing are exercised), bit-identical images"""
    pts = synth.make_scan(world["cm"], world["pose"], world["heading"], seed=21)      # 64 rings x 1024 azimuths
    rng = np.random.default_rng(12)
    pts[:, 2] = -2.0 + rng.normal(0, 0.4, len(pts)).astype(np.float32) * (rng.random(len(pts)) < 0.3)
    a, b = orc.render_geometric_polar(pts, 1024, 64, res, ANG, 100, 25), ref.render_geometric_polar(pts, 1024, 64, res, ANG, 100, 25)
    assert np.array_equal(a, b) and a[0].sum() > 100 and a[1].sum() > 100
    a, b = orc.render_geometric_cart(pts, 1024, 64, res, 150, 170), ref.render_geometric_cart(pts, 1024, 64, res, 150, 170)
    assert np.array_equal(a, b) and a[0].sum() > 100 and a[1].sum() > 100
    # a transposed organisation (64 columns of 1024 points) visits the points in another order
    a, b = orc.render_geometric_polar(pts, 64, 1024, res, ANG, 100, 25), ref.render_geometric_polar(pts, 64, 1024, res, ANG, 100, 25)
    assert np.array_equal(a, b)


def test_polar_gather_equals_the_reference(world):
    w = world
    rng = np.random.default_rng(1)
    centres = [(150.3, 120.8), (2.0, 3.0), (299.6, 259.4), (-40.0, 100.0), (-5000.0, -5000.0), (149.5, 130.5)]
    centres += [tuple(rng.uniform(-20, 320, 2)) for _ in range(10)]
    for k, (cx, cy) in enumerate(centres):
        scale, res = (2.0, 1.5) if k % 2 else (1.3, 4.0)
        d, m = w["map"].local_map_polar(cx, cy, scale, res)
        do, mo = orc.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], cx, cy, scale, res)
        assert same_bits(d, do.reshape(d.shape)) and np.array_equal(m, mo.reshape(m.shape)), (cx, cy)
        g = w["map"].local_geo_polar(cx, cy, scale, res)
        go, _ = orc.local_map_polar(w["geo"], np.zeros_like(w["mask"]), 1.0, w["tab"], cx, cy, scale, res)
        assert same_bits(g, go.reshape(g.shape))


def _tracking_kwargs(w):
    return dict(fixed_scale=2.0, init_pos_px=(float(w["pose"][0]), float(w["pose"][1])), init_pos_px_cov=6.0,
                init_pos_deg_theta=math.degrees(w["heading"]), init_pos_deg_cov=3.0)


# ---- the whole filter step on ONE shared engine: initializeParticles -> propagate -> update -> pose ------------------
def test_filter_step_equals_the_reference(world):
    w = world
    W, H, seed, N = w["W"], w["H"], 77, 500
    kw = _tracking_kwargs(w)
    f = ref.Filter(w["map"], N, seed, regularization=0.7, pos_cov=0.15, theta_cov=0.004, **kw)
    # initializeParticles (particle_filter.cpp:19-84, state_particle.cpp:3-49): states and engine position
    st0, _, _ = f.get()
    so, frozen, px, used = orc.init_particles(seed, w["layers"], 1.0, (W // 2, H // 2), N, **kw)
    assert len(st0) == N and np.array_equal(st0, so) and f.scale_frozen() == frozen and f.scale() == 2.0
    assert f.engine_peek() == orc.engine_peek(seed, used)
    # propagate (:86-92, state_particle.cpp:57-78)
    f.propagate(0.4, 0.05, 0.01)
    st1, ld1, _ = f.get()
    want, last_o, _, used_p = orc.propagate(so, 0.4, 0.05, 0.01, True, 0.15, 0.004, seed, discard=used)
    assert np.array_equal(st1, want) and same_bits(ld1, last_o) and (ld1 > 0).all()
    assert f.engine_peek() == orc.engine_peek(seed, used + used_p)
    # the mixture the count of the next resampling comes from (:151-158); fitted on this thread from the propagated set
    samples, _, covs = f.gmm()
    assert np.array_equal(samples, orc.gmm_samples(st1, min(1000, N)))         # the EM input matrix (:262-272)
    cov4 = np.zeros((1, 4, 4), np.float32)
    cov4[0, :3, :3] = covs[0]
    M = orc.adaptive_count(cov4, N, N)
    # update (:94-189): raw weights, normalised weights, resampled states
    f.update(w["scan"], 1.5)
    scored, ld_s, raw = f.get(scored_set=True)
    assert same_bits(ld_s, ld1)
    fp = orc.make_params(C_, regularization=0.7, map_width=W, map_height=H)
    st_o = st1.copy()
    raw_o = orc.score_all(st_o, fp, w["layers"], w["mask"], 1.0, w["tab"], 100, 25, w["scan"], 1.5, w["thetas"], w["shifts"])
    assert not np.isnan(raw).any() and np.max(np.abs(raw - raw_o) / raw_o) <= 1e-6
    assert np.array_equal(scored, st_o)                                         # theta / have_init as the reference left them
    wn = f.weights()
    wn_o, arg_o, _ = orc.normalize(raw.copy(), ld_s)                            # stage-wise: the reference's own raw weights
    assert np.max(np.abs(wn - wn_o) / wn_o) <= 1e-6 and int(np.argmax(wn)) == arg_o
    cur, _, _ = f.get()
    assert f.num_particles() == len(cur) == M and M < N
    u = orc.uniform_draw(seed, discard=used + used_p)                           # the ONE draw of :172-173, next on the engine
    assert np.array_equal(cur, scored[orc.resample_fast(wn, u, M)])             # indices: bit-exact on the reference's weights
    assert np.array_equal(orc.resample_fast(wn, u, M), orc.resample_literal(wn, u, M))
    assert f.engine_peek() == orc.engine_peek(seed, used + used_p + 1)
    # pose (:191-236)
    mean, cov, ml, cov_ml = f.pose()
    mo, co = orc.mean_cov(cur)
    assert same_bits(mean, mo) and same_bits(cov.reshape(-1), co.reshape(-1))
    mlo, _ = orc.ml_cov(scored, int(np.argmax(wn)))
    assert same_bits(ml, mlo)
    both = np.concatenate([scored[int(np.argmax(wn)):][:1], cur])               # computeCov: the current set about the ML pose
    _, cmo = orc.ml_cov(both, 0)
    assert np.allclose(cov_ml.reshape(-1) * (len(cur) - 1), cmo.reshape(-1) * len(cur), rtol=1e-5, atol=1e-5)


# ---- global localisation: free scale, no heading -> the 40-candidate theta search, the scale gate, NaN weights ----------
def test_theta_search_and_gates_equal_the_reference(world):
    w = world
    W, H, seed, N = w["W"], w["H"], 5, 400
    f = ref.Filter(w["map"], N, seed, regularization=0.7, fixed_scale=-1.0, scale_log_min=-0.1, scale_log_max=0.65, force_on_map=True)
    st0, _, _ = f.get()
    so, frozen, _, used = orc.init_particles(seed, w["layers"], 1.0, (W // 2, H // 2), N, fixed_scale=-1.0)
    assert np.array_equal(st0, so) and not frozen and f.scale() == -1.0 and (st0["have_init"] == 0).all()
    # push a few particles off the map (force_on_map gate) and onto unknown ground before the update
    st0["dx_m"][:7] = 1e4
    f.set(st0)
    f.propagate(0.2, 0.0, 0.0)                                                  # scale jitter on: four draws per particle
    st1, ld1, _ = f.get()
    want, last_o, _, used_p = orc.propagate(st0, 0.2, 0.0, 0.0, False, 0.3, 0.0314, seed, discard=used)
    assert np.array_equal(st1, want) and same_bits(ld1, last_o)
    f.update(w["scan"], 1.5)
    scored, ld_s, raw = f.get(scored_set=True)
    fp = orc.make_params(C_, regularization=0.7, fixed_scale=-1.0, force_on_map=True, map_width=W, map_height=H)
    fp.scale_log_min, fp.scale_log_max = -0.1, 0.65
    st_o = st1.copy()
    raw_o = orc.score_all(st_o, fp, w["layers"], w["mask"], 1.0, w["tab"], 100, 25, w["scan"], 1.5, w["thetas"], w["shifts"])
    assert np.array_equal(np.isnan(raw), np.isnan(raw_o)) and np.array_equal(raw == 0, raw_o == 0)
    assert (raw[:7] == 0).all() and (raw == 0).sum() > 7                        # off the map; scale outside 10^[-0.1, 0.65]
    ok = ~np.isnan(raw) & (raw != 0)
    assert ok.sum() > 100 and np.max(np.abs(raw[ok] - raw_o[ok]) / raw_o[ok]) <= 1e-6
    # the heading each particle chose: the first strict minimum over 40 candidates; a different summation order can only
    # flip near-ties
    searched = ok | np.isnan(raw)
    assert (scored["have_init"][searched] == 1).all() and (scored["have_init"][raw == 0] == 0).all()
    assert (scored["theta"][searched] == st_o["theta"][searched]).mean() >= 0.99
    wn = f.weights()
    wn_o, _, _ = orc.normalize(raw.copy(), ld_s)
    assert np.allclose(wn, wn_o, rtol=1e-6, atol=0) and not np.isnan(wn).any()
    # freezeScale (:343-357) on the resampled set
    cur, _, _ = f.get()
    f.freeze_scale()
    froz, _, _ = f.get()
    fo, g = orc.freeze_scale(cur)
    assert np.array_equal(froz, fo) and f.scale_frozen() and f.scale() == np.float32(g)


def test_nan_weights_take_the_mean_minus_lower_deviation(world):
    """particles whose polar footprint is mostly unknown get NaN (state_particle.cpp:117-120) and the normalisation
    replaces it by mean - bottom_stddev (particle_filter.cpp:107-134): the reference's loops against the oracle's"""
    w = world
    W, H, seed, N = w["W"], w["H"], 9, 300
    kw = _tracking_kwargs(w)
    f = ref.Filter(w["map"], N, seed, regularization=0.7, **kw)
    st, _, _ = f.get()
    st["init_x_px"][::5] = -150.0                                               # footprints hanging off the map edge
    st["init_y_px"][1::7] = H + 120.0
    ld = np.random.default_rng(3).uniform(0, 0.4, N).astype(np.float32)
    f.set(st, ld)
    f.update(w["scan"], 1.5)
    scored, ld_s, raw = f.get(scored_set=True)
    assert same_bits(ld_s, ld) and 20 < np.isnan(raw).sum() < N - 50
    fp = orc.make_params(C_, regularization=0.7, map_width=W, map_height=H)
    raw_o = orc.score_all(st.copy(), fp, w["layers"], w["mask"], 1.0, w["tab"], 100, 25, w["scan"], 1.5, w["thetas"], w["shifts"])
    assert np.array_equal(np.isnan(raw), np.isnan(raw_o))
    wn = f.weights()
    wn_o, arg_o, stats = orc.normalize(raw.copy(), ld_s)
    assert stats[5] == 0 and not np.isnan(wn).any() and np.max(np.abs(wn - wn_o) / wn_o) <= 1e-6


# ---- ParticleFilter::updateMap, the metric initial position ------------------------------------------------------------
def test_map_centre_shift_and_metric_initial_position(world):
    w = world
    W, H = w["W"], w["H"]
    kw = _tracking_kwargs(w)
    own = ref.Map.from_class_image(w["img"], w["lut"], C_, 1.0, center=(W // 2, H // 2))     # updateMap replaces the layers
    own.set_polar_table(w["tab"], 100, 25)
    f = ref.Filter(own, 64, 3, **kw)
    a, _, _ = f.get()
    assert np.array_equal(a, orc.init_particles(3, w["layers"], 1.0, (W // 2, H // 2), 64, **kw)[0])
    f.update_map(w["img"], (W // 2 + 3, H // 2 - 2))                            # last_map_center_ starts at (0, 0) (:12)
    b, _, _ = f.get()
    fl = np.float32
    assert same_bits(b["init_x_px"], a["init_x_px"] + fl(W // 2 + 3)) and same_bits(b["init_y_px"], a["init_y_px"] + fl(H // 2 - 2))
    assert own.info()[4] == (W // 2 + 3, H // 2 - 2)
    f.update_map(w["img"], (W // 2, H // 2))
    c, _, _ = f.get()
    assert same_bits(c["init_x_px"], b["init_x_px"] + fl(-3)) and same_bits(c["init_y_px"], b["init_y_px"] + fl(2))
    # metric position relative to the map centre (:27-54)
    m_xy = ((float(w["pose"][0]) - W // 2) / 2.0, (float(w["pose"][1]) - H // 2) / 2.0)
    kw2 = dict(kw, init_pos_px=(-1.0, -1.0), init_pos_m=m_xy)
    g = ref.Filter(w["map"], 48, 11, **kw2)
    got, _, _ = g.get()
    so, _, px, used = orc.init_particles(11, w["layers"], 1.0, (W // 2, H // 2), 48, **kw2)
    assert len(got) == 48 and np.array_equal(got, so) and g.init_px() == px and g.engine_peek() == orc.engine_peek(11, used)
    off = ref.Filter(w["map"], 48, 12, **dict(kw2, init_pos_m=(1e5, 0.0)))
    assert off.count() == 0 and off.num_particles() == 0
    assert len(orc.init_particles(12, w["layers"], 1.0, (W // 2, H // 2), 48, **dict(kw2, init_pos_m=(1e5, 0.0)))[0]) == 0
    off.update(w["scan"], 1.5)                                                  # no particles: returns at once (:96-99)
    assert off.count() == 0


# ---- ActiveLocalizer::getBestRelPos ------------------------------------------------------------------------------------
def test_active_localizer_equals_the_reference(world):
    w = world
    rng = np.random.default_rng(3)
    for n in (1, 2, 4):
        preds = np.stack([rng.uniform(0.2, 0.8, n) * w["W"], rng.uniform(0.2, 0.8, n) * w["H"], rng.uniform(-3, 3, n)], axis=1).astype(np.float32)
        rel = w["map"].active_best_rel_pos(preds)
        rel_o, _ = orc.active_best_rel_pos(w["layers"], w["mask"], 1.0, w["tab"], 100, 25, preds)
        assert rel == rel_o, (n, rel, rel_o)


# ---- src/top_down_map.cpp itself: a3 - a6, a8, the vector map, both caches -----------------------------------------------
@pytest.mark.parametrize("resolution", [1.0, 0.5, 2.0])
def test_distance_fields_equal_the_reference_map_code(world, resolution):
    """TopDownMap::updateMap = loadCompressedRasterMap + computeDists from the reference's source; cv::distanceTransform
    itself is the stand-in's exact transform (OpenCV's values, pinned by cv2 elsewhere) — what is checked here is
    everything AROUND it: the flipped / scaled sampling of the image, the lut, the unknown mask built from uint8 sums,
    the 8-bit conversion, x resolution, truncation at 50, zeroing under the mask"""
    w = world
    m = ref.Map.from_class_image(w["img"], w["lut"], C_, resolution, center=(7, -3))
    layers, mask = m.get()
    seeds = orc.class_image_to_layers(w["img"], w["lut"], C_, resolution)
    lo, mo = orc.compute_dists(seeds, resolution)
    assert layers.shape == lo.shape and same_bits(layers, lo) and np.array_equal(mask, mo)
    rows, cols, k, have, centre = m.info()
    assert (rows, cols) == orc.map_dims(w["H"], w["W"], resolution) and k == C_ and have and centre == (7, -3)
    # getGeoRasterMap + computeDists (:410-427, :47-58)
    geo = m.build_geo_from_binary(seeds)
    go, _ = orc.compute_dists(orc.geo_raster(seeds), resolution)
    assert same_bits(geo, go)
    # getClassesAtPoint, both overloads (:159-175; the float one divides by the resolution twice)
    rng = np.random.default_rng(2)
    for _ in range(200):
        x, y = (float(v) for v in rng.uniform(-10, max(w["W"], w["H"]) + 10, 2))
        assert m.classes_at(x, y) == orc.classes_at_point(lo, resolution, int(x), int(y))
        assert m.classes_at(x, y, as_float=True) == orc.classes_at_point(lo, resolution, int(np.float32(x) / np.float32(resolution)),
                                                                         int(np.float32(y) / np.float32(resolution)))


def test_map_without_a_non_road_cell_is_not_a_map(world):
    """the isZero(0) test of updateMap (:150) runs on the BINARY layer 1: all road -> have_map_ stays false"""
    img = np.full((40, 50), 1, dtype=np.uint8)
    assert not ref.Map.from_class_image(img, world["lut"], C_, 1.0).info()[3]
    img[3, 4] = 2
    assert ref.Map.from_class_image(img, world["lut"], C_, 1.0).info()[3]


def test_polar_table_and_cartesian_gather_through_sample_pts(world):
    """samplePts (:367-389) assigns a replicated LinSpaced row to strided maps of another shape — defined only with
    Eigen's assertions off; the stand-in implements the reading the oracle restates (destination (i, j) <- row[j mod n]),
    so this checks the code AROUND that reading: strides, rotation, centre offsets, the polar post-processing, rounding"""
    w = world
    for (nt, nr) in ((100, 25), (100, 50), (36, 7)):
        ang = np.float32(2 * math.pi / nt)
        for resolution in (1.0, 0.5):
            m = ref.Map.from_class_image(w["img"], w["lut"], C_, resolution)
            assert same_bits(m.polar_table(nt, nr, ang), orc.polar_table(nt, nr, ang, resolution).reshape(-1, 2))
    m = ref.Map.from_class_image(w["img"], w["lut"], C_, 1.0)
    for (cx, cy, rot, res, rows, cols) in [(150.0, 130.0, 0.0, 1.0, 50, 50), (20.3, 250.1, 0.7, 2.0, 40, 31), (217.8, 99.2, -2.1, 0.5, 33, 64)]:
        d, k = m.local_map_cart(cx, cy, rot, res, rows, cols)
        do, ko = orc.local_map_cart(w["layers"], w["mask"], 1.0, cx, cy, rot, res, rows, cols)
        assert same_bits(d, do.reshape(d.shape)) and np.array_equal(k, ko.reshape(k.shape)), (cx, cy, rot)


SVG_W, SVG_H = 230, 170
SVG_SHAPES = [("#101010", [(-5, -5), (240, -5), (240, 180), (-5, 180)]),                                        # class 0: background
              ("#2040c0", [(20.25, 30.5), (200.75, 35.25), (205.0, 60.75), (120.5, 55.5), (110.25, 140.75), (80.75, 139.25), (85.5, 52.5), (18.75, 58.25)]),
              ("#c04020", [(150.25, 90.25), (190.75, 90.25), (190.75, 130.75), (150.25, 130.75)]),
              ("#20c040", [(60.25, 100.5), (100.5, 70.25), (140.75, 110.5), (95.25, 160.75)]),
              ("#20c040", [(200.5, 120.25), (260.0, 125.5), (250.25, 200.0), (190.5, 190.25)])]
SVG_CLASS_HEX = ["#101010", "#2040c0", "#c04020", "#20c040"]


def _write_svg(path):
    body = "".join(f'<polygon fill="{col}" points="{" ".join(f"{x},{y}" for x, y in pts)}"/>\n' for col, pts in SVG_SHAPES)
    with open(path, "w") as f:
        f.write(f'<svg xmlns="http://www.w3.org/2000/svg" width="{SVG_W}" height="{SVG_H}">\n{body}</svg>\n')


def _packed(hexcol):                                          # what loadSvg compares with nanosvg's 0xBBGGRR (:80-86)
    r, g, b = int(hexcol[1:3], 16), int(hexcol[3:5], 16), int(hexcol[5:7], 16)
    return b << 16 | g << 8 | r


def test_static_map_constructor_svg_raster_cache_and_eig_cache(tmp_path):
    """the reference's static-map constructor (:9-64) end to end on an svg file: nanosvg (vendored in the reference) ->
    loadSvg -> getRasterMap / samplePts / getClasses -> saveRasterizedMaps -> geo maps -> computeDists -> saveCachedMaps;
    then again from the .eig cache, and once more from the raster cache directory — against the oracle's polygon
    rasteriser and this repository's readers of both cache formats"""
    from top_down_renderer_b200 import eigcache, rastercache
    home = tmp_path / "home"
    (home / ".ros").mkdir(parents=True)
    svg = str(tmp_path / "campus.svg")
    _write_svg(svg)
    lut = np.arange(C_, dtype=np.int32)
    colors = [_packed(c) for c in SVG_CLASS_HEX]
    excl = [0, 1]
    m = ref.Map.from_path(str(home), svg, lut, C_, 1.0, colors, exclusive=excl)
    layers, mask, geo = m.get(want_geo=True)
    # the oracle on the same polygons (y flipped as loadSvg does, :89), in class order
    cls_of = {c: i for i, c in enumerate(SVG_CLASS_HEX)}
    polys = [np.float32([(x, SVG_H - y) for x, y in pts]) for _, pts in SVG_SHAPES]
    pcls = [cls_of[col] for col, _ in SVG_SHAPES]
    order = sorted(range(len(polys)), key=lambda i: pcls[i])
    binl = orc.raster_polygons([polys[i] for i in order], [pcls[i] for i in order], SVG_W, SVG_H, 0.0, 1.0, C_, excl)
    lo, mo = orc.compute_dists(binl, 1.0)
    assert layers.shape == (C_, SVG_W, SVG_H) and same_bits(layers, lo) and np.array_equal(mask, mo)
    go, _ = orc.compute_dists(orc.geo_raster(binl), 1.0)
    assert same_bits(geo, go)
    assert all((binl[c] == 0).any() and (binl[c] == 1).any() for c in range(C_))
    # the raster cache the reference wrote (cv::imwrite through the stand-in's PNG codec): this repository's reader
    cache = str(tmp_path / "campus_raster_cache")
    assert np.array_equal(rastercache.load_rasterized_maps(cache, C_), binl)
    # the .eig cache it wrote: metadata + distance fields, geo, mask — this repository's reader
    xc = str(home / ".ros" / "xview_cache")
    assert eigcache.cache_is_valid(xc, svg, C_, 1.0)
    c_layers, c_geo, c_mask = eigcache.load_cache(xc, C_)
    assert same_bits(c_layers, lo) and same_bits(c_geo, go) and np.array_equal(c_mask, mo)
    # second construction: a cache hit (loadCacheMetaData + loadCachedMaps); a cache written by THIS repository loads too
    again = ref.Map.from_path(str(home), svg, lut, C_, 1.0, colors, exclusive=excl)
    l2, m2, g2 = again.get(want_geo=True)
    assert same_bits(l2, lo) and same_bits(g2, go) and np.array_equal(m2, mo) and again.info()[3]
    home2 = tmp_path / "home2"
    (home2 / ".ros" / "xview_cache").mkdir(parents=True)
    eigcache.save_cache(str(home2 / ".ros" / "xview_cache"), "some/other/map.svg", lo, go, mo, 1.0)
    ours = ref.Map.from_path(str(home2), "some/other/map.svg", lut, C_, 1.0, colors, exclusive=excl)
    l3, m3, g3 = ours.get(want_geo=True)
    assert same_bits(l3, lo) and same_bits(g3, go) and np.array_equal(m3, mo)
    # third: the raster cache directory as the map path (loadRasterizedMaps :213-224), written by THIS repository
    home3 = tmp_path / "home3"
    (home3 / ".ros").mkdir(parents=True)
    own_cache = str(tmp_path / "own_raster_cache")
    rastercache.save_rasterized_maps(own_cache, binl)
    ras = ref.Map.from_path(str(home3), own_cache, lut, C_, 1.0, colors, exclusive=excl)
    l4, m4 = ras.get()
    assert same_bits(l4, lo) and np.array_equal(m4, mo)


# ---- ties and specials: where two implementations of the same loop usually part ways -----------------------------------
def test_class_images_at_bin_boundaries_and_special_coordinates(world):
    """points on the rounding boundaries of the angular / radial / Cartesian bins (+- a few ulp), NaN and infinite
    coordinates, the (0, 0) return: the reference's float -> int conversions (x86 cvttss2si: NaN / overflow -> INT_MIN)
    against the oracle's restatement of them"""
    rng = np.random.default_rng(5)
    n = 60000
    pts = np.zeros((n, 8), dtype=np.float32)
    k = rng.integers(-51, 52, n)
    th = (k + 0.5) * float(ANG) + rng.normal(0, 2e-6, n)
    kr = rng.integers(0, 27, n)
    r = (kr + 0.5) * 4.0 + rng.normal(0, 2e-5, n)
    pts[:, 0] = (r * np.sin(th)).astype(np.float32)
    pts[:, 1] = (r * np.cos(th)).astype(np.float32)
    half = n // 2                                                 # second half: Cartesian cell boundaries
    pts[half:, 0] = ((rng.integers(-40, 40, n - half) + 0.5) * 1.5 + rng.normal(0, 1e-6, n - half)).astype(np.float32)
    pts[half:, 1] = ((rng.integers(-30, 30, n - half) + 0.5) * 1.5 + rng.normal(0, 1e-6, n - half)).astype(np.float32)
    pts[:, 4] = rng.integers(0, C_ + 2, n).astype(np.float32)    # classes 4, 5 map to -1
    pts[:10, 0] = np.nan
    pts[10:20, 1] = np.inf
    pts[20:30, 0] = -np.inf
    pts[30:40, 0:2] = 0
    pts[40:50, 0] = 3e38
    lut = np.full(256, -1, dtype=np.int32)
    lut[:C_] = np.arange(C_)
    for res in (4.0, 1.5):
        a, b = ref.render_polar(pts, res, ANG, 100, 25, lut, C_), orc.render_polar(pts, res, ANG, 100, 25, lut, C_)
        assert np.array_equal(a, b) and a.sum() > 10000
        a, b = ref.render_cart(pts, res, 61, 81, lut, C_), orc.render_cart(pts, res, 61, 81, lut, C_)
        assert np.array_equal(a, b.reshape(a.shape)) and a.sum() > 10000


def test_polar_gather_at_pixel_boundaries(world):
    """centres and scales that put many lattice points exactly on x.5 pixel boundaries (round half away from zero), on the
    map border and just outside it"""
    w = world
    for cx, cy, scale, res in [(0.0, 0.0, 1.0, 1.0), (100.5, 80.5, 1.0, 0.5), (299.5, 259.5, 2.0, 0.25), (-0.5, -0.5, 1.0, 1.0),
                               (150.0, 130.0, 0.5, 1.0), (w["W"] - 0.5, 10.0, 1.0, 2.0), (1e6, 1e6, 2.0, 4.0), (float("nan"), 5.0, 2.0, 4.0)]:
        d, m = w["map"].local_map_polar(cx, cy, scale, res)
        do, mo = orc.local_map_polar(w["layers"], w["mask"], 1.0, w["tab"], cx, cy, scale, res)
        assert same_bits(d, do.reshape(d.shape)) and np.array_equal(m, mo.reshape(m.shape)), (cx, cy, scale, res)


@pytest.mark.parametrize("seed,N,motion", [(101, 64, (0.0, 0.0, 0.0)), (202, 333, (1.5, -0.7, 0.2)), (303, 1000, (0.05, 0.0, -0.01))])
def test_filter_steps_over_seeds_sizes_and_motions(world, seed, N, motion):
    """two consecutive propagate + update steps (the second on the resampled, adaptively sized set), several seeds"""
    w = world
    W, H = w["W"], w["H"]
    kw = _tracking_kwargs(w)
    f = ref.Filter(w["map"], N, seed, regularization=0.7, pos_cov=0.3, theta_cov=0.0314, **kw)
    st, frozen, _, used = orc.init_particles(seed, w["layers"], 1.0, (W // 2, H // 2), N, **kw)
    fp = orc.make_params(C_, regularization=0.7, map_width=W, map_height=H)
    for step in range(2):
        f.propagate(*motion)
        st, ld, _, used_p = orc.propagate(st, *motion, True, 0.3, 0.0314, seed, discard=used)
        used += used_p
        got, got_ld, _ = f.get()
        assert np.array_equal(got, st) and same_bits(got_ld, ld), step
        _, _, covs = f.gmm()
        cov4 = np.zeros((1, 4, 4), np.float32)
        cov4[0, :3, :3] = covs[0]
        M = orc.adaptive_count(cov4, len(st), N)
        f.update(w["scan"], 1.5)
        scored, _, raw = f.get(scored_set=True)
        st_o = st.copy()
        raw_o = orc.score_all(st_o, fp, w["layers"], w["mask"], 1.0, w["tab"], 100, 25, w["scan"], 1.5, w["thetas"], w["shifts"])
        assert np.array_equal(np.isnan(raw), np.isnan(raw_o)) and np.allclose(raw, raw_o, rtol=1e-6, atol=0, equal_nan=True)
        wn = f.weights()
        wn_o, _, _ = orc.normalize(raw.copy(), ld)
        assert np.allclose(wn, wn_o, rtol=1e-6, atol=0)
        u = orc.uniform_draw(seed, discard=used)
        used += 1
        cur, _, _ = f.get()
        assert len(cur) == M == f.num_particles()
        st = scored[orc.resample_fast(wn, u, M)]
        assert np.array_equal(cur, st) and f.engine_peek() == orc.engine_peek(seed, used), step
    mean, cov, _, _ = f.pose()
    mo, co = orc.mean_cov(st)
    assert same_bits(mean, mo) and same_bits(cov.reshape(-1), co.reshape(-1))


@pytest.mark.parametrize("num_classes,resolution,res,weights", [(6, 1.0, 4.0, [1.0, 0.5, 2.0, 1.5, 0.25, 3.0]), (4, 0.5, 1.5, [1.0, 1.0, 1.0, 1.0]),
                                                                  (5, 2.0, 0.5, [0.0, 1.0, 4.0, 0.1, 1.0])])
def test_weights_over_class_counts_resolutions_headings(num_classes, resolution, res, weights):
    """computeWeight / getCostForRot with a known heading (state_particle.cpp:112-219): class weights, map resolutions other
    than 1, radial bin sizes, headings far outside [0, 2 pi) (the while-loops that normalise the row shift), free scale
    with its gate — injected states, the reference's weights against the oracle's"""
    rng = np.random.default_rng(num_classes)
    cm = synth.make_class_map(220, 260, num_classes, seed=40 + num_classes)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(num_classes)
    layers, mask = orc.compute_dists(orc.class_image_to_layers(img, lut, num_classes, resolution), resolution)
    pose, heading = synth.default_pose(cm, seed=3)
    pts = synth.make_scan(cm, pose, heading, seed=3, n_rings=16, n_az=256)
    scan = orc.render_polar(pts, res, ANG, 100, 25, lut, num_classes)
    tab = orc.polar_table(100, 25, ANG, resolution)
    m = ref.Map.from_class_image(img, lut, num_classes, resolution)
    assert same_bits(m.get()[0], layers)
    assert same_bits(m.polar_table(100, 25, ANG), tab.reshape(-1, 2))
    N = 300
    f = ref.Filter(m, N, 8, regularization=0.15, fixed_scale=-1.0, scale_log_min=-0.2, scale_log_max=0.5, class_weights=weights)
    assert f.count() == N
    st = np.zeros(N, dtype=synth.STATE_DTYPE)
    st["init_x_px"] = rng.uniform(-20, cm.shape[1] + 20, N)
    st["init_y_px"] = rng.uniform(-20, cm.shape[0] + 20, N)
    st["dx_m"], st["dy_m"] = rng.normal(0, 3, N), rng.normal(0, 3, N)
    st["theta"] = rng.uniform(-25, 25, N)                           # many turns either way
    st["theta"][:8] = [0.0, -0.0, 2 * math.pi, -2 * math.pi, math.pi / 100, -math.pi / 100, 199 * math.pi / 100, 1e3]
    st["scale"] = 10.0 ** rng.uniform(-0.4, 0.7, N)                 # some outside 10^[-0.2, 0.5]: gated to 0
    st["have_init"] = 1
    ld = rng.uniform(0, 0.4, N).astype(np.float32)
    f.set(st, ld)
    f.update(scan, res)
    scored, _, raw = f.get(scored_set=True)
    H, W = cm.shape
    fp = orc.make_params(num_classes, regularization=0.15, class_weights=weights, fixed_scale=-1.0, scale_log_min=-0.2, scale_log_max=0.5,
                         map_width=(W / resolution) * resolution, map_height=(H / resolution) * resolution)
    thetas, shifts = orc.search_list(100)
    st_o = st.copy()
    raw_o = orc.score_all(st_o, fp, layers, mask, resolution, tab, 100, 25, scan, res, thetas, shifts)
    assert np.array_equal(np.isnan(raw), np.isnan(raw_o)) and np.array_equal(raw == 0, raw_o == 0) and (raw == 0).sum() > 10
    assert (~np.isnan(raw) & (raw != 0)).sum() > 40
    assert np.allclose(raw, raw_o, rtol=1e-6, atol=0, equal_nan=True)
    assert np.array_equal(scored, st_o)                              # a known heading is left alone
    for t in st["theta"][:40]:                                       # the row shift itself (:123-128)
        assert 0 <= orc.rot_to_shift(float(t), 100) < 100
    wn, wn_o = f.weights(), orc.normalize(raw.copy(), ld)[0]
    assert np.allclose(wn, wn_o, rtol=1e-6, atol=0)


@pytest.mark.parametrize("seed,resolution", [(1, 1.0), (2, 1.0), (3, 0.5)])
def test_random_svg_polygons_with_vertices_on_sample_points(tmp_path, seed, resolution):
    """getRasterMap / getClasses (top_down_map.cpp:328-408) on random concave polygons whose vertices lie on quarter pixels —
    a quarter of them exactly ON sample points (x.5) and many edges horizontal or vertical: the ties of the even-odd
    rule's `<` comparisons and the divisions by zero of horizontal edges, the reference's expression against the oracle's"""
    rng = np.random.default_rng(seed)
    Wm, Hm = 120, 90
    shapes = []
    for _ in range(14):
        k = int(rng.integers(3, 9))
        cx, cy = rng.uniform(-10, Wm + 10), rng.uniform(-10, Hm + 10)
        a = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(4, 45) * rng.uniform(0.3, 1.0, k)
        p = np.stack([cx + rad * np.cos(a), cy + rad * np.sin(a)], axis=1)
        p = np.round(p * 4) / 4                                    # quarter pixels: exact in the svg text and in fp32
        if rng.random() < 0.5:
            p = np.floor(p) + 0.5                                  # on the sample lattice
        shapes.append((int(rng.integers(0, C_)), p))
    svg = str(tmp_path / "random.svg")
    with open(svg, "w") as f:
        f.write(f'<svg xmlns="http://www.w3.org/2000/svg" width="{Wm}" height="{Hm}">\n')
        for c, p in shapes:
            f.write(f'<polygon fill="{SVG_CLASS_HEX[c]}" points="{" ".join(f"{x:.2f},{y:.2f}" for x, y in p)}"/>\n')
        f.write("</svg>\n")
    home = tmp_path / "home"
    (home / ".ros").mkdir(parents=True)
    excl = [0, 1, 2]
    m = ref.Map.from_path(str(home), svg, np.arange(C_, dtype=np.int32), C_, resolution, [_packed(c) for c in SVG_CLASS_HEX], exclusive=excl)
    layers, mask = m.get()
    order = sorted(range(len(shapes)), key=lambda i: shapes[i][0])
    polys = [np.float32([(x, Hm - y) for x, y in shapes[i][1]]) for i in order]
    binl = orc.raster_polygons(polys, [shapes[i][0] for i in order], Wm, Hm, 0.0, resolution, C_, excl)
    lo, mo = orc.compute_dists(binl, resolution)
    assert layers.shape == lo.shape and same_bits(layers, lo) and np.array_equal(mask, mo)
    from top_down_renderer_b200 import rastercache
    assert np.array_equal(rastercache.load_rasterized_maps(str(tmp_path / "random_raster_cache"), C_), binl)
