"""The ADAPTERS as compiled code: top_down_renderer_b200/adapters/*.cpp are the bodies that replace the reference's .cpp
files — written against the reference's UNCHANGED class declarations (read from /root/reference/include at build time,
with oracle/ref_shim standing in for ROS / Eigen / PCL) and calling the C ABI.  `make -C oracle _adapters` links them with
the CPU stand-in of the C ABI (here) and with libtdr_b200.so (on the GPU box), behind the same C interface as the
reference build, so the reference's own bodies and the adapters can be run side by side."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refbuild as ref
from top_down_renderer_b200 import synth

ANG = np.float32(2 * math.pi / 100)


def _scan():
    cm = synth.make_class_map(260, 300, 4, seed=21)
    pose, heading = synth.default_pose(cm, seed=21)
    pts = synth.make_scan(cm, pose, heading, seed=21, n_rings=32, n_az=256)
    pts[::19, :2] = 0
    return pts, synth.identity_lut(4)


def _check(kind):
    pts, lut = _scan()
    want_ref = {}
    if ref.available():                                             # the reference's own bodies, outside the adapter block
        for res in (4.0, 1.5, 0.5):
            want_ref[res] = (ref.render_polar(pts, res, ANG, 100, 25, lut, 4), ref.render_cart(pts, res, 48, 64, lut, 4))
    with ref.using_adapters(kind):
        for res in (4.0, 1.5, 0.5):
            got = ref.render_polar(pts, res, ANG, 100, 25, lut, 4)
            assert np.array_equal(got, orc.render_polar(pts, res, ANG, 100, 25, lut, 4)) and got.sum() > 1000
            cart = ref.render_cart(pts, res, 48, 64, lut, 4)
            assert np.array_equal(cart, orc.render_cart(pts, res, 48, 64, lut, 4).reshape(cart.shape)) and cart.sum() > 100
            if res in want_ref:
                assert np.array_equal(got, want_ref[res][0]) and np.array_equal(cart, want_ref[res][1])
        # other raster shapes: the size of the caller's images defines the raster
        assert np.array_equal(ref.render_polar(pts, 2.0, np.float32(2 * math.pi / 36), 36, 9, lut, 4),
                              orc.render_polar(pts, 2.0, np.float32(2 * math.pi / 36), 36, 9, lut, 4))
        # the geometric renderers (scan_renderer_polar.cpp:6-81, scan_renderer.cpp:7-53) over the organised 256 x 32 cloud
        gp = pts.copy()
        rng = np.random.default_rng(4)
        gp[:, 2] = -2.0 + rng.normal(0, 0.4, len(gp)).astype(np.float32) * (rng.random(len(gp)) < 0.3)
        for res in (4.0, 1.0):
            geo = ref.render_geometric_polar(gp, 256, 32, res, ANG, 100, 25)
            assert np.array_equal(geo, orc.render_geometric_polar(gp, 256, 32, res, ANG, 100, 25)) and geo[0].sum() > 50 and geo[1].sum() > 50
            gc = ref.render_geometric_cart(gp, 256, 32, res, 96, 80)
            assert np.array_equal(gc, orc.render_geometric_cart(gp, 256, 32, res, 96, 80)) and gc[0].sum() > 50


def _check_map(kind, tmp_path, resolutions=(1.0, 0.5), static_paths=True):
    """TopDownMap / TopDownMapPolar through the adapters: the dynamic path, the gathers, the static constructor and caches"""
    from top_down_renderer_b200 import eigcache, rastercache
    from tests.test_ref_build import SVG_CLASS_HEX, SVG_H, SVG_SHAPES, SVG_W, _packed, _write_svg, same_bits
    C_ = 4
    cm = synth.make_class_map(200, 240, C_, seed=9)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C_)
    with ref.using_adapters(kind):
        for resolution in resolutions:
            seeds = orc.class_image_to_layers(img, lut, C_, resolution)
            lo, mo = orc.compute_dists(seeds, resolution)
            m = ref.Map.from_class_image(img, lut, C_, resolution, center=(5, -2))          # updateMap: a3 + a4 on the device side
            layers, mask = m.get()
            assert same_bits(layers, lo) and np.array_equal(mask, mo) and m.info()[3] and m.info()[4] == (5, -2)
            tab = m.polar_table(100, 25, ANG)                                                # samplePtsPolar
            assert same_bits(tab, orc.polar_table(100, 25, ANG, resolution).reshape(-1, 2))
            go, _ = orc.compute_dists(orc.geo_raster(seeds), resolution)
            for cx, cy in [(120.3, 90.8), (2.0, 3.0), (-40.0, 100.0), (239.5, 199.5)]:
                d, k = m.local_map_polar(cx, cy, 2.0, 1.5)                                   # a7
                do, ko = orc.local_map_polar(lo, mo, resolution, tab, cx, cy, 2.0, 1.5)
                assert same_bits(d, do.reshape(d.shape)) and np.array_equal(k, ko.reshape(k.shape))
                g = m.local_geo_polar(cx, cy, 2.0, 1.5)
                gw, _ = orc.local_map_polar(go, np.zeros_like(mo), resolution, tab, cx, cy, 2.0, 1.5)
                assert same_bits(g, gw.reshape(g.shape))
                assert m.classes_at(cx, cy) == orc.classes_at_point(lo, resolution, int(cx), int(cy))
            d, k = m.local_map_cart(100.0, 80.0, 0.7, 2.0, 40, 31)                           # a8 through the base class
            do, ko = orc.local_map_cart(lo, mo, resolution, 100.0, 80.0, 0.7, 2.0, 40, 31)
            assert same_bits(d, do.reshape(d.shape)) and np.array_equal(k, ko.reshape(k.shape))
        if not static_paths:
            return
        # two maps alive at once: the device holds one, the other is re-installed from its host copies on demand
        a = ref.Map.from_class_image(img, lut, C_, 1.0)
        b = ref.Map.from_class_image(img[::-1].copy(), lut, C_, 1.0)
        a.polar_table(100, 25, ANG), b.polar_table(100, 25, ANG)
        la, ma = orc.compute_dists(orc.class_image_to_layers(img, lut, C_, 1.0), 1.0)
        tab = orc.polar_table(100, 25, ANG, 1.0)
        d, _ = a.local_map_polar(120.0, 90.0, 2.0, 1.5)
        assert same_bits(d, orc.local_map_polar(la, ma, 1.0, tab, 120.0, 90.0, 2.0, 1.5)[0].reshape(d.shape))
        # the static constructor on an svg file, then from each cache
        home = tmp_path / "home"
        (home / ".ros").mkdir(parents=True)
        svg = str(tmp_path / "campus.svg")
        _write_svg(svg)
        colors, excl = [_packed(c) for c in SVG_CLASS_HEX], [0, 1]
        m = ref.Map.from_path(str(home), svg, np.arange(C_, dtype=np.int32), C_, 1.0, colors, exclusive=excl)
        layers, mask, geo = m.get(want_geo=True)
        cls_of = {c: i for i, c in enumerate(SVG_CLASS_HEX)}
        polys = [np.float32([(x, SVG_H - y) for x, y in pts]) for _, pts in SVG_SHAPES]
        pcls = [cls_of[col] for col, _ in SVG_SHAPES]
        order = sorted(range(len(polys)), key=lambda i: pcls[i])
        binl = orc.raster_polygons([polys[i] for i in order], [pcls[i] for i in order], SVG_W, SVG_H, 0.0, 1.0, C_, excl)
        lo, mo = orc.compute_dists(binl, 1.0)
        go, _ = orc.compute_dists(orc.geo_raster(binl), 1.0)
        assert same_bits(layers, lo) and np.array_equal(mask, mo) and same_bits(geo, go) and m.info()[3]
        assert np.array_equal(rastercache.load_rasterized_maps(str(tmp_path / "campus_raster_cache"), C_), binl)
        xc = str(home / ".ros" / "xview_cache")
        assert eigcache.cache_is_valid(xc, svg, C_, 1.0)
        c_layers, c_geo, c_mask = eigcache.load_cache(xc, C_)
        assert same_bits(c_layers, lo) and same_bits(c_geo, go) and np.array_equal(c_mask, mo)
        again = ref.Map.from_path(str(home), svg, np.arange(C_, dtype=np.int32), C_, 1.0, colors, exclusive=excl)      # cache hit
        l2, m2, g2 = again.get(want_geo=True)
        assert same_bits(l2, lo) and same_bits(g2, go) and np.array_equal(m2, mo)
        tab = again.polar_table(100, 25, ANG)
        d, k = again.local_map_polar(SVG_W * 0.45, SVG_H * 0.55, 1.5, 2.0)                   # gathers from the cached fields
        do, ko = orc.local_map_polar(lo, mo, 1.0, tab, SVG_W * 0.45, SVG_H * 0.55, 1.5, 2.0)
        assert same_bits(d, do.reshape(d.shape)) and np.array_equal(k, ko.reshape(k.shape))
        home3 = tmp_path / "home3"
        (home3 / ".ros").mkdir(parents=True)
        ras = ref.Map.from_path(str(home3), str(tmp_path / "campus_raster_cache"), np.arange(C_, dtype=np.int32), C_, 1.0, colors, exclusive=excl)
        l4, m4 = ras.get()
        assert same_bits(l4, lo) and np.array_equal(m4, mo)
    # and the caches the ADAPTER wrote load in the REFERENCE's own code
    if ref.available() and kind == "cpu":
        theirs = ref.Map.from_path(str(home), svg, np.arange(C_, dtype=np.int32), C_, 1.0, colors, exclusive=excl)
        l5, m5, g5 = theirs.get(want_geo=True)
        assert same_bits(l5, lo) and same_bits(g5, go) and np.array_equal(m5, mo)


@pytest.mark.skipif(not ref.adapters_available("cpu"), reason="no /root/reference and no prebuilt adapters")
def test_scan_renderer_adapters_on_the_cpu_standin():
    _check("cpu")


@pytest.mark.skipif(not ref.adapters_available("cpu"), reason="no /root/reference and no prebuilt adapters")
def test_map_adapters_on_the_cpu_standin(tmp_path):
    _check_map("cpu", tmp_path)


def _gpu_adapters():
    if not ref.adapters_available("gpu"):
        pytest.skip("no prebuilt oracle/_ref/libtdr_adapters_gpu.so")
    try:
        ref.adapters("gpu")
    except OSError as e:
        pytest.skip(f"adapters do not load here: {e}")


@pytest.mark.gpu
def test_scan_renderer_adapters_on_the_device():
    _gpu_adapters()
    _check("gpu")


@pytest.mark.gpu
def test_map_adapters_on_the_device(tmp_path):
    _gpu_adapters()
    _check_map("gpu", tmp_path, resolutions=(1.0,))


@pytest.mark.gpu
def test_map_adapters_gathers_at_half_resolution_on_the_device(tmp_path):
    """gathers on a map whose resolution is not 1 (centre / resolution, radius / resolution): the device kernels divide as
    the reference does, but no GPU test fed them such a map before this one — it runs last for that reason"""
    _gpu_adapters()
    _check_map("gpu", tmp_path, resolutions=(0.5,), static_paths=False)


# ---- ParticleFilter / StateParticle through the adapters -------------------------------------------------------------------
REF_BUILD_TESTS = ["test_filter_step_equals_the_reference", "test_theta_search_and_gates_equal_the_reference",
                   "test_nan_weights_take_the_mean_minus_lower_deviation", "test_map_centre_shift_and_metric_initial_position",
                   "test_active_localizer_equals_the_reference", "test_polar_gather_equals_the_reference"]


@pytest.mark.skipif(not ref.adapters_available("cpu"), reason="no /root/reference and no prebuilt adapters")
def test_filter_adapters_pass_the_reference_builds_own_tests_on_the_cpu_standin():
    """the tests that hold the REFERENCE's own ParticleFilter / TopDownMapPolar / ActiveLocalizer against the oracle
    (tests/test_ref_build.py), run unchanged on the adapter classes: same harness, same assertions — initialisation,
    propagate and the engine position bit for bit, weights, adaptive count, resampled states, pose, gates, NaN handling,
    freezeScale, the map-centre shift, the metric initial position, getBestRelPos"""
    import tests.test_ref_build as t
    with ref.using_adapters("cpu"):
        world = t.world._get_wrapped_function()()
        for name in REF_BUILD_TESTS:
            getattr(t, name)(world)
        for args in [(101, 64, (0.0, 0.0, 0.0)), (202, 333, (1.5, -0.7, 0.2))]:
            t.test_filter_steps_over_seeds_sizes_and_motions(world, *args)


def _check_filter_on_device(kind):
    """the adapter ParticleFilter with the DEVICE behind the C ABI: host-side parts (initialisation, RNG stream, adaptive
    count, mirror bookkeeping) bit for bit, device stages to the bars of tests/test_gpu_parity.py"""
    from oracle import numpy_twin as twin
    from tests.common import make_world, N_R, N_THETA, rel_err
    wd = make_world(h=300, w=400, C=4, seed=11, res=2.0)
    seed, N = 23, 400
    kw = dict(fixed_scale=2.0, init_pos_px=(float(wd.pose[0]), float(wd.pose[1])), init_pos_px_cov=8.0,
              init_pos_deg_theta=math.degrees(wd.heading), init_pos_deg_cov=4.0)
    with ref.using_adapters(kind):
        m = ref.Map.from_class_image(wd.img, wd.lut, wd.C, 1.0, center=(wd.w // 2, wd.h // 2))
        m.polar_table(N_THETA, N_R, np.float32(2 * math.pi / N_THETA))
        layers, mask = m.get()
        assert np.array_equal(layers.view(np.uint32), wd.layers.view(np.uint32)) and np.array_equal(mask, wd.mask)
        f = ref.Filter(m, N, seed, regularization=0.7, pos_cov=0.15, theta_cov=0.004, **kw)
        st0, _, _ = f.get()
        so, frozen, _, used = orc.init_particles(seed, wd.layers, 1.0, (wd.w // 2, wd.h // 2), N, **kw)
        assert np.array_equal(st0, so) and f.scale_frozen() == frozen and f.engine_peek() == orc.engine_peek(seed, used)
        f.propagate(0.4, 0.05, 0.01)
        st1, ld1, _ = f.get()
        want, last_o, z, used_p = orc.propagate(so, 0.4, 0.05, 0.01, True, 0.15, 0.004, seed, discard=used)
        tw, last_t = twin.propagate_with_z(so, 0.4, 0.05, 0.01, True, 0.15, 0.004, z)
        assert f.engine_peek() == orc.engine_peek(seed, used + used_p)                       # the same variates were drawn
        for k in ("dx_m", "dy_m", "theta"):
            assert np.abs(st1[k] - want[k]).max() <= 1e-6, k                                 # 1 ulp of cosf / sinf at most
            assert np.array_equal(st1[k], tw[k]) or np.array_equal(st1[k], want[k]), k       # the kernel's form or glibc's
        for k in ("init_x_px", "init_y_px", "scale", "have_init"):
            assert np.array_equal(st1[k], so[k]), k
        assert np.abs(ld1 - last_o).max() <= 1e-6
        _, _, covs = f.gmm()
        cov4 = np.zeros((1, 4, 4), np.float32)
        cov4[0, :3, :3] = covs[0]
        M = orc.adaptive_count(cov4, N, N)
        f.update(wd.scan, wd.res)
        scored, ld_s, raw = f.get(scored_set=True)
        st_o = st1.copy()
        raw_o = orc.score_all(st_o, wd.fp, wd.layers, wd.mask, 1.0, wd.tab, N_THETA, N_R, wd.scan, wd.res, wd.thetas, wd.shifts)
        assert rel_err(raw, raw_o).max() <= 1e-5                                             # a9, a10
        wn = f.weights()
        wn_o, _, _ = orc.normalize(raw.copy(), ld_s)                                         # a11, stage-wise on the device's raw weights
        assert rel_err(wn, wn_o).max() <= 1e-6
        cur, _, _ = f.get()
        assert len(cur) == M == f.num_particles()
        u = orc.uniform_draw(seed, discard=used + used_p)
        idx = orc.resample_fast(wn, u, M)                                                    # a12, stage-wise on the device's weights
        for k in ("init_x_px", "init_y_px", "dx_m", "dy_m", "theta", "scale", "have_init"):
            assert np.array_equal(cur[k], scored[k][idx]), k
        assert f.engine_peek() == orc.engine_peek(seed, used + used_p + 1)
        mean, cov, ml, _ = f.pose()                                                          # a13
        mo, _ = orc.mean_cov(cur)
        assert abs(mean[0] - mo[0]) <= 0.002 and abs(mean[1] - mo[1]) <= 0.002 and abs(mean[2] - mo[2]) <= math.radians(0.01)
        mlo, _ = orc.ml_cov(scored, int(np.argmax(wn)))
        assert np.array_equal(ml, mlo)                                                       # the ML particle is a host object


@pytest.mark.skipif(not ref.adapters_available("cpu"), reason="no /root/reference and no prebuilt adapters")
def test_filter_adapter_device_checks_hold_on_the_cpu_standin():
    _check_filter_on_device("cpu")                      # the GPU test's own assertions, proven where they can run today


@pytest.mark.gpu
def test_filter_adapters_on_the_device():
    _gpu_adapters()
    _check_filter_on_device("gpu")
