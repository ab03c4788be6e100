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


def _check_map(kind, tmp_path):
    """TopDownMap / TopDownMapPolar through the adapters: the dynamic path, the gathers, the static constructor and caches"""
    from top_down_renderer_b200 import eigcache, rastercache
    from tests.test_ref_build import SVG_CLASS_HEX, SVG_H, SVG_SHAPES, SVG_W, _packed, _write_svg, same_bits
    C_ = 4
    cm = synth.make_class_map(200, 240, C_, seed=9)
    img, lut = synth.to_cv_image(cm), synth.identity_lut(C_)
    with ref.using_adapters(kind):
        for resolution in (1.0, 0.5):
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
    _check_map("gpu", tmp_path)
