"""TEST INFRASTRUCTURE ONLY — `oracle/_ref/libtdr_ref.so`: seven translation units of the reference compiled UNMODIFIED
from /root/reference/src (scan_renderer.cpp, scan_renderer_polar.cpp, top_down_map.cpp, top_down_map_polar.cpp,
state_particle.cpp, particle_filter.cpp, active_localizer.cpp) against the stand-in headers in oracle/ref_shim/ (the reference's real
dependencies — ROS, Eigen, OpenCV, PCL — are not installed), behind the C interface of oracle/ref_shim/ref_harness.cpp.
oracle/ref_shim/README.md says what that pins and what it cannot.  The library is built only where /root/reference
exists (this container); the prebuilt file travels to the GPU box.  Nothing outside tests/ loads it."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

from .oracle import STATE_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libtdr_ref.so")
REFERENCE = "/root/reference"
UNITS = ["scan_renderer", "scan_renderer_polar", "top_down_map", "top_down_map_polar", "state_particle", "particle_filter", "active_localizer"]


def available() -> bool:
    return os.path.exists(SO) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(force: bool = False) -> str:
    """compiles the reference's sources where they lie; outputs only under oracle/_ref/ (git-ignored)"""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref"] + (["-B"] if force else []))
    if not os.path.exists(SO):
        raise FileNotFoundError(f"{SO}: not built and {REFERENCE} is not present")
    return SO


class RefFilterParams(C.Structure):
    _fields_ = [("pos_cov", C.c_float), ("theta_cov", C.c_float), ("regularization", C.c_float),
                ("init_pos_px_x", C.c_float), ("init_pos_px_y", C.c_float), ("init_pos_px_cov", C.c_float),
                ("init_pos_m_x", C.c_float), ("init_pos_m_y", C.c_float), ("init_pos_deg_theta", C.c_float),
                ("init_pos_deg_cov", C.c_float), ("force_on_map", C.c_int), ("fixed_scale", C.c_float),
                ("scale_log_min", C.c_float), ("scale_log_max", C.c_float), ("class_weights", C.c_float * 16),
                ("n_class_weights", C.c_int)]


_lib = None
_f32p, _u8p, _i32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_double)


def _prepare(lb, with_filter=True):
    lb.ref_map_create.restype = C.c_void_p
    lb.ref_map_from_class_image.restype = C.c_void_p
    lb.ref_map_from_path.restype = C.c_void_p
    lb.ref_map_classes_at.restype = C.c_uint
    if with_filter:
        lb.ref_filter_create.restype = C.c_void_p
        for name in ("ref_filter_count", "ref_filter_get", "ref_filter_weights"):
            getattr(lb, name).restype = C.c_long
        lb.ref_filter_scale.restype = C.c_float
        lb.ref_filter_engine_peek.restype = C.c_uint32
    return lb


_forced = None


def lib():
    """the reference build — or, inside `with using_adapters(kind):`, the adapters behind the same C interface"""
    global _lib
    if _forced is not None:
        return _forced
    if _lib is None:
        _lib = _prepare(C.CDLL(build()))
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def render_polar(pts, res, ang_res, n_theta, n_r, lut, num_classes):
    """ScanRendererPolar::renderSemanticTopDown (scan_renderer_polar.cpp:83-109) -> (C, n_r, n_theta) = col-major images"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.shape[1] == 8
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    out = np.zeros((num_classes, n_r, n_theta), dtype=np.float32)
    lib().ref_render_polar(_p(pts, _f32p), C.c_long(len(pts)), C.c_float(res), C.c_float(ang_res), n_theta, n_r, _p(lut, _i32p),
                           len(lut), num_classes, _p(out, _f32p))
    return out


def render_geometric_polar(pts, width, height, res, ang_res, n_theta, n_r):
    """ScanRendererPolar::renderGeometricTopDown (scan_renderer_polar.cpp:6-81) -> (2, n_r, n_theta)"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    out = np.zeros((2, n_r, n_theta), dtype=np.float32)
    lib().ref_render_geometric_polar(_p(pts, _f32p), width, height, C.c_float(res), C.c_float(ang_res), n_theta, n_r, _p(out, _f32p))
    return out


def render_geometric_cart(pts, width, height, res, rows, cols):
    """ScanRenderer::renderGeometricTopDown (scan_renderer.cpp:7-53) -> (2, cols, rows)"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    out = np.zeros((2, cols, rows), dtype=np.float32)
    lib().ref_render_geometric_cart(_p(pts, _f32p), width, height, C.c_float(res), rows, cols, _p(out, _f32p))
    return out


def render_cart(pts, res, rows, cols, lut, num_classes):
    """ScanRenderer::renderSemanticTopDown (scan_renderer.cpp:55-78) -> (C, cols, rows)"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    out = np.zeros((num_classes, cols, rows), dtype=np.float32)
    lib().ref_render_cart(_p(pts, _f32p), C.c_long(len(pts)), C.c_float(res), rows, cols, _p(lut, _i32p), len(lut), num_classes,
                          _p(out, _f32p))
    return out


class Map:
    """TopDownMapPolar.  Map(...) installs given distance fields / mask / offset table (inputs); Map.from_class_image and
    Map.from_path run the reference's own map code (top_down_map.cpp) on an image / a map file."""

    @classmethod
    def _wrap(cls, handle, n_theta, n_r):
        self = cls.__new__(cls)
        self.h = C.c_void_p(handle)
        self.n_theta, self.n_r = n_theta, n_r
        self.rows, self.cols, self.C, _, _ = self.info()
        return self

    @classmethod
    def from_class_image(cls, img, lut, num_classes, resolution, center=(0, 0), n_theta=100, n_r=25):
        """TopDownMapPolar(params) + updateMap(image, centre): loadCompressedRasterMap + computeDists (:116-157, :289-326)"""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        lut = np.ascontiguousarray(lut, dtype=np.int32)
        h = lib().ref_map_from_class_image(_p(img, _u8p), img.shape[0], img.shape[1], img.shape[1], _p(lut, _i32p), len(lut),
                                           int(num_classes), C.c_float(resolution), int(center[0]), int(center[1]))
        return cls._wrap(h, n_theta, n_r)

    @classmethod
    def from_path(cls, home, map_path, lut, num_classes, resolution, class_colors, exclusive=(), n_theta=100, n_r=25):
        """the static-map constructor (:9-64) with $HOME = home for ~/.ros/xview_cache"""
        lut = np.ascontiguousarray(lut, dtype=np.int32)
        col = np.ascontiguousarray(class_colors, dtype=np.uint32)
        ex = np.ascontiguousarray(exclusive, dtype=np.int32)
        assert len(col) == len(lut)
        h = lib().ref_map_from_path(home.encode(), map_path.encode(), _p(lut, _i32p), len(lut), int(num_classes), C.c_float(resolution),
                                    col.ctypes.data_as(C.POINTER(C.c_uint32)), _p(ex, _i32p), len(ex))
        return cls._wrap(h, n_theta, n_r)

    def info(self):
        r, c, k, have = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        ctr = (C.c_int * 2)()
        lib().ref_map_info(self.h, C.byref(r), C.byref(c), C.byref(k), C.byref(have), ctr)
        return r.value, c.value, k.value, bool(have.value), (ctr[0], ctr[1])

    def get(self, want_geo=False):
        """-> (class_maps_ (C, cols, rows), class_mask_ (cols, rows)[, geo_maps_ (2, cols, rows)])"""
        rows, cols, k, _, _ = self.info()
        layers, mask = np.zeros((k, cols, rows), dtype=np.float32), np.zeros((cols, rows), dtype=np.uint8)
        geo = np.zeros((2, cols, rows), dtype=np.float32) if want_geo else None
        lib().ref_map_get(self.h, _p(layers, _f32p), _p(mask, _u8p), _p(geo, _f32p) if want_geo else None)
        return (layers, mask, geo) if want_geo else (layers, mask)

    def build_geo_from_binary(self, binary_layers):
        """getGeoRasterMap + computeDists (:410-427, :47-58) on the given binary class maps; the distance fields are put back"""
        dist, _ = self.get()
        b = np.ascontiguousarray(binary_layers, dtype=np.float32)
        lib().ref_map_set_class_maps(self.h, _p(b, _f32p), self.rows, self.cols, self.C)
        lib().ref_map_build_geo(self.h)
        lib().ref_map_set_class_maps(self.h, _p(dist, _f32p), self.rows, self.cols, self.C)
        return self.get(want_geo=True)[2]

    def polar_table(self, n_theta, n_r, ang_res):
        """samplePtsPolar (top_down_map_polar.cpp:7-19) through samplePts (top_down_map.cpp:367-389)"""
        tab = np.zeros(2 * n_theta * n_r, dtype=np.float32)
        lib().ref_map_polar_table(self.h, n_theta, n_r, C.c_float(ang_res), _p(tab, _f32p))
        self.n_theta, self.n_r = n_theta, n_r
        return tab.reshape(-1, 2)

    def set_polar_table(self, tab, n_theta, n_r):
        t = np.ascontiguousarray(tab, dtype=np.float32).reshape(-1)
        lib().ref_map_set_polar_table(self.h, _p(t, _f32p), n_theta, n_r)
        self.n_theta, self.n_r = n_theta, n_r

    def classes_at(self, x, y, as_float=False):
        bits = lib().ref_map_classes_at(self.h, int(as_float), C.c_float(x), C.c_float(y))
        return [c for c in range(self.C) if bits >> c & 1]

    def local_map_cart(self, cx, cy, rot, res, rows, cols):
        d, m = np.zeros((self.C, cols, rows), dtype=np.float32), np.zeros((cols, rows), dtype=np.uint8)
        lib().ref_map_local_cart(self.h, C.c_float(cx), C.c_float(cy), C.c_float(rot), C.c_float(res), rows, cols, _p(d, _f32p), _p(m, _u8p))
        return d, m

    def __init__(self, layers, mask, resolution, tab, n_theta, n_r, geo=None, center=(0, 0)):
        layers = np.ascontiguousarray(layers, dtype=np.float32)
        self.C, self.cols, self.rows = layers.shape
        self.n_theta, self.n_r = n_theta, n_r
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        tab = np.ascontiguousarray(tab, dtype=np.float32).reshape(-1)
        geo_p = _p(np.ascontiguousarray(geo, dtype=np.float32), _f32p) if geo is not None else None
        self.h = C.c_void_p(lib().ref_map_create(_p(layers, _f32p), _p(mask, _u8p), geo_p, self.rows, self.cols, self.C,
                                                 C.c_float(resolution), _p(tab, _f32p), n_theta, n_r, int(center[0]), int(center[1])))

    def local_map_polar(self, cx, cy, scale, res):
        d = np.zeros((self.C, self.n_r, self.n_theta), dtype=np.float32)
        m = np.zeros((self.n_r, self.n_theta), dtype=np.uint8)
        lib().ref_map_local_polar(self.h, C.c_float(cx), C.c_float(cy), C.c_float(scale), C.c_float(res), self.n_theta, self.n_r,
                                  _p(d, _f32p), _p(m, _u8p))
        return d, m

    def local_geo_polar(self, cx, cy, scale, res):
        g = np.zeros((2, self.n_r, self.n_theta), dtype=np.float32)
        lib().ref_map_local_geo_polar(self.h, C.c_float(cx), C.c_float(cy), C.c_float(scale), C.c_float(res), self.n_theta, self.n_r,
                                      _p(g, _f32p))
        return g

    def active_best_rel_pos(self, preds):
        p = np.ascontiguousarray(preds, dtype=np.float32)
        rel = np.zeros(2, dtype=np.float32)
        lib().ref_active_best_rel_pos(self.h, _p(p, _f32p), len(p), _p(rel, _f32p))
        return float(rel[0]), float(rel[1])


class Filter:
    """ParticleFilter (particle_filter.cpp) on a Map; the engine is std::mt19937(seed)"""

    def __init__(self, ref_map: Map, N, seed, initialize=True, pos_cov=0.3, theta_cov=0.0314, regularization=0.15,
                 init_pos_px=(-1.0, -1.0), init_pos_px_cov=-1.0, init_pos_m=(math.inf, math.inf), init_pos_deg_theta=math.inf,
                 init_pos_deg_cov=10.0, force_on_map=False, fixed_scale=-1.0, scale_log_min=-0.1, scale_log_max=1.0,
                 class_weights=None):
        self.map = ref_map
        cw = list(class_weights) if class_weights is not None else [1.0] * ref_map.C
        rp = RefFilterParams(pos_cov, theta_cov, regularization, init_pos_px[0], init_pos_px[1], init_pos_px_cov, init_pos_m[0],
                             init_pos_m[1], init_pos_deg_theta, init_pos_deg_cov, int(force_on_map), fixed_scale, scale_log_min,
                             scale_log_max, (C.c_float * 16)(*(cw + [0.0] * (16 - len(cw)))), len(cw))
        self.h = C.c_void_p(lib().ref_filter_create(ref_map.h, int(N), C.byref(rp), C.c_uint32(seed), int(bool(initialize))))

    def count(self):
        return int(lib().ref_filter_count(self.h))

    def num_particles(self):
        return int(lib().ref_filter_num_particles(self.h))

    def get(self, scored_set=False):
        """-> (states, last_dist, raw weights) of the current set, or (after an update) of the set that was scored"""
        n = int(lib().ref_filter_get(self.h, int(scored_set), None, None, None, C.c_long(0)))
        st, ld, w = np.zeros(n, dtype=STATE_DTYPE), np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        lib().ref_filter_get(self.h, int(scored_set), st.ctypes.data_as(C.c_void_p), _p(ld, _f32p), _p(w, _f32p), C.c_long(n))
        return st, ld, w

    def set(self, states, last_dist=None):
        st = np.ascontiguousarray(states)
        ld = _p(np.ascontiguousarray(last_dist, dtype=np.float32), _f32p) if last_dist is not None else None
        lib().ref_filter_set(self.h, st.ctypes.data_as(C.c_void_p), ld, C.c_long(len(st)))

    def weights(self):
        n = int(lib().ref_filter_weights(self.h, None, C.c_long(0)))
        w = np.zeros(n, dtype=np.float32)
        lib().ref_filter_weights(self.h, _p(w, _f32p), C.c_long(n))
        return w

    def propagate(self, tx, ty, omega):
        lib().ref_filter_propagate(self.h, C.c_float(tx), C.c_float(ty), C.c_float(omega))

    def update(self, scan, res):
        s = np.ascontiguousarray(scan, dtype=np.float32)
        Cn, n_r, n_theta = s.shape
        lib().ref_filter_update(self.h, _p(s, _f32p), n_theta, n_r, Cn, C.c_float(res))

    def pose(self):
        mean, cov, ml, cov_ml = (np.zeros(k, dtype=np.float32) for k in (4, 16, 4, 16))
        lib().ref_filter_pose(self.h, _p(mean, _f32p), _p(cov, _f32p), _p(ml, _f32p), _p(cov_ml, _f32p))
        return mean, cov.reshape(4, 4), ml, cov_ml.reshape(4, 4)

    def freeze_scale(self):
        lib().ref_filter_freeze_scale(self.h)

    def scale(self):
        return float(lib().ref_filter_scale(self.h))

    def scale_frozen(self):
        return bool(lib().ref_filter_scale_frozen(self.h))

    def update_map(self, img, center):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        lib().ref_filter_update_map(self.h, _p(img, _u8p), img.shape[0], img.shape[1], int(center[0]), int(center[1]))

    def init_px(self):
        px = np.zeros(2, dtype=np.float32)
        lib().ref_filter_init_px(self.h, _p(px, _f32p))
        return float(px[0]), float(px[1])

    def engine_peek(self):
        return int(lib().ref_filter_engine_peek(self.h))

    def gmm(self):
        samples = np.zeros((1000, 4), dtype=np.float64)
        means, covs = np.zeros((8, 3), dtype=np.float32), np.zeros((8, 3, 3), dtype=np.float32)
        n = int(lib().ref_filter_gmm(self.h, _p(samples, _f64p), 1000, _p(means, _f32p), _p(covs, _f32p), 8))
        return samples[:n], means[:1], covs[:1]


# ---- the ADAPTERS (top_down_renderer_b200/adapters/*.cpp): bodies over the C ABI for the reference's unchanged class
# declarations, behind the same C interface (oracle/ref_shim/adapter_harness.cpp)
def adapters_so(kind: str) -> str:
    assert kind in ("cpu", "gpu")
    return os.path.join(_HERE, "_ref", f"libtdr_adapters_{kind}.so")


def adapters_available(kind: str) -> bool:
    return os.path.exists(adapters_so(kind)) or os.path.isdir(os.path.join(REFERENCE, "include"))


_adp = {}


def adapters(kind: str):
    """kind "cpu": linked with the CPU stand-in of the C ABI (the oracle answers); "gpu": with libtdr_b200.so"""
    if kind not in _adp:
        if os.path.isdir(os.path.join(REFERENCE, "include")):
            subprocess.check_call(["make", "-C", _HERE, "-s", "_ref/" + os.path.basename(adapters_so(kind))])
        _adp[kind] = C.CDLL(adapters_so(kind))
    return _adp[kind]


import contextlib  # noqa: E402


@contextlib.contextmanager
def using_adapters(kind: str):
    """inside the block render_polar / render_cart / Map run the ADAPTERS (adapters/*.cpp over the C ABI: "cpu" = the CPU
    stand-in answers, "gpu" = libtdr_b200) through the very harness that drives the reference's own bodies"""
    global _forced
    prev = _forced
    _forced = _prepare(adapters(kind))
    try:
        yield
    finally:
        _forced = prev
