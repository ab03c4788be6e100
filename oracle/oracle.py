"""ctypes front-end for oracle/tdr_oracle.cpp (TEST INFRASTRUCTURE ONLY).

The oracle is the CPU restatement of the reference hot path; see the header of
tdr_oracle.cpp for what pins it (the reference's own sources built in oracle/_ref, cv2, libm / libstdc++, twins).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tdr_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class OrcState(C.Structure):
    _fields_ = [("init_x_px", C.c_float), ("init_y_px", C.c_float), ("dx_m", C.c_float),
                ("dy_m", C.c_float), ("theta", C.c_float), ("scale", C.c_float),
                ("have_init", C.c_uint8), ("pad", C.c_uint8 * 3)]


STATE_DTYPE = np.dtype([("init_x_px", "<f4"), ("init_y_px", "<f4"), ("dx_m", "<f4"), ("dy_m", "<f4"),
                        ("theta", "<f4"), ("scale", "<f4"), ("have_init", "u1"), ("pad", "u1", (3,))])
assert STATE_DTYPE.itemsize == 28


class OrcFilterParams(C.Structure):
    _fields_ = [("regularization", C.c_float), ("force_on_map", C.c_int), ("fixed_scale", C.c_float),
                ("scale_log_min", C.c_float), ("scale_log_max", C.c_float), ("map_width", C.c_float),
                ("map_height", C.c_float), ("num_classes", C.c_int), ("class_weights", C.c_float * 16)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_cost_for_shift.restype = C.c_float
        _lib.orc_compute_weight.restype = C.c_float
        _lib.orc_uniform_draw.restype = C.c_float
        _lib.orc_normalize.restype = C.c_long
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def make_params(num_classes, regularization=0.7, class_weights=None, force_on_map=False, fixed_scale=2.0,
                scale_log_min=-0.1, scale_log_max=1.0, map_width=0.0, map_height=0.0):
    fp = OrcFilterParams()
    fp.regularization = regularization
    fp.force_on_map = int(force_on_map)
    fp.fixed_scale = fixed_scale
    fp.scale_log_min = scale_log_min
    fp.scale_log_max = scale_log_max
    fp.map_width = map_width
    fp.map_height = map_height
    fp.num_classes = num_classes
    cw = class_weights if class_weights is not None else [1.0] * num_classes
    for i, v in enumerate(cw):
        fp.class_weights[i] = v
    return fp


# ---- a1 / a2 -------------------------------------------------------------------------------
def render_polar(pts: np.ndarray, res, ang_res, n_theta, n_r, lut, C_, intensity_off=16):
    """pts: (n, stride/4) float32 AoS (PointXYZI = 8 floats); returns (C, n_r, n_theta) float32
    whose memory is C column-major n_theta x n_r images (index [c, r, theta])."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    out = np.empty((C_, n_r, n_theta), dtype=np.float32)
    lib().orc_render_polar(_p(pts, c_u8_p), C.c_int(pts.strides[0]), C.c_int(intensity_off), C.c_long(pts.shape[0]),
                           C.c_float(res), C.c_float(ang_res), n_theta, n_r, _p(lut, c_int_p), len(lut), C_,
                           _p(out, c_float_p))
    return out


def render_cart(pts, res, rows, cols, lut, C_, intensity_off=16):
    """returns (C, cols, rows): column-major rows x cols images (index [c, x, y])."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    out = np.empty((C_, cols, rows), dtype=np.float32)
    lib().orc_render_cart(_p(pts, c_u8_p), C.c_int(pts.strides[0]), C.c_int(intensity_off), C.c_long(pts.shape[0]),
                          C.c_float(res), rows, cols, _p(lut, c_int_p), len(lut), C_, _p(out, c_float_p))
    return out


def render_geometric_polar(pts, width, height, res, ang_res, n_theta, n_r):
    """ScanRendererPolar::renderGeometricTopDown (scan_renderer_polar.cpp:6-81) over an organised cloud
    (pts[row * width + col]); returns (2, n_r, n_theta): [0] flat ground fill, [1] obstacle steps."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.shape[0] == width * height
    out = np.empty((2, n_r, n_theta), dtype=np.float32)
    lib().orc_render_geometric_polar(_p(pts, c_u8_p), C.c_int(pts.strides[0]), width, height, C.c_float(res), C.c_float(ang_res),
                                     n_theta, n_r, _p(out, c_float_p))
    return out


def render_geometric_cart(pts, width, height, res, rows, cols):
    """ScanRenderer::renderGeometricTopDown (scan_renderer.cpp:7-53); returns (2, cols, rows)"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    assert pts.shape[0] == width * height
    out = np.empty((2, cols, rows), dtype=np.float32)
    lib().orc_render_geometric_cart(_p(pts, c_u8_p), C.c_int(pts.strides[0]), width, height, C.c_float(res), rows, cols,
                                    _p(out, c_float_p))
    return out


# ---- a3 / a4 / a5 ----------------------------------------------------------------------------
def map_dims(h_img, w_img, res):
    r, c = C.c_int(), C.c_int()
    lib().orc_map_dims(h_img, w_img, C.c_float(res), C.byref(r), C.byref(c))
    return r.value, c.value


def class_image_to_layers(img: np.ndarray, lut, C_, res=1.0):
    """img: (H, W) uint8 row-major.  Returns layers (C, cols, rows) float32 = column-major rows x cols."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    rows, cols = map_dims(img.shape[0], img.shape[1], res)
    layers = np.empty((C_, cols, rows), dtype=np.float32)
    lib().orc_class_image_to_layers(_p(img, c_u8_p), img.shape[0], img.shape[1], C.c_int(img.strides[0]),
                                    _p(lut, c_int_p), len(lut), C_, C.c_float(res), _p(layers, c_float_p))
    return layers


def compute_dists(layers: np.ndarray, resolution=1.0):
    """layers (C, cols, rows) binary float32 (copied).  Returns (dist layers, mask (cols, rows) uint8)."""
    layers = np.array(layers, dtype=np.float32, order="C", copy=True)
    C_, cols, rows = layers.shape
    mask = np.empty((cols, rows), dtype=np.uint8)
    lib().orc_compute_dists(_p(layers, c_float_p), rows, cols, C_, C.c_float(resolution), _p(mask, c_u8_p))
    return layers, mask


def edt_sq(bin_img: np.ndarray):
    b = np.ascontiguousarray(bin_img, dtype=np.uint8)
    n1, n0 = b.shape
    out = np.empty((n1, n0), dtype=np.int64)
    lib().orc_edt_sq(_p(b, c_u8_p), n0, n1, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def geo_raster(class_layers):
    cl = np.ascontiguousarray(class_layers, dtype=np.float32)
    C_, cols, rows = cl.shape
    geo = np.empty((2, cols, rows), dtype=np.float32)
    lib().orc_geo_raster(_p(cl, c_float_p), rows, cols, C_, _p(geo, c_float_p))
    return geo


# ---- a6 / a7 / a8 -----------------------------------------------------------------------------
def polar_table(n_theta, n_r, ang_res, resolution=1.0):
    tab = np.empty((n_theta * n_r, 2), dtype=np.float32)
    lib().orc_polar_table(n_theta, n_r, C.c_float(ang_res), C.c_float(resolution), _p(tab, c_float_p))
    return tab


def local_map_polar(layers, mask, resolution, tab, cx, cy, scale, res):
    C_, cols, rows = layers.shape
    P = tab.shape[0]
    d = np.empty((C_, P), dtype=np.float32)
    m = np.empty((P,), dtype=np.uint8)
    lib().orc_local_map_polar(_p(layers, c_float_p), _p(mask, c_u8_p), rows, cols, C_, C.c_float(resolution),
                              _p(tab, c_float_p), P, C.c_float(cx), C.c_float(cy), C.c_float(scale),
                              C.c_float(res), _p(d, c_float_p), _p(m, c_u8_p))
    return d, m


def local_map_cart(layers, mask, resolution, cx, cy, rot, res, out_rows, out_cols):
    C_, cols, rows = layers.shape
    P = out_rows * out_cols
    d = np.empty((C_, P), dtype=np.float32)
    m = np.empty((P,), dtype=np.uint8)
    lib().orc_local_map_cart(_p(layers, c_float_p), _p(mask, c_u8_p), rows, cols, C_, C.c_float(resolution),
                             C.c_float(cx), C.c_float(cy), C.c_float(rot), C.c_float(res), out_rows, out_cols,
                             _p(d, c_float_p), _p(m, c_u8_p))
    return d, m


# ---- a9 / a10 ---------------------------------------------------------------------------------
def search_list(n_theta=100):
    th = np.zeros(64, dtype=np.float32)
    sh = np.zeros(64, dtype=np.int32)
    n = lib().orc_search_list(n_theta, _p(th, c_float_p), _p(sh, c_int_p), 64)
    return th[:n].copy(), sh[:n].copy()


def rot_to_shift(rot, n_theta=100):
    return lib().orc_rot_to_shift(C.c_float(rot), n_theta)


def cost_for_shift(scan, classes, known, n_theta, n_r, class_weights, shift):
    scan = np.ascontiguousarray(scan, dtype=np.float32)
    classes = np.ascontiguousarray(classes, dtype=np.float32)
    known = np.ascontiguousarray(known, dtype=np.float32)
    cw = np.ascontiguousarray(class_weights, dtype=np.float32)
    return lib().orc_cost_for_shift(_p(scan, c_float_p), _p(classes, c_float_p), _p(known, c_float_p), n_theta, n_r,
                                    scan.shape[0], _p(cw, c_float_p), int(shift))


def score_all(states, fp, layers, mask, resolution, tab, n_theta, n_r, scan, res, thetas, shifts,
              geo_layers=None, n_threads=0):
    """states: STATE_DTYPE array (modified in place: theta / have_init).  Returns weights float32."""
    assert states.dtype == STATE_DTYPE and states.flags.c_contiguous
    C_, cols, rows = layers.shape
    scan = np.ascontiguousarray(scan, dtype=np.float32)
    thetas = np.ascontiguousarray(thetas, dtype=np.float32)
    shifts = np.ascontiguousarray(shifts, dtype=np.int32)
    w = np.empty(len(states), dtype=np.float32)
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    geo = _p(geo_layers, c_float_p) if geo_layers is not None else None
    lib().orc_score_all(states.ctypes.data_as(C.POINTER(OrcState)), C.c_long(len(states)), C.byref(fp),
                        _p(layers, c_float_p), _p(mask, c_u8_p), geo, rows, cols, C.c_float(resolution),
                        _p(tab, c_float_p), n_theta, n_r, _p(scan, c_float_p), C.c_float(res),
                        _p(thetas, c_float_p), _p(shifts, c_int_p), len(shifts), _p(w, c_float_p), n_threads)
    return w


def cost_grid(centers, scale, fp, layers, mask, resolution, tab, n_theta, n_r, scan, res, shifts, n_threads=0):
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    shifts = np.ascontiguousarray(shifts, dtype=np.int32)
    scan = np.ascontiguousarray(scan, dtype=np.float32)
    C_, cols, rows = layers.shape
    out = np.empty((centers.shape[0], len(shifts)), dtype=np.float32)
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    lib().orc_cost_grid(_p(centers, c_float_p), C.c_long(centers.shape[0]), C.c_float(scale), C.byref(fp),
                        _p(layers, c_float_p), _p(mask, c_u8_p), rows, cols, C.c_float(resolution),
                        _p(tab, c_float_p), n_theta, n_r, _p(scan, c_float_p), C.c_float(res),
                        _p(shifts, c_int_p), len(shifts), _p(out, c_float_p), n_threads)
    return out


# ---- a11 / a12 / a13 --------------------------------------------------------------------------
def normalize(w, last_dist):
    w = np.array(w, dtype=np.float32, copy=True)
    ld = np.ascontiguousarray(last_dist, dtype=np.float32)
    stats = np.zeros(6, dtype=np.float32)
    arg = lib().orc_normalize(_p(w, c_float_p), _p(ld, c_float_p), C.c_long(len(w)), _p(stats, c_float_p))
    return w, int(arg), stats


def resample_literal(w, shift, M):
    w = np.ascontiguousarray(w, dtype=np.float32)
    idx = np.empty(M, dtype=np.int32)
    lib().orc_resample_literal(_p(w, c_float_p), C.c_long(len(w)), C.c_float(shift), int(M), _p(idx, c_int_p))
    return idx


def resample_fast(w, shift, M, want_prefix=False):
    w = np.ascontiguousarray(w, dtype=np.float32)
    idx = np.empty(M, dtype=np.int32)
    pre = np.empty(len(w), dtype=np.float32) if want_prefix else None
    lib().orc_resample_fast(_p(w, c_float_p), C.c_long(len(w)), C.c_float(shift), int(M), _p(idx, c_int_p),
                            _p(pre, c_float_p) if want_prefix else None)
    return (idx, pre) if want_prefix else idx


def uniform_draw(seed, discard=0):
    """the resampling offset of particle_filter.cpp:172-173 from mt19937(seed) after `discard` earlier outputs"""
    lib().orc_uniform_draw_from.restype = C.c_float
    return float(lib().orc_uniform_draw_from(C.c_uint32(seed), C.c_uint64(discard)))


def engine_peek(seed, discard):
    """raw 32-bit output number `discard` of std::mt19937(seed)"""
    lib().orc_engine_peek.restype = C.c_uint32
    return int(lib().orc_engine_peek(C.c_uint32(seed), C.c_uint64(discard)))


def mean_cov(states):
    mean = np.zeros(4, dtype=np.float32)
    cov = np.zeros(16, dtype=np.float32)
    lib().orc_mean_cov(states.ctypes.data_as(C.POINTER(OrcState)), C.c_long(len(states)), _p(mean, c_float_p),
                       _p(cov, c_float_p))
    return mean, cov.reshape(4, 4)


def ml_cov(states, argmax):
    ml = np.zeros(4, dtype=np.float32)
    cov = np.zeros(16, dtype=np.float32)
    lib().orc_ml_cov(states.ctypes.data_as(C.POINTER(OrcState)), C.c_long(len(states)), C.c_long(argmax),
                     _p(ml, c_float_p), _p(cov, c_float_p))
    return ml, cov.reshape(4, 4)


def refine_bin(xy, cls, res, cx, cy, width, height, C_):
    xy = np.ascontiguousarray(xy, dtype=np.float32)
    cls = np.ascontiguousarray(cls, dtype=np.int32)
    out = np.empty((C_, height, width), dtype=np.uint8)
    lib().orc_refine_bin(_p(xy, c_float_p), _p(cls, c_int_p), C.c_long(len(cls)), C.c_float(res), C.c_float(cx),
                         C.c_float(cy), width, height, C_, _p(out, c_u8_p))
    return out


def propagate(states, tx, ty, omega, scale_freeze, pos_cov, theta_cov, seed, discard=None):
    """StateParticle::propagate over a particle set with ONE shared mt19937(seed), in particle order
    (state_particle.cpp:57-78, particle_filter.cpp:86-92).  Returns (new states, last_dist, z) where z[n, 4] are the
    standard normal variates behind the draws (theta, dx, dy, scale).  With `discard` (the engine outputs earlier calls
    consumed) the engine resumes there and the outputs consumed here are returned as a fourth value."""
    st = np.ascontiguousarray(states.copy())
    n = len(st)
    last = np.empty(n, dtype=np.float32)
    z = np.empty((n, 4), dtype=np.float32)
    draws = C.c_uint64(0)
    lib().orc_propagate_from(st.ctypes.data_as(C.c_void_p), _p(last, c_float_p), C.c_long(n), C.c_float(tx), C.c_float(ty),
                             C.c_float(omega), int(bool(scale_freeze)), C.c_float(pos_cov), C.c_float(theta_cov),
                             C.c_uint32(seed), C.c_uint64(discard or 0), _p(z, c_float_p), C.byref(draws))
    return (st, last, z) if discard is None else (st, last, z, int(draws.value))


def raster_polygons(polys, poly_class, map_w, map_h, rot, resolution, C_, exclusive):
    """getRasterMap + getClasses (top_down_map.cpp:328-408): polys = list of (n_k, 2) float32 vertex arrays (x, y), their
    flattened classes, the svg size -> C binary layers (C, cols, rows) = col-major rows x cols, 0 inside, 1 outside"""
    start = np.zeros(len(polys) + 1, dtype=np.int32)
    start[1:] = np.cumsum([len(p) for p in polys])
    verts = (np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in polys]) if polys
             else np.zeros((0, 2), np.float32))
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    pc = np.ascontiguousarray(poly_class, dtype=np.int32)
    ex = np.ascontiguousarray(exclusive, dtype=np.int32)
    rows, cols = int(map_h / resolution), int(map_w / resolution)
    out = np.empty((C_, cols, rows), dtype=np.float32)
    lib().orc_raster_polygons(_p(verts, c_float_p), _p(start, c_int_p), _p(pc, c_int_p), len(polys), int(map_w), int(map_h),
                              C.c_float(rot), C.c_float(resolution), C_, _p(ex, c_int_p), len(ex), _p(out, c_float_p))
    return out


def active_best_rel_pos(layers, mask, resolution, tab, n_theta, n_r, preds):
    """ActiveLocalizer::getBestRelPos (active_localizer.cpp:45-82) -> ((dist, theta), best_diff)"""
    C_, cols, rows = layers.shape
    preds = np.ascontiguousarray(preds, dtype=np.float32).reshape(-1, 3)
    rel = np.zeros(2, dtype=np.float32)
    best = C.c_float()
    lib().orc_active_best_rel_pos(_p(layers, c_float_p), _p(mask, c_u8_p), rows, cols, C_, C.c_float(resolution),
                                  _p(np.ascontiguousarray(tab, dtype=np.float32), c_float_p), n_theta, n_r, _p(preds, c_float_p),
                                  len(preds), _p(rel, c_float_p), C.byref(best))
    return (float(rel[0]), float(rel[1])), best.value


def gmm_samples(states, num_samples):
    """the num_samples x 4 double matrix computeGMM hands to cv::ml::EM (particle_filter.cpp:262-272)"""
    st = np.ascontiguousarray(states)
    out = np.empty((num_samples, 4), dtype=np.float64)
    lib().orc_gmm_samples(st.ctypes.data_as(C.c_void_p), C.c_long(len(st)), int(num_samples), out.ctypes.data_as(C.c_void_p))
    return out


def adaptive_count(covs, last_num_particles, max_num_particles):
    """num_particles_ of the next resampling from the GMM covariances (particle_filter.cpp:151-158); covs: (k, 4, 4)"""
    cv = np.ascontiguousarray(covs, dtype=np.float32).reshape(-1, 16)
    lib().orc_adaptive_count.restype = C.c_int
    return int(lib().orc_adaptive_count(_p(cv, c_float_p), len(cv), int(last_num_particles), int(max_num_particles)))


class OrcInitParams(C.Structure):
    _fields_ = [("init_pos_px_x", C.c_float), ("init_pos_px_y", C.c_float), ("init_pos_px_cov", C.c_float),
                ("init_pos_m_x", C.c_float), ("init_pos_m_y", C.c_float), ("init_pos_deg_theta", C.c_float),
                ("init_pos_deg_cov", C.c_float), ("fixed_scale", C.c_float)]


def classes_at_point(layers, resolution, px, py):
    """TopDownMap::getClassesAtPoint(Vector2i) (top_down_map.cpp:159-170) on distance layers (C, cols, rows)"""
    L = np.ascontiguousarray(layers, dtype=np.float32)
    C_, cols, rows = L.shape
    lib().orc_classes_at_point.restype = C.c_uint
    bits = lib().orc_classes_at_point(_p(L, c_float_p), rows, cols, C_, C.c_float(resolution), int(px), int(py))
    return [c for c in range(C_) if bits >> c & 1]


def init_particles(seed, layers, resolution, map_center, max_n, init_pos_px=(-1.0, -1.0), init_pos_px_cov=-1.0,
                   init_pos_m=(math.inf, math.inf), init_pos_deg_theta=math.inf, init_pos_deg_cov=10.0, fixed_scale=-1.0):
    """ParticleFilter::initializeParticles (particle_filter.cpp:19-84) from std::mt19937(seed) on distance layers
    (C, cols, rows).  Returns (states, scale_frozen, (init_pos_px_x, init_pos_px_y) as the filter ends up with, the number
    of engine outputs consumed — `discard` of the calls that continue on the filter's shared engine)."""
    L = np.ascontiguousarray(layers, dtype=np.float32)
    C_, cols, rows = L.shape
    ip = OrcInitParams(init_pos_px[0], init_pos_px[1], init_pos_px_cov, init_pos_m[0], init_pos_m[1], init_pos_deg_theta,
                       init_pos_deg_cov, fixed_scale)
    out = np.zeros(max(int(max_n), 1), dtype=STATE_DTYPE)
    frozen = C.c_int(0)
    px = np.zeros(2, dtype=np.float32)
    draws = C.c_uint64(0)
    lib().orc_init_particles.restype = C.c_long
    n = lib().orc_init_particles(C.c_uint32(seed), _p(L, c_float_p), rows, cols, C_, C.c_float(resolution), int(map_center[0]),
                                 int(map_center[1]), C.byref(ip), int(max_n), out.ctypes.data_as(C.c_void_p), C.byref(frozen),
                                 _p(px, c_float_p), C.byref(draws))
    return out[:n].copy(), bool(frozen.value), (float(px[0]), float(px[1])), int(draws.value)


def freeze_scale(states):
    """ParticleFilter::freezeScale (particle_filter.cpp:343-357); returns (states with the common scale, that scale)"""
    st = np.ascontiguousarray(states).copy()
    lib().orc_freeze_scale.restype = C.c_float
    g = lib().orc_freeze_scale(st.ctypes.data_as(C.c_void_p), C.c_long(len(st)))
    return st, float(g)
