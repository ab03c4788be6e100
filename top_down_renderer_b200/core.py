"""numpy-facing wrapper of the C ABI (include/tdr.h): one Context = one tdr_ctx.

Array conventions follow the reference's Eigen types:
  * class / distance layers: numpy (C, cols, rows) C-contiguous == C column-major rows x cols arrays
  * polar images: (C, n_r, n_theta) C-contiguous == C column-major n_theta x n_r images
  * states: structured array of STATE_DTYPE (28-byte State records)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import TdrFilterParams, check

STATE_DTYPE = np.dtype([("init_x_px", "<f4"), ("init_y_px", "<f4"), ("dx_m", "<f4"), ("dy_m", "<f4"),
                        ("theta", "<f4"), ("scale", "<f4"), ("have_init", "u1"), ("pad", "u1", (3,))])

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)
_b = C.POINTER(C.c_uint8)


def _pf(a):
    return a.ctypes.data_as(_f)


def _pi(a):
    return a.ctypes.data_as(_i)


def _pb(a):
    return a.ctypes.data_as(_b)


class Context:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.tdr_create(C.byref(h), int(device)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tdr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- context
    def sync(self):
        check(self._lib.tdr_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.tdr_stream(self._h) or 0)

    def launch_count(self) -> int:
        n = C.c_int64()
        check(self._lib.tdr_launch_count(self._h, C.byref(n)))
        return n.value

    def set_score_impl(self, impl):
        """0 auto, 1 CUDA cores only, 2 tensor cores whenever usable"""
        check(self._lib.tdr_set_score_impl(self._h, int(impl)))

    def profile_enable(self, on=True):
        check(self._lib.tdr_profile_enable(self._h, int(bool(on))))

    def profile_stage_ms(self):
        """device ms of the last step's stages: render, score, normalise, resample (synchronises)."""
        ms = (C.c_float * 4)()
        check(self._lib.tdr_profile_stage_ms(self._h, ms))
        return [float(v) for v in ms]

    # ---- map
    def map_set_class_image(self, img, flatten_lut, num_classes, resolution=1.0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        lut = np.ascontiguousarray(flatten_lut, dtype=np.int32)
        check(self._lib.tdr_map_set_class_image(self._h, _pb(img), img.shape[0], img.shape[1], C.c_int(img.strides[0]),
                                                _pi(lut), len(lut), int(num_classes), C.c_float(resolution)))

    def map_set_binary_layers(self, layers, resolution=1.0):
        layers = np.ascontiguousarray(layers, dtype=np.float32)
        c, cols, rows = layers.shape
        check(self._lib.tdr_map_set_binary_layers(self._h, _pf(layers), rows, cols, c, C.c_float(resolution)))

    def map_set_dist_layers(self, layers, mask, resolution=1.0):
        layers = np.ascontiguousarray(layers, dtype=np.float32)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        c, cols, rows = layers.shape
        check(self._lib.tdr_map_set_dist_layers(self._h, _pf(layers), _pb(mask), rows, cols, c, C.c_float(resolution)))

    def map_info(self):
        r, c, k, res = C.c_int(), C.c_int(), C.c_int(), C.c_float()
        check(self._lib.tdr_map_info(self._h, C.byref(r), C.byref(c), C.byref(k), C.byref(res)))
        return r.value, c.value, k.value, res.value

    def map_get_layers(self):
        rows, cols, k, _ = self.map_info()
        layers = np.empty((k, cols, rows), dtype=np.float32)
        mask = np.empty((cols, rows), dtype=np.uint8)
        check(self._lib.tdr_map_get_layers(self._h, _pf(layers), _pb(mask)))
        return layers, mask

    def map_get_geo_layers(self):
        rows, cols, _, _ = self.map_info()
        geo = np.empty((2, cols, rows), dtype=np.float32)
        check(self._lib.tdr_map_get_geo_layers(self._h, _pf(geo)))
        return geo

    def map_set_polar_table(self, tab, n_theta, n_r):
        tab = np.ascontiguousarray(tab, dtype=np.float32)
        assert tab.size == 2 * n_theta * n_r
        check(self._lib.tdr_map_set_polar_table(self._h, _pf(tab), int(n_theta), int(n_r)))
        self._P = n_theta * n_r

    def map_set_polygons(self, polys, poly_class, map_w, map_h, rot, num_classes, resolution, exclusive, want_layers=True):
        """vector map (getRasterMap + getClasses + computeDists); returns the binary class layers (C, cols, rows)"""
        start = np.zeros(len(polys) + 1, dtype=np.int32)
        start[1:] = np.cumsum([len(p) for p in polys])
        verts = (np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in polys]) if len(polys)
                 else np.zeros((0, 2), np.float32))
        verts = np.ascontiguousarray(verts, dtype=np.float32)
        pc = np.ascontiguousarray(poly_class, dtype=np.int32)
        ex = np.ascontiguousarray(exclusive, dtype=np.int32)
        rows, cols = int(map_h / resolution), int(map_w / resolution)
        out = np.empty((num_classes, cols, rows), dtype=np.float32) if want_layers else None
        check(self._lib.tdr_map_set_polygons(self._h, _pf(verts), _pi(start), _pi(pc), len(polys), int(map_w), int(map_h),
                                             C.c_float(rot), int(num_classes), C.c_float(resolution), _pi(ex), len(ex),
                                             _pf(out) if want_layers else None))
        return out

    def map_local_polar(self, centers_xy, scale, res):
        centers = np.ascontiguousarray(centers_xy, dtype=np.float32).reshape(-1, 2)
        n = centers.shape[0]
        _, _, k, _ = self.map_info()
        d = np.empty((n, k, self._P), dtype=np.float32)
        m = np.empty((n, self._P), dtype=np.uint8)
        check(self._lib.tdr_map_local_polar(self._h, _pf(centers), n, C.c_float(scale), C.c_float(res), _pf(d), _pb(m)))
        return d, m

    def map_set_geo_dist_layers(self, geo):
        geo = np.ascontiguousarray(geo, dtype=np.float32)
        check(self._lib.tdr_map_set_geo_dist_layers(self._h, _pf(geo)))

    def map_local_geo_polar(self, centers_xy, scale, res):
        centers = np.ascontiguousarray(centers_xy, dtype=np.float32).reshape(-1, 2)
        n = centers.shape[0]
        g = np.empty((n, 2, self._P), dtype=np.float32)
        check(self._lib.tdr_map_local_geo_polar(self._h, _pf(centers), n, C.c_float(scale), C.c_float(res), _pf(g)))
        return g

    def active_best_rel_pos(self, preds_xyt):
        """ActiveLocalizer::getBestRelPos -> ((dist, theta), best_diff)"""
        preds = np.ascontiguousarray(preds_xyt, dtype=np.float32).reshape(-1, 3)
        rel = np.zeros(2, dtype=np.float32)
        best = C.c_float()
        check(self._lib.tdr_active_best_rel_pos(self._h, _pf(preds), preds.shape[0], _pf(rel), C.byref(best)))
        return (float(rel[0]), float(rel[1])), best.value

    def map_local_cart(self, cx, cy, rot, res, rows, cols):
        _, _, k, _ = self.map_info()
        d = np.empty((k, rows * cols), dtype=np.float32)
        m = np.empty((rows * cols,), dtype=np.uint8)
        check(self._lib.tdr_map_local_cart(self._h, C.c_float(cx), C.c_float(cy), C.c_float(rot), C.c_float(res),
                                           int(rows), int(cols), _pf(d), _pb(m)))
        return d, m

    # ---- scan
    def scan_set_points(self, pts, intensity_off=16):
        """pts: (n, k) float32 AoS (PointXYZI: k = 8).  Asynchronous: keep `pts` alive until sync()."""
        assert pts.dtype == np.float32 and pts.flags.c_contiguous
        self._pts_keepalive = pts
        check(self._lib.tdr_scan_set_points(self._h, C.c_void_p(pts.ctypes.data), C.c_int(pts.shape[1] * 4),
                                            C.c_int(intensity_off), C.c_int64(pts.shape[0])))

    def scan_set_points_ptr(self, ptr, stride, intensity_off, n):
        check(self._lib.tdr_scan_set_points(self._h, C.c_void_p(ptr), C.c_int(stride), C.c_int(intensity_off), C.c_int64(n)))

    def scan_set_lut(self, lut, num_classes):
        lut = np.ascontiguousarray(lut, dtype=np.int32)
        check(self._lib.tdr_scan_set_lut(self._h, _pi(lut), len(lut), int(num_classes)))
        self._scan_C = int(num_classes)

    def scan_render_polar(self, res, ang_res, n_theta, n_r, want=True):
        out = np.empty((self._scan_C, n_r, n_theta), dtype=np.float32) if want else None
        check(self._lib.tdr_scan_render_polar(self._h, C.c_float(res), C.c_float(ang_res), int(n_theta), int(n_r),
                                              _pf(out) if want else None))
        return out

    def scan_render_cart(self, res, rows, cols):
        out = np.empty((self._scan_C, cols, rows), dtype=np.float32)
        check(self._lib.tdr_scan_render_cart(self._h, C.c_float(res), int(rows), int(cols), _pf(out)))
        return out

    def scan_render_geometric_polar(self, width, height, res, ang_res, n_theta, n_r):
        """ScanRendererPolar::renderGeometricTopDown over the resident points as a width x height organised cloud"""
        out = np.empty((2, n_r, n_theta), dtype=np.float32)
        check(self._lib.tdr_scan_render_geometric_polar(self._h, int(width), int(height), C.c_float(res), C.c_float(ang_res),
                                                        int(n_theta), int(n_r), _pf(out)))
        return out

    def scan_render_geometric_cart(self, width, height, res, rows, cols):
        out = np.empty((2, cols, rows), dtype=np.float32)
        check(self._lib.tdr_scan_render_geometric_cart(self._h, int(width), int(height), C.c_float(res), int(rows), int(cols), _pf(out)))
        return out

    def refine_bin(self, xy, cls, res, cx, cy, width, height, num_classes):
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        cls = np.ascontiguousarray(cls, dtype=np.int32)
        out = np.empty((num_classes, height, width), dtype=np.uint8)
        check(self._lib.tdr_refine_bin(self._h, _pf(xy), _pi(cls), C.c_int64(len(cls)), C.c_float(res), C.c_float(cx),
                                       C.c_float(cy), int(width), int(height), int(num_classes), _pb(out)))
        return out

    # the same in pieces (cfg5: batches larger than one buffer) + the distance-field rebuild on the device
    def refine_begin(self, res, cx, cy, width, height, num_classes):
        self._refine_shape = (int(num_classes), int(height), int(width))
        check(self._lib.tdr_refine_begin(self._h, C.c_float(res), C.c_float(cx), C.c_float(cy), int(width), int(height), int(num_classes)))

    def refine_add(self, xy, cls):
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        cls = np.ascontiguousarray(cls, dtype=np.int32)
        check(self._lib.tdr_refine_add(self._h, _pf(xy), _pi(cls), C.c_int64(len(cls))))

    def refine_add_ptr(self, xy_ptr, cls_ptr, n):
        """host pointers (e.g. pinned torch tensors)"""
        check(self._lib.tdr_refine_add(self._h, C.cast(C.c_void_p(xy_ptr), C.POINTER(C.c_float)),
                                       C.cast(C.c_void_p(cls_ptr), C.POINTER(C.c_int32)), C.c_int64(n)))

    def refine_add_dev(self, xy_ptr, cls_ptr, n):
        check(self._lib.tdr_refine_add_dev(self._h, C.c_void_p(xy_ptr), C.c_void_p(cls_ptr), C.c_int64(n)))

    def refine_counts(self):
        out = np.empty(self._refine_shape, dtype=np.uint8)
        check(self._lib.tdr_refine_counts(self._h, _pb(out)))
        return out

    def refine_rebuild_map(self, resolution):
        check(self._lib.tdr_refine_rebuild_map(self._h, C.c_float(resolution)))

    def scan_set_polar_images(self, imgs):
        imgs = np.ascontiguousarray(imgs, dtype=np.float32)
        c, n_r, n_theta = imgs.shape
        check(self._lib.tdr_scan_set_polar_images(self._h, _pf(imgs), n_theta, n_r, c))

    # ---- filter
    def pf_set_params(self, num_classes, regularization=0.7, class_weights=None, force_on_map=False, fixed_scale=2.0,
                      scale_log_min=-0.1, scale_log_max=1.0):
        p = TdrFilterParams()
        p.regularization = regularization
        p.force_on_map = int(bool(force_on_map))
        p.fixed_scale = fixed_scale
        p.scale_log_min = scale_log_min
        p.scale_log_max = scale_log_max
        p.num_classes = int(num_classes)
        cw = list(class_weights) if class_weights is not None else [1.0] * num_classes
        for k, v in enumerate(cw):
            p.class_weights[k] = v
        check(self._lib.tdr_pf_set_params(self._h, C.byref(p)))

    def pf_set_search(self, thetas, shifts):
        th = np.ascontiguousarray(thetas, dtype=np.float32)
        sh = np.ascontiguousarray(shifts, dtype=np.int32)
        check(self._lib.tdr_pf_set_search(self._h, _pf(th), _pi(sh), len(sh)))

    def pf_set_states(self, states, last_dist=None):
        assert states.dtype == STATE_DTYPE
        states = np.ascontiguousarray(states)
        ld = np.ascontiguousarray(last_dist, dtype=np.float32) if last_dist is not None else None
        check(self._lib.tdr_pf_set_states(self._h, C.c_void_p(states.ctypes.data), _pf(ld) if ld is not None else None,
                                          C.c_int64(len(states))))

    def pf_count(self):
        n = C.c_int64()
        check(self._lib.tdr_pf_count(self._h, C.byref(n)))
        return n.value

    def pf_checkpoint(self):
        check(self._lib.tdr_pf_checkpoint(self._h))

    def pf_restore(self):
        check(self._lib.tdr_pf_restore(self._h))

    def pf_get_states(self, n=None):
        n = self.pf_count() if n is None else n
        st = np.zeros(n, dtype=STATE_DTYPE)
        check(self._lib.tdr_pf_get_states(self._h, C.c_void_p(st.ctypes.data), C.c_int64(n)))
        return st

    def pf_score(self, res, want=True):
        n = self.pf_count()
        w = np.empty(n, dtype=np.float32) if want else None
        check(self._lib.tdr_pf_score(self._h, C.c_float(res), _pf(w) if want else None))
        return w

    def pf_set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float32)
        check(self._lib.tdr_pf_set_weights(self._h, _pf(w), C.c_int64(len(w))))

    def pf_get_weights(self, n):
        w = np.empty(n, dtype=np.float32)
        check(self._lib.tdr_pf_get_weights(self._h, _pf(w), C.c_int64(n)))
        return w

    def pf_normalize(self):
        arg = C.c_int64()
        stats = np.zeros(6, dtype=np.float32)
        check(self._lib.tdr_pf_normalize(self._h, C.byref(arg), _pf(stats)))
        return arg.value, stats

    def pf_resample(self, u, M, want=True):
        idx = np.empty(M, dtype=np.int32) if want else None
        check(self._lib.tdr_pf_resample(self._h, C.c_float(u), C.c_int64(M), _pi(idx) if want else None))
        return idx

    def pf_normalize_resample(self, u, M):
        arg = C.c_int64()
        idx = np.empty(M, dtype=np.int32)
        check(self._lib.tdr_pf_normalize_resample(self._h, C.c_float(u), C.c_int64(M), C.byref(arg), _pi(idx)))
        return arg.value, idx

    def pf_pose(self, want_ml=True):
        mean = np.zeros(4, dtype=np.float32)
        cov = np.zeros(16, dtype=np.float32)
        ml = np.zeros(4, dtype=np.float32)
        cov_ml = np.zeros(16, dtype=np.float32)
        check(self._lib.tdr_pf_pose(self._h, _pf(mean), _pf(cov), _pf(ml) if want_ml else None,
                                    _pf(cov_ml) if want_ml else None))
        return mean, cov.reshape(4, 4), ml, cov_ml.reshape(4, 4)

    def pf_update(self, res, u, M):
        check(self._lib.tdr_pf_update(self._h, C.c_float(res), C.c_float(u), C.c_int64(M)))

    def step(self, res, ang_res, n_theta, n_r, u, M):
        check(self._lib.tdr_step(self._h, C.c_float(res), C.c_float(ang_res), int(n_theta), int(n_r), C.c_float(u),
                                 C.c_int64(M)))

    # ---- multi-GPU shards (device pointers; the host layer issues the all-gather)
    def pf_export_shard(self, dev_ptr, capacity_floats, with_weights=True):
        check(self._lib.tdr_pf_export_shard(self._h, C.c_void_p(dev_ptr), C.c_int64(capacity_floats), int(with_weights)))

    def pf_update_gathered(self, dev_ptr, n_ranks, n_local, u, M, i0, i1):
        check(self._lib.tdr_pf_update_gathered(self._h, C.c_void_p(dev_ptr), int(n_ranks), C.c_int64(n_local),
                                               C.c_float(u), C.c_int64(M), C.c_int64(i0), C.c_int64(i1)))

    # the same update as two collectives (weights + last_dist first, the states while the normalisation runs)
    def pf_export_split(self, wl_ptr, states_ptr):
        check(self._lib.tdr_pf_export_split(self._h, C.c_void_p(wl_ptr), C.c_void_p(states_ptr)))

    def pf_normalize_gathered(self, wl_all_ptr, n_ranks, n_local):
        check(self._lib.tdr_pf_normalize_gathered(self._h, C.c_void_p(wl_all_ptr), int(n_ranks), C.c_int64(n_local)))

    def pf_resample_gathered(self, states_all_ptr, n_ranks, n_local, u, M, i0, i1):
        check(self._lib.tdr_pf_resample_gathered(self._h, C.c_void_p(states_all_ptr), int(n_ranks), C.c_int64(n_local),
                                                 C.c_float(u), C.c_int64(M), C.c_int64(i0), C.c_int64(i1)))

    def pf_pose_gathered(self, dev_ptr, n_ranks, n_local, want_ml=True):
        mean = np.zeros(4, dtype=np.float32)
        cov = np.zeros(16, dtype=np.float32)
        ml = np.zeros(4, dtype=np.float32)
        cov_ml = np.zeros(16, dtype=np.float32)
        check(self._lib.tdr_pf_pose_gathered(self._h, C.c_void_p(dev_ptr), int(n_ranks), C.c_int64(n_local), _pf(mean),
                                             _pf(cov), _pf(ml) if want_ml else None, _pf(cov_ml) if want_ml else None))
        return mean, cov.reshape(4, 4), ml, cov_ml.reshape(4, 4)

    # ---- the sharded filter below the ABI (csrc/shard.cu): NCCL + peer-mapped state slots inside the library
    def pf_set_shard_count(self, n_ranks):
        check(self._lib.tdr_pf_set_shard_count(self._h, int(n_ranks)))

    @staticmethod
    def shard_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(_lib.load().tdr_shard_unique_id(buf))
        return bytes(buf)

    def shard_init(self, rank, world, unique_id: bytes, particles_per_rank):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self._lib.tdr_shard_init(self._h, int(rank), int(world), buf, C.c_int64(particles_per_rank)))

    def shard_step(self, res, ang_res, n_theta, n_r, u, M_total):
        check(self._lib.tdr_shard_step(self._h, C.c_float(res), C.c_float(ang_res), int(n_theta), int(n_r), C.c_float(u), C.c_int64(M_total)))

    def shard_pose(self, want_ml=True):
        mean = np.zeros(4, dtype=np.float32)
        cov = np.zeros(16, dtype=np.float32)
        ml = np.zeros(4, dtype=np.float32)
        cov_ml = np.zeros(16, dtype=np.float32)
        check(self._lib.tdr_shard_pose(self._h, _pf(mean), _pf(cov), _pf(ml) if want_ml else None, _pf(cov_ml) if want_ml else None))
        return mean, cov.reshape(4, 4), ml, cov_ml.reshape(4, 4)

    def shard_finalize(self):
        self._lib.tdr_shard_finalize(self._h)

    # ---- grid
    def grid_costs(self, centers_xy, scale, res, shifts, want=True):
        centers = np.ascontiguousarray(centers_xy, dtype=np.float32).reshape(-1, 2)
        sh = np.ascontiguousarray(shifts, dtype=np.int32)
        out = np.empty((centers.shape[0], len(sh)), dtype=np.float32) if want else None
        check(self._lib.tdr_grid_costs(self._h, _pf(centers), C.c_int64(centers.shape[0]), C.c_float(scale),
                                       C.c_float(res), _pi(sh), len(sh), _pf(out) if want else None))
        return out

    def grid_run_resident(self, n, scale, res, shifts):
        """asynchronous re-run over the resident centres (no H2D / D2H)"""
        sh = np.ascontiguousarray(shifts, dtype=np.int32)
        check(self._lib.tdr_grid_costs(self._h, None, C.c_int64(n), C.c_float(scale), C.c_float(res), _pi(sh), len(sh), None))

    def grid_set_costs_buffer(self, dev_ptr, capacity_floats):
        check(self._lib.tdr_grid_set_costs_buffer(self._h, C.c_void_p(dev_ptr), C.c_int64(capacity_floats)))

    def grid_best_dev(self, dev_ptr, n):
        c = C.c_float()
        i = C.c_int64()
        check(self._lib.tdr_grid_best_dev(self._h, C.c_void_p(dev_ptr), C.c_int64(n), C.byref(c), C.byref(i)))
        return c.value, i.value

    # fused weight all-gather over peer memory
    def grid_peer_alloc(self, n_floats):
        p = C.c_void_p()
        h = (C.c_uint8 * 64)()
        check(self._lib.tdr_grid_peer_alloc(self._h, C.c_int64(n_floats), C.byref(p), h))
        return p.value, bytes(h)

    def grid_peer_open(self, handle: bytes):
        p = C.c_void_p()
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        check(self._lib.tdr_grid_peer_open(self._h, h, C.byref(p)))
        return p.value

    def grid_peer_set(self, ptrs, row_offset):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        check(self._lib.tdr_grid_peer_set(self._h, arr, len(ptrs), C.c_int64(row_offset)))

    def grid_peer_clear(self):
        check(self._lib.tdr_grid_peer_clear(self._h))

    def grid_peer_exchange(self) -> int:
        """arg-min reduction + barrier of the fused grid over peer memory (no collective library); the packed key"""
        k = C.c_uint64()
        check(self._lib.tdr_grid_peer_exchange(self._h, C.byref(k)))
        return int(k.value)

    def copy_from_device(self, dev_ptr, n_floats):
        """debug / test helper: D2H of n floats from a raw device pointer (synchronises the context stream first)"""
        self.sync()
        out = np.empty(n_floats, dtype=np.float32)
        libcudart = C.CDLL("libcudart.so.12")
        rc = libcudart.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(dev_ptr), C.c_size_t(n_floats * 4), C.c_int(2))
        if rc != 0:
            raise RuntimeError(f"cudaMemcpy D2H failed: {rc}")
        return out

    def grid_best(self):
        c = C.c_float()
        i = C.c_int64()
        check(self._lib.tdr_grid_best(self._h, C.byref(c), C.byref(i)))
        return c.value, i.value

    # motion model (StateParticle::propagate, state_particle.cpp:57-78)
    def pf_propagate(self, trans, omega, scale_freeze, pos_cov, theta_cov, z):
        """z: (n, 4) standard normal variates (theta, dx, dy, scale) — the parity path"""
        z = np.ascontiguousarray(z, dtype=np.float32).reshape(-1, 4)
        check(self._lib.tdr_pf_propagate(self._h, C.c_float(trans[0]), C.c_float(trans[1]), C.c_float(omega), int(bool(scale_freeze)),
                                         C.c_float(pos_cov), C.c_float(theta_cov), _pf(z), C.c_int64(z.shape[0])))

    def pf_propagate_rng(self, trans, omega, scale_freeze, pos_cov, theta_cov, seed, step, want_z=False):
        """device RNG (Philox + Box-Muller); returns the variates used when want_z"""
        z = np.empty((self.pf_count(), 4), dtype=np.float32) if want_z else None
        check(self._lib.tdr_pf_propagate_rng(self._h, C.c_float(trans[0]), C.c_float(trans[1]), C.c_float(omega), int(bool(scale_freeze)),
                                             C.c_float(pos_cov), C.c_float(theta_cov), C.c_uint64(seed), C.c_uint64(step),
                                             _pf(z) if want_z else None))
        return z

    def pf_gmm_samples(self, num_samples):
        out = np.empty((num_samples, 4), dtype=np.float64)
        check(self._lib.tdr_pf_gmm_samples(self._h, int(num_samples), out.ctypes.data_as(C.c_void_p)))
        return out

    def pf_get_last_dist(self):
        out = np.empty(self.pf_count(), dtype=np.float32)
        check(self._lib.tdr_pf_get_last_dist(self._h, _pf(out), C.c_int64(len(out))))
        return out

    def grid_best_key(self):
        """packed (min cost, first flat index) of the last tensor-core grid launch (synchronises)"""
        k = C.c_uint64()
        check(self._lib.tdr_grid_best_key(self._h, C.byref(k)))
        return k.value

    def grid_key_decode(self, key):
        c = C.c_float()
        i = C.c_int64()
        check(self._lib.tdr_grid_key_decode(C.c_uint64(key & 0xFFFFFFFFFFFFFFFF), C.byref(c), C.byref(i)))
        return c.value, i.value

    def dev_ptr(self, which):
        p = C.c_void_p()
        n = C.c_int64()
        check(self._lib.tdr_dev_ptr(self._h, int(which), C.byref(p), C.byref(n)))
        return (p.value or 0), n.value

    def pf_set_weights_dev(self, ptr, n):
        check(self._lib.tdr_pf_set_weights_dev(self._h, C.c_void_p(ptr), C.c_int64(n)))
