"""ctypes binding of libtdr_b200.so (the C ABI of include/tdr.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable,
every operation raises.  The library is built in-tree by ``build()`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libtdr_b200.so")
HOSTMATH_PATH = os.path.join(_PKG, "libtdr_hostmath.so")

TDR_OK, TDR_EINVAL, TDR_ENOGPU, TDR_ECUDA, TDR_ESTATE, TDR_EUNSUPPORTED = 0, -1, -2, -3, -4, -5
TDR_BUF_WEIGHTS, TDR_BUF_GRID_COSTS, TDR_BUF_STATES_SOA, TDR_BUF_SCAN_IMAGES = 0, 1, 2, 3


class TdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtdr_b200 error {code}: {msg}")
        self.code = code


class TdrState(C.Structure):
    """State, include/top_down_render/state_particle.h:9-17"""
    _fields_ = [("init_x_px", C.c_float), ("init_y_px", C.c_float), ("dx_m", C.c_float), ("dy_m", C.c_float),
                ("theta", C.c_float), ("scale", C.c_float), ("have_init", C.c_uint8), ("pad_", C.c_uint8 * 3)]


class TdrFilterParams(C.Structure):
    _fields_ = [("regularization", C.c_float), ("force_on_map", C.c_int32), ("fixed_scale", C.c_float),
                ("scale_log_min", C.c_float), ("scale_log_max", C.c_float), ("num_classes", C.c_int32),
                ("class_weights", C.c_float * 16)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libtdr_b200.so (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", _CSRC, "-j8"] + (["-B"] if force else [])
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libtdr_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def build_hostmath() -> str:
    r = subprocess.run(["make", "-C", _CSRC, "../libtdr_hostmath.so"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libtdr_hostmath.so failed:\n" + r.stdout + r.stderr)
    return HOSTMATH_PATH


_lib = None

# every symbol include/tdr.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "tdr_abi_version", "tdr_last_error", "tdr_create", "tdr_destroy", "tdr_sync", "tdr_stream", "tdr_launch_count", "tdr_set_score_impl", "tdr_profile_enable", "tdr_profile_stage_ms",
    "tdr_map_set_class_image", "tdr_map_set_binary_layers", "tdr_map_set_dist_layers", "tdr_map_set_polygons", "tdr_map_get_layers",
    "tdr_map_info", "tdr_map_get_geo_layers", "tdr_map_set_polar_table", "tdr_map_local_polar", "tdr_map_local_cart", "tdr_map_local_geo_polar", "tdr_map_set_geo_dist_layers", "tdr_active_best_rel_pos",
    "tdr_scan_set_points", "tdr_scan_set_lut", "tdr_scan_render_polar", "tdr_scan_render_cart", "tdr_scan_render_geometric_polar", "tdr_scan_render_geometric_cart",
    "tdr_scan_set_polar_images", "tdr_refine_bin", "tdr_refine_begin", "tdr_refine_add", "tdr_refine_add_dev", "tdr_refine_counts", "tdr_refine_rebuild_map", "tdr_pf_set_params", "tdr_pf_set_search", "tdr_pf_set_states", "tdr_pf_get_states",
    "tdr_pf_count", "tdr_pf_get_prev_states", "tdr_pf_keep_raw_weights", "tdr_pf_get_raw_weights", "tdr_pf_checkpoint", "tdr_pf_restore", "tdr_pf_score", "tdr_pf_set_weights", "tdr_pf_get_weights", "tdr_pf_normalize", "tdr_pf_resample", "tdr_pf_normalize_resample",
    "tdr_pf_pose", "tdr_pf_update", "tdr_step", "tdr_grid_costs", "tdr_grid_best", "tdr_grid_set_costs_buffer", "tdr_grid_best_dev", "tdr_grid_best_key", "tdr_grid_key_decode", "tdr_grid_peer_alloc", "tdr_grid_peer_open",
    "tdr_grid_peer_set", "tdr_grid_peer_clear", "tdr_dev_ptr",
    "tdr_pf_propagate", "tdr_pf_propagate_dev", "tdr_pf_propagate_rng", "tdr_pf_get_last_dist", "tdr_pf_gmm_samples",
    "tdr_pf_set_weights_dev", "tdr_pf_export_shard", "tdr_pf_update_gathered", "tdr_pf_export_split", "tdr_pf_normalize_gathered", "tdr_pf_resample_gathered", "tdr_pf_pose_gathered", "tdr_grid_peer_exchange", "tdr_pf_set_shard_count", "tdr_shard_unique_id", "tdr_shard_init", "tdr_shard_step", "tdr_shard_pose", "tdr_shard_finalize",
]


def load():
    """dlopen libtdr_b200.so.  Raises if it has not been built — the product path never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with top_down_renderer_b200.build() "
                           "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.tdr_last_error.restype = C.c_char_p
    lib.tdr_stream.restype = C.c_void_p
    lib.tdr_stream.argtypes = [C.c_void_p]
    lib.tdr_destroy.restype = None
    lib.tdr_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        raise TdrError(code, load().tdr_last_error().decode("utf-8", "replace"))
