"""The reference's map cache: `~/.ros/xview_cache/` as TopDownMap::saveCachedMaps /
loadCacheMetaData / loadCachedMaps write and read it (top_down_map.cpp:226-286), so that caches written by either
implementation load in the other.

File format `.eig` (write_binary / read_binary, top_down_map.h:29-50): Eigen::Index rows, Eigen::Index cols (two
little-endian int64), then rows*cols scalars in COLUMN-major order — float32 for `class_map<c>.eig` (the distance
fields after computeDists) and `geo_map<0|1>.eig`, int8 for `class_mask.eig`.  `cached_data.txt`: the map path, the
number of classes, the resolution (operator<< of a float), one per line; a cache is valid when the path and the class
count are equal and the resolution differs by at most 0.01 (:233-239).

Arrays here use this repo's layout for col-major rows x cols images: numpy shape (cols, rows), C-contiguous."""
from __future__ import annotations

import os

import numpy as np


def write_eig(path: str, a: np.ndarray) -> None:
    """a: (cols, rows) C-contiguous == the column-major rows x cols matrix"""
    a = np.ascontiguousarray(a)
    cols, rows = a.shape
    with open(path, "wb") as f:
        f.write(np.array([rows, cols], dtype="<i8").tobytes())
        f.write(a.tobytes())


def read_eig(path: str, dtype) -> np.ndarray:
    with open(path, "rb") as f:
        rows, cols = (int(v) for v in np.frombuffer(f.read(16), dtype="<i8"))
        a = np.frombuffer(f.read(rows * cols * np.dtype(dtype).itemsize), dtype=dtype)
    if a.size != rows * cols:
        raise ValueError(f"{path}: truncated ({a.size} of {rows * cols} scalars)")
    return a.reshape(cols, rows).copy()


def save_cache(cache_dir: str, map_path: str, layers: np.ndarray, geo: np.ndarray, mask: np.ndarray, resolution: float) -> None:
    """layers (C, cols, rows) float32 distance fields, geo (2, cols, rows), mask (cols, rows) 0/1"""
    os.makedirs(cache_dir, exist_ok=True)
    with open(os.path.join(cache_dir, "cached_data.txt"), "w") as f:
        f.write(f"{map_path}\n{layers.shape[0]}\n{_cxx_float(resolution)}\n")
    for c in range(layers.shape[0]):
        write_eig(os.path.join(cache_dir, f"class_map{c}.eig"), layers[c].astype(np.float32))
    for c in range(2):
        write_eig(os.path.join(cache_dir, f"geo_map{c}.eig"), geo[c].astype(np.float32))
    write_eig(os.path.join(cache_dir, "class_mask.eig"), mask.astype(np.int8))


def cache_is_valid(cache_dir: str, map_path: str, num_classes: int, resolution: float) -> bool:
    try:
        with open(os.path.join(cache_dir, "cached_data.txt")) as f:
            lines = f.read().split("\n")
        return lines[0] == map_path and int(lines[1]) == num_classes and abs(float(lines[2]) - resolution) <= 0.01
    except (OSError, ValueError, IndexError):
        return False


def load_cache(cache_dir: str, num_classes: int):
    layers = np.stack([read_eig(os.path.join(cache_dir, f"class_map{c}.eig"), np.float32) for c in range(num_classes)])
    geo = np.stack([read_eig(os.path.join(cache_dir, f"geo_map{c}.eig"), np.float32) for c in range(2)])
    mask = read_eig(os.path.join(cache_dir, "class_mask.eig"), np.int8).astype(np.uint8)
    return layers, geo, mask


def _cxx_float(v: float) -> str:
    """std::ostream << float: %g with 6 significant digits"""
    return "%g" % np.float32(v)
