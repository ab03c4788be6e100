"""Seeded synthetic inputs of the shapes BASELINE.json names.

Pure numpy; shared by tests/, bench.py and __graft_entry__.smoke().  Nothing here is
on the product path — it only manufactures maps, scans and particle sets.

Conventions (reference file:line):
  * class image: uint8 row-major cv::Mat as received by TopDownMap::updateMap
    (top_down_map.cpp:146); image row 0 is the TOP, map row 0 the BOTTOM (vertical
    flip in loadCompressedRasterMap, top_down_map.cpp:137).
  * scan: pcl::PointXYZI AoS, 32 B/point = 8 floats: x y z pad intensity pad pad pad
    (scan_renderer.h:12); intensity carries the raw class id.
  * State: 28-byte struct (state_particle.h:9-17).
"""
from __future__ import annotations

import math

import numpy as np

STATE_DTYPE = np.dtype([("init_x_px", "<f4"), ("init_y_px", "<f4"), ("dx_m", "<f4"), ("dy_m", "<f4"),
                        ("theta", "<f4"), ("scale", "<f4"), ("have_init", "u1"), ("pad", "u1", (3,))])
UNKNOWN = 255
ROAD = 1  # particles are initialised on class 1 (state_particle.cpp:28-31)


def identity_lut(num_classes: int) -> np.ndarray:
    """256-entry flatten LUT, -1 default (top_down_render.cpp:57-62)."""
    lut = -np.ones(256, dtype=np.int32)
    lut[:num_classes] = np.arange(num_classes, dtype=np.int32)
    return lut


def make_class_map(h: int, w: int, num_classes: int, seed: int = 1234, unknown_frac: float = 0.10) -> np.ndarray:
    """Class-index image in MAP orientation (row 0 = bottom).  Use to_cv_image() for the loader's input."""
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), dtype=np.uint8)  # class 0 = open terrain
    # coarse land-use blocks of the non-road classes
    n_blocks = max(8, (h * w) // 20000)
    for _ in range(n_blocks):
        c = int(rng.integers(2, num_classes)) if num_classes > 2 else 0
        bh, bw = int(rng.integers(8, 80)), int(rng.integers(8, 80))
        y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        m[y0:y0 + bh, x0:x0 + bw] = c
    # unknown regions (~unknown_frac of the pixels), a few large blobs + many small ones
    target = unknown_frac * h * w
    covered = 0
    while covered < target:
        bh, bw = int(rng.integers(10, max(12, h // 12))), int(rng.integers(10, max(12, w // 12)))
        y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        m[y0:y0 + bh, x0:x0 + bw] = UNKNOWN
        covered += min(bh, h - y0) * min(bw, w - x0)
    # road grid (class 1) drawn last so it is connected
    y = int(rng.integers(20, 60))
    while y < h - 8:
        wd = int(rng.integers(4, 10))
        m[y:y + wd, :] = ROAD
        y += int(rng.integers(60, 180))
    x = int(rng.integers(20, 60))
    while x < w - 8:
        wd = int(rng.integers(4, 10))
        m[:, x:x + wd] = ROAD
        x += int(rng.integers(60, 180))
    return m


def to_cv_image(class_map: np.ndarray) -> np.ndarray:
    """Map orientation -> cv::Mat orientation (row 0 = top)."""
    return np.ascontiguousarray(class_map[::-1, :])


def ring_ranges(n_rings: int = 64, sensor_h: float = 2.0, max_range: float = 100.0) -> np.ndarray:
    """Fixed elevation table: ring k looks down by elev_k; range = h / tan(elev), clipped."""
    elev = np.deg2rad(np.linspace(22.5, 0.6, n_rings))
    return np.minimum(sensor_h / np.tan(elev), max_range).astype(np.float32)


def make_scan(class_map: np.ndarray, pose_xy_px, heading: float, scale_px_per_m: float = 2.0,
              n_rings: int = 64, n_az: int = 1024, seed: int = 1234, drop_frac: float = 0.10) -> np.ndarray:
    """64x1024 PointXYZI scan (n, 8) float32 taken from a ground-truth pose in the map."""
    rng = np.random.default_rng(seed + 7)
    h, w = class_map.shape
    rho = ring_ranges(n_rings)                                   # metres
    az = (np.arange(n_az, dtype=np.float64) * (2 * math.pi / n_az))
    rr, aa = np.meshgrid(rho.astype(np.float64), az, indexing="ij")
    rr = rr * (1.0 + 0.01 * rng.standard_normal(rr.shape))       # range noise
    xs = rr * np.cos(aa)
    ys = rr * np.sin(aa)
    # world offset = R(heading) * (x, y); map px = pose + scale * world
    wx = math.cos(heading) * xs - math.sin(heading) * ys
    wy = math.sin(heading) * xs + math.cos(heading) * ys
    px = np.rint(pose_xy_px[0] + scale_px_per_m * wx).astype(np.int64)
    py = np.rint(pose_xy_px[1] + scale_px_per_m * wy).astype(np.int64)
    inb = (px >= 0) & (px < w) & (py >= 0) & (py < h)
    cls = np.full(px.shape, UNKNOWN, dtype=np.uint8)
    cls[inb] = class_map[py[inb], px[inb]]
    pts = np.zeros((n_rings * n_az, 8), dtype=np.float32)
    pts[:, 0] = xs.reshape(-1)
    pts[:, 1] = ys.reshape(-1)
    pts[:, 2] = -2.0
    pts[:, 4] = cls.reshape(-1).astype(np.float32)
    drop = rng.random(pts.shape[0]) < drop_frac                  # invalid returns: x = y = 0 (skipped, :95)
    pts[drop, 0:3] = 0
    return pts


def road_pixels(class_map: np.ndarray) -> np.ndarray:
    return np.flatnonzero(class_map == ROAD)


def particles_tracking(n: int, pose_xy_px, heading: float, scale: float = 2.0, sigma_px: float = 5.0,
                       sigma_deg: float = 5.0, seed: int = 1234):
    rng = np.random.default_rng(seed + 11)
    st = np.zeros(n, dtype=STATE_DTYPE)
    dx = rng.normal(0, 1.0, n).astype(np.float32)                # odometry metres
    dy = rng.normal(0, 1.0, n).astype(np.float32)
    cx = (pose_xy_px[0] + rng.normal(0, sigma_px, n)).astype(np.float32)
    cy = (pose_xy_px[1] + rng.normal(0, sigma_px, n)).astype(np.float32)
    st["dx_m"], st["dy_m"] = dx, dy
    st["init_x_px"] = cx - dx * np.float32(scale)
    st["init_y_px"] = cy - dy * np.float32(scale)
    st["theta"] = (heading + np.deg2rad(rng.normal(0, sigma_deg, n))).astype(np.float32)
    st["scale"] = np.float32(scale)
    st["have_init"] = 1
    last_dist = rng.uniform(0, 0.4, n).astype(np.float32)
    return st, last_dist


def particles_global(n: int, class_map: np.ndarray, scale: float = 2.0, seed: int = 1234):
    """Uniform on road pixels, heading unknown (have_init = false -> 40-shift search)."""
    rng = np.random.default_rng(seed + 13)
    h, w = class_map.shape
    road = road_pixels(class_map)
    pick = road[rng.integers(0, road.size, n)]
    st = np.zeros(n, dtype=STATE_DTYPE)
    st["init_x_px"] = (pick % w).astype(np.float32) + rng.random(n, dtype=np.float32)
    st["init_y_px"] = (pick // w).astype(np.float32) + rng.random(n, dtype=np.float32)
    st["scale"] = np.float32(scale)
    st["have_init"] = 0
    last_dist = rng.uniform(0, 0.4, n).astype(np.float32)
    return st, last_dist


def grid_centers(h: int, w: int, stride: int) -> np.ndarray:
    """Dense lattice of hypothesis centres (cfg4), (n, 2) float32 as (x, y)."""
    xs = np.arange(stride // 2, w, stride, dtype=np.float32)
    ys = np.arange(stride // 2, h, stride, dtype=np.float32)
    gx, gy = np.meshgrid(xs, ys, indexing="xy")
    return np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1).astype(np.float32)


def default_pose(class_map: np.ndarray, seed: int = 1234):
    """A ground-truth pose on a road pixel near the map centre."""
    h, w = class_map.shape
    road = road_pixels(class_map)
    ry, rx = road // w, road % w
    d = (ry - h / 2) ** 2 + (rx - w / 2) ** 2
    k = int(np.argmin(d))
    return (float(rx[k]) + 0.25, float(ry[k]) + 0.25), 0.6
