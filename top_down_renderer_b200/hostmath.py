"""Host-side inputs of the device library that the reference computes on the CPU once
(SURVEY.md decision 3): the polar offset table and the theta-search candidate list.

The C++ adapters compute these with the reference's own expressions (Eigen cos/sin);
this module is the numpy equivalent used by bench.py and the Python host mirror.  Nothing here touches
the oracle.
"""
from __future__ import annotations

import math

import numpy as np


def polar_table(n_theta: int, n_r: int, ang_res, resolution: float = 1.0) -> np.ndarray:
    """ang_sample_pts_: TopDownMap::samplePts + TopDownMapPolar::samplePtsPolar
    (top_down_map.cpp:367-389, top_down_map_polar.cpp:7-19).  Returns (P, 2): [:, 0] row (y) offset,
    [:, 1] col (x) offset, p = r*n_theta + theta."""
    P = n_theta * n_r
    p = np.arange(P)
    row = (p % n_theta).astype(np.float32)
    col = (p // n_theta).astype(np.float32)
    ang = (row - np.float32((n_theta - 1) / 2.0)) * np.float32(ang_res)
    rho = col * np.float32(1.0 / resolution)
    tab = np.empty((P, 2), dtype=np.float32)
    tab[:, 0] = np.cos(ang).astype(np.float32) * rho
    tab[:, 1] = np.sin(ang).astype(np.float32) * rho
    return tab


def rot_to_shift(rot, n_theta: int) -> int:
    """state_particle.cpp:123-128: round(rot*n_theta/2/pi) wrapped into [0, n_theta)"""
    v = float(np.float32(np.float32(np.float32(rot) * np.float32(n_theta)) / np.float32(2.0))) / math.pi
    r = math.floor(abs(v) + 0.5) * (1.0 if v >= 0 else -1.0)   # std::round: half away from zero
    s = int(math.fmod(r, n_theta))
    return s + n_theta if s < 0 else s


def search_list(n_theta: int = 100):
    """the float loop `for (float t = 0; t < 2*M_PI; t += 2*M_PI/40)` of StateParticle::computeWeight
    (state_particle.cpp:197): t accumulates as (float)((double)t + 2pi/40).  Returns (thetas, shifts)."""
    thetas, shifts = [], []
    t = np.float32(0.0)
    while float(t) < 2 * math.pi:
        thetas.append(t)
        shifts.append(rot_to_shift(t, n_theta))
        t = np.float32(float(t) + 2 * math.pi / 40)
    return np.array(thetas, dtype=np.float32), np.array(shifts, dtype=np.int32)
