"""Shared fixtures: the cfg1-style synthetic world evaluated by the oracle (the checker)."""
import functools
import math

import numpy as np

from oracle import oracle as orc
from top_down_renderer_b200 import synth

ANG_RES = np.float32(2 * math.pi / 100)
N_THETA, N_R = 100, 25


class World:
    pass


@functools.lru_cache(maxsize=8)
def make_world(h=1000, w=1000, C=4, resolution=1.0, seed=1234, res=4.0):
    wd = World()
    wd.C, wd.h, wd.w, wd.resolution, wd.res = C, h, w, resolution, res
    wd.class_map = synth.make_class_map(h, w, C, seed=seed)
    wd.img = synth.to_cv_image(wd.class_map)
    wd.lut = synth.identity_lut(C)
    wd.bin_layers = orc.class_image_to_layers(wd.img, wd.lut, C, resolution)
    wd.layers, wd.mask = orc.compute_dists(wd.bin_layers, resolution)
    wd.pose, wd.heading = synth.default_pose(wd.class_map, seed=seed)
    wd.pts = synth.make_scan(wd.class_map, wd.pose, wd.heading, seed=seed)
    wd.scan = orc.render_polar(wd.pts, res, ANG_RES, N_THETA, N_R, wd.lut, C)
    wd.tab = orc.polar_table(N_THETA, N_R, ANG_RES, resolution)
    wd.thetas, wd.shifts = orc.search_list(N_THETA)
    wd.cols, wd.rows = wd.layers.shape[1], wd.layers.shape[2]
    wd.fp = orc.make_params(C, regularization=0.7, map_width=wd.cols * resolution, map_height=wd.rows * resolution)
    return wd


def make_ctx(wd, device=0):
    from top_down_renderer_b200.core import Context
    ctx = Context(device)
    ctx.map_set_class_image(wd.img, wd.lut, wd.C, wd.resolution)
    ctx.map_set_polar_table(wd.tab, N_THETA, N_R)
    ctx.scan_set_lut(wd.lut, wd.C)
    ctx.pf_set_params(wd.C, regularization=0.7)
    ctx.pf_set_search(wd.thetas, wd.shifts)
    return ctx


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    d[both_nan] = 0
    d[np.isnan(a) ^ np.isnan(b)] = np.inf
    d[(a == b)] = 0
    return d


def assert_heading_flips_are_ties(wd, st_before, theta_dev, theta_oracle, res, tol=1e-5, scan=None):
    """Where the device chose another candidate heading than the oracle, the two candidates' costs (the oracle's own, for
    that particle) must agree to `tol` relative: a flip is only legitimate as a tie broken by rounding
    (state_particle.cpp:193-204 takes the first strict minimum).  Returns the number of flips."""
    diff = np.flatnonzero(np.asarray(theta_dev) != np.asarray(theta_oracle))
    if diff.size == 0:
        return 0
    src = st_before[diff]
    scales = np.unique(src["scale"])
    assert scales.size == 1, "helper written for one scale"
    cen = np.stack([(src["dx_m"] * src["scale"]).astype(np.float32) + src["init_x_px"],
                    (src["dy_m"] * src["scale"]).astype(np.float32) + src["init_y_px"]], axis=1).astype(np.float32)
    costs = orc.cost_grid(cen, float(scales[0]), wd.fp, wd.layers, wd.mask, wd.resolution, wd.tab, N_THETA, N_R,
                          wd.scan if scan is None else scan, res, wd.shifts).astype(np.float64)
    th = np.asarray(wd.thetas, dtype=np.float32)
    kd = np.array([int(np.flatnonzero(th == t)[0]) for t in np.asarray(theta_dev)[diff]])
    ko = np.array([int(np.flatnonzero(th == t)[0]) for t in np.asarray(theta_oracle)[diff]])
    cd, co = costs[np.arange(diff.size), kd], costs[np.arange(diff.size), ko]
    gap = np.abs(cd - co) / np.maximum(np.abs(co), 1e-300)
    gap[np.isnan(cd) & np.isnan(co)] = 0
    assert np.all(gap <= tol), (diff[np.argmax(gap)], float(np.nanmax(gap)))
    return int(diff.size)
