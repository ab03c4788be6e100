"""The CUDA path against the REFERENCE'S OWN SOURCE on the B200: oracle/_ref/libtdr_ref.so (the reference's seven hot-path
translation units compiled unmodified against stand-in headers, oracle/ref_shim/README.md) is prebuilt in the container
that has /root/reference and travels with the snapshot; here its classes run on the GPU box's host beside libtdr_b200.

Same bars as tests/test_gpu_parity.py.  Weights pass through Eigen reductions, which the reference build's stand-in Eigen
sums sequentially (the oracle restates Eigen's own SSE2 order), so the device is held to 1e-5 PLUS the oracle's own
distance from that build (about 1e-7); everything else is bit for bit."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refbuild as ref
from tests.common import ANG_RES, N_R, N_THETA, make_ctx, make_world, rel_err

pytestmark = pytest.mark.gpu
SEED, N = 41, 600


@pytest.fixture(scope="module")
def rw():
    if not ref.available():
        pytest.skip("no prebuilt oracle/_ref and no /root/reference")
    try:
        ref.lib()
    except (OSError, FileNotFoundError) as e:
        pytest.skip(f"oracle/_ref does not load here: {e}")
    wd = make_world(h=300, w=400, C=4, seed=11, res=2.0)
    rmap = ref.Map.from_class_image(wd.img, wd.lut, wd.C, 1.0, center=(wd.w // 2, wd.h // 2))
    rmap.set_polar_table(wd.tab, N_THETA, N_R)
    ctx = make_ctx(wd)
    yield wd, rmap, ctx
    ctx.close()


def test_distance_fields_equal_the_reference_build(rw):
    wd, rmap, ctx = rw
    layers, mask = ctx.map_get_layers()
    r_layers, r_mask = rmap.get()
    assert layers.shape == r_layers.shape and np.array_equal(mask, r_mask)
    assert np.array_equal(layers.view(np.uint32), r_layers.view(np.uint32))


@pytest.mark.parametrize("res", [2.0, 0.5])
def test_class_images_equal_the_reference_build(rw, res):
    wd, rmap, ctx = rw
    ctx.scan_set_points(wd.pts)
    got = ctx.scan_render_polar(res, ANG_RES, N_THETA, N_R)
    want = ref.render_polar(wd.pts, res, ANG_RES, N_THETA, N_R, wd.lut, wd.C)
    assert got.sum() > 1000 and np.array_equal(got, want.reshape(got.shape))


def test_polar_gather_equals_the_reference_build(rw):
    wd, rmap, ctx = rw
    centers = np.float32([[wd.pose[0], wd.pose[1]], [3.2, 7.9], [wd.w - 0.6, wd.h - 0.4], [-60.0, 120.0], [199.5, 150.5]])
    d, m = ctx.map_local_polar(centers, 2.0, 2.0)
    for i in range(len(centers)):
        dr, mr = rmap.local_map_polar(float(centers[i, 0]), float(centers[i, 1]), 2.0, 2.0)
        assert np.array_equal(np.asarray(m[i]).reshape(-1), mr.reshape(-1)), i
        assert np.array_equal(np.asarray(d[i]).reshape(-1).view(np.uint32), dr.reshape(-1).view(np.uint32)), i


def test_filter_step_equals_the_reference_build(rw):
    """ParticleFilter of the reference (initialise -> propagate -> update on its own engine) and the device, stage by stage
    on the reference's own intermediate values"""
    wd, rmap, ctx = rw
    kw = dict(fixed_scale=2.0, init_pos_px=(float(wd.pose[0]), float(wd.pose[1])), init_pos_px_cov=8.0,
              init_pos_deg_theta=math.degrees(wd.heading), init_pos_deg_cov=4.0)
    f = ref.Filter(rmap, N, SEED, regularization=0.7, pos_cov=0.15, theta_cov=0.004, **kw)
    f.propagate(0.4, 0.05, 0.01)
    st1, ld1, _ = f.get()
    assert len(st1) == N
    f.update(wd.scan, wd.res)
    scored, ld_s, raw = f.get(scored_set=True)
    wn = f.weights()
    cur, _, _ = f.get()
    M = len(cur)
    assert len(scored) == N and len(wn) == N and 0 < M <= N
    # a9 / a10: raw weights of the same particle set
    ctx.scan_set_polar_images(wd.scan)
    ctx.pf_set_states(st1, ld1)
    w = ctx.pf_score(wd.res)
    w_o = orc.score_all(st1.copy(), wd.fp, wd.layers, wd.mask, 1.0, wd.tab, N_THETA, N_R, wd.scan, wd.res, wd.thetas, wd.shifts)
    slack = float(rel_err(w_o, raw).max())                 # the oracle's distance from the sequentially summing build
    assert slack <= 1e-6
    assert rel_err(w, w_o).max() <= 1e-5 and rel_err(w, raw).max() <= 1.001e-5 + slack
    # a11: normalisation of the reference's own raw weights
    ctx.pf_set_states(scored, ld_s)
    ctx.pf_set_weights(raw)
    arg, _ = ctx.pf_normalize()
    wn_g = ctx.pf_get_weights(N)
    assert rel_err(wn_g, wn).max() <= 1e-6 and wn[arg] >= wn.max() * (1 - 1e-6)
    # a12: systematic resampling of the reference's own normalised weights with the engine's next uniform
    so, _, _, used = orc.init_particles(SEED, wd.layers, 1.0, (wd.w // 2, wd.h // 2), N, **kw)
    _, _, _, used_p = orc.propagate(so, 0.4, 0.05, 0.01, True, 0.15, 0.004, SEED, discard=used)
    u = orc.uniform_draw(SEED, discard=used + used_p)
    ctx.pf_set_states(scored, ld_s)
    ctx.pf_set_weights(wn)
    idx = ctx.pf_resample(u, M)
    new = ctx.pf_get_states()
    assert len(new) == M
    for k in ("init_x_px", "init_y_px", "dx_m", "dy_m", "theta", "scale", "have_init"):
        assert np.array_equal(new[k], cur[k]) and np.array_equal(scored[k][idx], cur[k]), k
    # a13: pose of the resampled set
    mean, _, _, _ = ctx.pf_pose(want_ml=False)
    r_mean, _, _, _ = f.pose()
    assert abs(mean[0] - r_mean[0]) <= 0.002 and abs(mean[1] - r_mean[1]) <= 0.002
    assert abs(mean[2] - r_mean[2]) <= math.radians(0.01) and mean[3] == r_mean[3]


def test_ref_mini_fixture_through_the_c_abi():
    """tests/golden/ref_mini.npz — arrays computed by the reference's own classes (tests/golden/make_ref_golden.py) — against
    the CUDA path; needs neither the reference nor its build on the GPU box"""
    import os
    from top_down_renderer_b200.core import Context
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mini.npz"))
    Cn, res, ang = int(g["num_classes"]), float(g["res"]), g["ang_res"]
    N = len(g["propagated"])
    thetas, shifts = orc.search_list(N_THETA)
    c = Context(0)
    try:
        c.map_set_class_image(g["img"], g["lut"], Cn, 1.0)
        layers, mask = c.map_get_layers()
        assert np.array_equal(layers.view(np.uint32), g["layers"].view(np.uint32)) and np.array_equal(mask, g["mask"])      # a3, a4
        c.map_set_polar_table(g["tab"], N_THETA, N_R)
        c.scan_set_lut(g["lut"], Cn)
        c.scan_set_points(g["pts"])
        assert np.array_equal(c.scan_render_polar(res, ang, N_THETA, N_R), g["scan"])                                       # a1
        assert np.array_equal(c.scan_render_cart(res, 40, 56), g["cart"])                                                   # a2
        d, m = c.map_local_polar(g["centres"], 2.0, res)                                                                    # a7
        for i in range(len(g["centres"])):
            assert np.array_equal(np.asarray(m[i]).reshape(-1), g["local_m"][i].reshape(-1)), i
            assert np.array_equal(np.asarray(d[i]).reshape(-1).view(np.uint32), g["local_d"][i].reshape(-1).view(np.uint32)), i
        c.pf_set_params(Cn, regularization=0.7)
        c.pf_set_search(thetas, shifts)
        c.scan_set_polar_images(g["scan"])
        c.pf_set_states(g["propagated"], g["last_dist"])
        w = c.pf_score(res)                                                                                                 # a9, a10
        assert rel_err(w, g["raw"]).max() <= 1.1e-5          # 1e-5 to Eigen's order + the reference build's sequential sums (1e-6)
        c.pf_set_states(g["scored"], g["last_dist"])
        c.pf_set_weights(g["raw"])
        c.pf_normalize()                                                                                                    # a11
        assert rel_err(c.pf_get_weights(N), g["weights_norm"]).max() <= 1e-6
        # a12: the uniform is the engine output the reference consumed next: generate_canonical<float, 24> of it
        u = np.float32(int(g["engine_peek"][1])) / np.float32(4294967296.0)
        assert u < 1
        M = len(g["resampled"])
        c.pf_set_states(g["scored"], g["last_dist"])
        c.pf_set_weights(g["weights_norm"])
        c.pf_resample(float(u), M)
        new = c.pf_get_states()
        assert len(new) == M
        for k in ("init_x_px", "init_y_px", "dx_m", "dy_m", "theta", "scale", "have_init"):
            assert np.array_equal(new[k], g["resampled"][k]), k
        mean, _, _, _ = c.pf_pose(want_ml=False)                                                                            # a13
        assert abs(mean[0] - g["mean"][0]) <= 0.002 and abs(mean[1] - g["mean"][1]) <= 0.002
        assert abs(mean[2] - g["mean"][2]) <= math.radians(0.01)
        # the search set: force_on_map / NaN / all-NaN search (weight 1 / (FLT_MAX + reg), a denormal)
        c.pf_set_params(Cn, regularization=0.7, force_on_map=True)
        c.pf_set_states(g["search_in"], g["search_last_dist"])
        sw = c.pf_score(res)
        r = g["search_raw"]
        assert np.array_equal(np.isnan(sw), np.isnan(r)) and np.array_equal(sw == 0, r == 0)
        assert rel_err(sw, r).max() <= 1.1e-5
        got = c.pf_get_states()
        assert np.array_equal(got["have_init"], g["search_scored"]["have_init"])
        assert (got["theta"] == g["search_scored"]["theta"]).mean() >= 0.95       # near-ties between candidates may flip with the summation order
    finally:
        c.close()
