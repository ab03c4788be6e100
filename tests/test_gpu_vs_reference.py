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
