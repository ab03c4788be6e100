"""The ADAPTERS as compiled code: top_down_renderer_b200/adapters/*.cpp are the bodies that replace the reference's .cpp
files — written against the reference's UNCHANGED class declarations (read from /root/reference/include at build time,
with oracle/ref_shim standing in for ROS / Eigen / PCL) and calling the C ABI.  `make -C oracle _adapters` links them with
the CPU stand-in of the C ABI (here) and with libtdr_b200.so (on the GPU box), behind the same C interface as the
reference build, so the reference's own bodies and the adapters can be run side by side."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import refbuild as ref
from top_down_renderer_b200 import synth

ANG = np.float32(2 * math.pi / 100)


def _scan():
    cm = synth.make_class_map(260, 300, 4, seed=21)
    pose, heading = synth.default_pose(cm, seed=21)
    pts = synth.make_scan(cm, pose, heading, seed=21, n_rings=32, n_az=256)
    pts[::19, :2] = 0
    return pts, synth.identity_lut(4)


def _check(kind):
    pts, lut = _scan()
    for res in (4.0, 1.5, 0.5):
        got = ref.adapter_render_polar(kind, pts, res, ANG, 100, 25, lut, 4)
        assert np.array_equal(got, orc.render_polar(pts, res, ANG, 100, 25, lut, 4)) and got.sum() > 1000
        if ref.available():
            assert np.array_equal(got, ref.render_polar(pts, res, ANG, 100, 25, lut, 4))          # the reference's own body
        cart = ref.adapter_render_cart(kind, pts, res, 48, 64, lut, 4)
        assert np.array_equal(cart, orc.render_cart(pts, res, 48, 64, lut, 4).reshape(cart.shape)) and cart.sum() > 100
        if ref.available():
            assert np.array_equal(cart, ref.render_cart(pts, res, 48, 64, lut, 4))
    # other raster shapes: the size of the caller's images defines the raster
    assert np.array_equal(ref.adapter_render_polar(kind, pts, 2.0, np.float32(2 * math.pi / 36), 36, 9, lut, 4),
                          orc.render_polar(pts, 2.0, np.float32(2 * math.pi / 36), 36, 9, lut, 4))


@pytest.mark.skipif(not ref.adapters_available("cpu"), reason="no /root/reference and no prebuilt adapters")
def test_scan_renderer_adapters_on_the_cpu_standin():
    _check("cpu")


@pytest.mark.gpu
def test_scan_renderer_adapters_on_the_device():
    if not ref.adapters_available("gpu"):
        pytest.skip("no prebuilt oracle/_ref/libtdr_adapters_gpu.so")
    try:
        ref.adapters("gpu")
    except OSError as e:
        pytest.skip(f"adapters do not load here: {e}")
    _check("gpu")
