"""Timings of the rows built beyond the first path: propagate on 1e6 device-resident particles and the
vector-map path on a 4000 x 4000 px, 6-class map.  CUDA events on the context stream, best of 5 after a warm-up;
CPU figures: the oracle (1 thread) on a bounded sample.   python tools/time_widen.py > profiles/r01_widen_timings.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from top_down_renderer_b200 import synth  # noqa: E402
from top_down_renderer_b200.core import Context  # noqa: E402


def timed(stream, fn, reps=5):
    best = 1e9
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    torch.cuda.set_device(0)
    c = Context(0)
    stream = torch.cuda.ExternalStream(c.stream, device=0)
    # ---- propagate
    n = 1_000_000
    st, ld = synth.particles_tracking(n, (2000.0, 2000.0), 0.6)
    c.pf_set_states(st, ld)
    z = np.random.default_rng(1).standard_normal((n, 4)).astype(np.float32)
    zt = torch.from_numpy(z).pin_memory()
    t_rng = timed(stream, lambda: c.pf_propagate_rng((0.7, -0.2), 0.03, False, 0.3, 0.1, 7, 1))
    t0 = time.perf_counter()
    for _ in range(5):
        c.pf_propagate((0.7, -0.2), 0.03, False, 0.3, 0.1, z)          # host variates: 16 MB H2D + kernel, synchronous
    t_inj = (time.perf_counter() - t0) / 5 * 1e3
    ns = 100_000
    t0 = time.perf_counter()
    orc.propagate(st[:ns], 0.7, -0.2, 0.03, False, 0.3, 0.1, 1)
    t_cpu = (time.perf_counter() - t0) * 1e3 * n / ns
    print(f"propagate, {n} particles: device RNG {t_rng:.3f} ms ({n * 36 / t_rng / 1e6:.0f} GB/s of 36 B/particle state traffic); "
          f"injected variates from host {t_inj:.2f} ms (16 MB H2D); oracle (reference loop, 1 thread, shared mt19937) {t_cpu:.0f} ms")
    # ---- vector map
    rng = np.random.default_rng(3)
    W = H = 4000
    polys, cls = [], []
    for _ in range(3000):
        k = int(rng.integers(4, 16))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(10, 120) * rng.uniform(0.4, 1.0, k)
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        polys.append(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1).astype(np.float32))
        cls.append(int(rng.integers(0, 6)))
    n_edges = sum(len(p) for p in polys)
    excl = [0] * 6 + [0, 1, 2, 3]
    c.map_set_polygons(polys, cls, W, H, 0.0, 6, 1.0, excl, want_layers=False)
    t0 = time.perf_counter()
    for _ in range(3):
        c.map_set_polygons(polys, cls, W, H, 0.0, 6, 1.0, excl, want_layers=False)
    t_map = (time.perf_counter() - t0) / 3 * 1e3
    sub = 400                                                           # oracle: O(pixels x edges) on a 400 x 400 corner
    sp = [p for p in polys if p[:, 0].max() < sub + 120 and p[:, 1].max() < sub + 120]
    sc = [k for p, k in zip(polys, cls) if p[:, 0].max() < sub + 120 and p[:, 1].max() < sub + 120]
    t0 = time.perf_counter()
    orc.raster_polygons(sp, sc, sub, sub, 0.0, 1.0, 6, excl)
    t_cpu = (time.perf_counter() - t0) * 1e3
    print(f"vector map, {W}x{H} px, 6 classes, {len(polys)} polygons / {n_edges} edges: rasterise + seeds + distance fields "
          f"{t_map:.2f} ms (host wall clock incl. H2D of the polygons); oracle rasterisation alone of a {sub}x{sub} corner with its "
          f"{len(sp)} polygons: {t_cpu:.0f} ms (1 thread)")
    c.close()


if __name__ == "__main__":
    main()
