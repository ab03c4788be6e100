"""A second, deliberately naive restatement of the hot path in numpy / pure Python (TEST INFRASTRUCTURE ONLY).

It exists to cross-check oracle/tdr_oracle.cpp on small inputs with *different code*: explicit np.roll for
the shift-correlation, brute-force nearest-seed search for the distance transform, a literal O(N*M)
resampler.  Citations are file:line in the reference checkout.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

_libm = C.CDLL("libm.so.6")
_libm.atan2f.restype = C.c_float
_libm.atan2f.argtypes = [C.c_float, C.c_float]
_libm.roundf.restype = C.c_float
_libm.roundf.argtypes = [C.c_float]
f32 = np.float32


def render_polar(pts, res, ang_res, n_theta, n_r, lut, num_classes):
    """scan_renderer_polar.cpp:83-109; returns (C, n_r, n_theta)"""
    img = np.zeros((num_classes, n_r, n_theta), dtype=np.float32)
    for p in pts:
        x, y, inten = f32(p[0]), f32(p[1]), p[4]
        if x == 0 and y == 0:
            continue
        theta = f32(_libm.atan2f(x, y))
        r = np.sqrt(f32(f32(x * x) + f32(y * y)), dtype=np.float32)
        ti = int(f32(_libm.roundf(f32(theta / f32(ang_res))) + f32(n_theta // 2)))
        ri = int(_libm.roundf(f32(r / f32(res))))
        if 0 <= ti < n_theta and 0 <= ri < n_r:
            cls = lut[int(inten)]
            if cls >= 0:
                img[cls, ri, ti] += 1
    return img


def edt_layers_bruteforce(bin_layers, resolution):
    """top_down_map.cpp:289-326 with a brute-force nearest-zero search; layers (C, cols, rows) of 0/1"""
    Cn, cols, rows = bin_layers.shape
    mask = (bin_layers.astype(np.uint8).sum(axis=0) > Cn - 1).astype(np.uint8)
    out = np.empty_like(bin_layers, dtype=np.float32)
    xs, ys = np.meshgrid(np.arange(cols), np.arange(rows), indexing="ij")
    for c in range(Cn):
        seeds = np.argwhere(bin_layers[c] == 0)
        if len(seeds) == 0:
            d = np.full((cols, rows), 50.0, dtype=np.float32)
        else:
            d2 = np.min((xs[..., None] - seeds[:, 0]) ** 2 + (ys[..., None] - seeds[:, 1]) ** 2, axis=-1)
            d = np.minimum(np.sqrt(d2.astype(np.float32), dtype=np.float32) * f32(resolution), f32(50.0))
        d[mask != 0] = 0
        out[c] = d
    return out, mask


def local_map_polar(layers, mask, resolution, tab, cx, cy, scale, res):
    """top_down_map_polar.cpp:21-53; layers (C, cols, rows); returns (C, P), (P,)"""
    Cn, cols, rows = layers.shape
    P = tab.shape[0]
    d = np.zeros((Cn, P), dtype=np.float32)
    m = np.ones(P, dtype=np.uint8)
    oy, ox = f32(f32(cy) / f32(resolution)), f32(f32(cx) / f32(resolution))
    for p in range(P):
        r = int(_libm.roundf(f32(f32(f32(tab[p, 0] * f32(scale)) * f32(res)) + oy)))
        c = int(_libm.roundf(f32(f32(f32(tab[p, 1] * f32(scale)) * f32(res)) + ox)))
        if 0 <= r < rows and 0 <= c < cols:
            d[:, p] = layers[:, c, r]
            m[p] = mask[c, r]
    return d, m


def cost_for_shift(scan, classes, known, n_theta, n_r, class_weights, shift):
    """state_particle.cpp:112-155 with an explicit roll: scan row r pairs with map row (r - s) mod n_theta.
    float64 accumulation (the reference's fp32 SIMD order is not part of the contract; 1e-5 is)."""
    known = known.reshape(n_r, n_theta).astype(np.float64)
    if known.sum() / known.size < 0.5:
        return float("nan")
    cost = norm = 0.0
    for c in range(scan.shape[0]):
        sc = scan[c].reshape(n_r, n_theta).astype(np.float64)
        mp = classes[c].reshape(n_r, n_theta).astype(np.float64)
        cost += (sc * np.roll(mp, shift, axis=1)).sum() * 0.01 * float(class_weights[c])
        norm += (sc * np.roll(known, shift, axis=1)).sum()
    return cost / norm if norm != 0 else float("nan")


def normalize(w, last_dist):
    """particle_filter.cpp:107-147 (loop counters read as i = 0), scalar Python, sequential fp32"""
    w = np.array(w, dtype=np.float32)
    n = len(w)
    s, nv = f32(0), 0
    for v in w:
        if v == v:
            s = f32(s + v); nv += 1
    with np.errstate(all="ignore"):
        mean = f32(s / f32(nv)) if nv else f32(np.nan)
    bs, nu = f32(0), 0
    for v in w:
        if v == v and v < mean:
            bs = f32(float(bs) + float(f32(v - mean)) ** 2); nu += 1
    with np.errstate(all="ignore"):
        bs = np.sqrt(f32(bs / f32(nu)), dtype=np.float32) if nu else f32(np.nan)
    if s == 0 or nu < 1:
        w[:] = 1
    else:
        w[np.isnan(w)] = f32(mean - bs)
    w = (w / _eigen_sum(w)).astype(np.float32)
    d = np.minimum(last_dist.astype(np.float32) * f32(5), f32(1))
    w = (d * w).astype(np.float32) + ((f32(1) - d) / f32(n)).astype(np.float32)
    w = (w / _eigen_sum(w)).astype(np.float32)
    return w, int(np.argmax(w))


def _eigen_sum(x):
    """Eigen linear-vectorised float sum, SSE2 packets of 4, 2x unrolled (Eigen/src/Core/Redux.h)"""
    x = np.asarray(x, dtype=np.float32)
    n = len(x)
    a4, a8 = (n // 4) * 4, (n // 8) * 8
    if a4 == 0:
        r = x[0]
        for v in x[1:]:
            r = f32(r + v)
        return r
    p0 = x[0:4].copy()
    if a4 > 4:
        p1 = x[4:8].copy()
        for i in range(8, a8, 8):
            p0 = (p0 + x[i:i + 4]).astype(np.float32)
            p1 = (p1 + x[i + 4:i + 8]).astype(np.float32)
        p0 = (p0 + p1).astype(np.float32)
        if a4 > a8:
            p0 = (p0 + x[a8:a8 + 4]).astype(np.float32)
    r = f32(f32(p0[0] + p0[2]) + f32(p0[1] + p0[3]))
    for v in x[a4:]:
        r = f32(r + v)
    return r


def resample_literal(w, u, M):
    """particle_filter.cpp:172-185, the O(N*M) double loop"""
    n = len(w)
    idx = np.empty(M, dtype=np.int32)
    for i in range(M):
        sample = f32(f32(f32(i) + f32(u)) / f32(M))
        run = f32(0)
        j = 0
        while True:
            run = f32(run + w[j])
            if run > sample or j == n - 1:
                break
            j += 1
        idx[i] = j
    return idx


def propagate_with_z(states, tx, ty, omega, scale_freeze, pos_cov, theta_cov, z):
    """StateParticle::propagate (state_particle.cpp:57-78) with the noise given as standard normal variates
    z[n, 4] = (theta, dx, dy, scale): every draw is `z * stddev + mean` in fp32, as libstdc++ forms it.  Vectorised
    float32 numpy; cos / sin through float64 and rounded (glibc's cosf / sinf agree up to rare 1-ulp cases)."""
    f = np.float32
    st = states.copy()
    z = np.asarray(z, dtype=f).reshape(-1, 4)
    th = st["theta"].astype(f)
    c, s = np.cos(th.astype(np.float64)).astype(f), np.sin(th.astype(np.float64)).astype(f)
    gx = c * f(tx) - s * f(ty)
    gy = s * f(tx) + c * f(ty)
    lx, ly = st["dx_m"].astype(f), st["dy_m"].astype(f)
    dx, dy = lx + gx, ly + gy
    dist = np.sqrt(gx * gx + gy * gy)
    sd_pos, sd_th = f(pos_cov) * dist, f(theta_cov) * dist
    st["theta"] = th + ((z[:, 0] * sd_th + f(0)) + f(omega))
    dx = dx + (z[:, 1] * sd_pos + f(0))
    dy = dy + (z[:, 2] * sd_pos + f(0))
    st["dx_m"], st["dy_m"] = dx, dy
    if not scale_freeze:
        with np.errstate(divide="ignore"):
            sd_sc = np.minimum(2.0 / dist.astype(np.float64), 0.02).astype(f)
        st["scale"] = st["scale"].astype(f) * (z[:, 3] * sd_sc + f(1))
    mx, my = lx - dx, ly - dy
    return st, np.sqrt(mx * mx + my * my)


def raster_polygons(polys, poly_class, map_w, map_h, resolution, num_classes, exclusive):
    """getRasterMap + getClasses (top_down_map.cpp:328-408) for rot = 0, vectorised numpy in float32: one (rows, cols)
    grid of sample points, the even-odd rule as a crossing COUNT per polygon (parity instead of the reference's sign
    flips), classes combined with logical or.  Returns (C, cols, rows) like the oracle."""
    f = np.float32
    rows, cols = int(map_h / resolution), int(map_w / resolution)

    def lin(size):
        lo, hi = f(-resolution * (size - 1) / 2.0), f(resolution * (size - 1) / 2.0)
        if size == 1:
            return np.array([lo], dtype=f)
        step = (hi - lo) / f(size - 1)
        v = lo + np.arange(size, dtype=f) * step
        v[-1] = hi
        return v.astype(f)
    py = (lin(rows) + f(map_h) / f(2))[:, None] + np.zeros((1, cols), dtype=f)       # (rows, cols)
    px = (lin(cols) + f(map_w) / f(2))[None, :] + np.zeros((rows, 1), dtype=f)
    inside = np.zeros((num_classes, rows, cols), dtype=bool)
    for poly, c in zip(polys, poly_class):
        p = np.asarray(poly, dtype=f)
        crossings = np.zeros((rows, cols), dtype=np.int32)
        for i in range(len(p)):
            xi, yi = p[i]
            xj, yj = p[i - 1]
            straddle = (py < yi) != (py < yj)
            with np.errstate(divide="ignore", invalid="ignore"):
                xint = xi + ((xj - xi) * (py - yi) / (yj - yi))
            crossings += (straddle & (px < xint)).astype(np.int32)
        inside[c] |= (crossings % 2) == 1
    lay = np.where(inside, f(0), f(1)).astype(f)
    for under in exclusive:
        for cls in exclusive:
            if under < cls:
                lay[under] = lay[under] + (f(1) - lay[cls])
        lay[under] = np.minimum(lay[under], f(1))
    return np.ascontiguousarray(lay.transpose(0, 2, 1))


def active_best_rel_pos(layers, mask, resolution, tab, n_theta, n_r, preds):
    """ActiveLocalizer::getBestRelPos (active_localizer.cpp:45-82) with np.roll for the heading rotation and float64 sums
    of the pairwise absolute differences.  Returns ((dist, theta), best_diff)."""
    f = np.float32
    preds = np.asarray(preds, dtype=f).reshape(-1, 3)
    C_ = layers.shape[0]
    best_diff, best = 0.0, (0.0, 0.0)
    dist = f(50)
    while best_diff < 6000 and dist < 150:
        theta = f(0)
        while float(theta) < 2 * math.pi:
            maps = []
            for x, y, th in preds:
                px = f(x + dist * f(math.cos(f(theta + th))))      # cosf stand-in: double cos rounded to float
                py = f(y + dist * f(math.sin(f(theta + th))))
                d, _ = local_map_polar(layers, mask, resolution, tab, px, py, f(1), f(2))
                v = float(f(f(th * f(n_theta)) / f(2))) / math.pi
                shift = int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)) % n_theta
                maps.append(np.roll(d.reshape(C_, n_r, n_theta), shift, axis=2).astype(np.float64))
            total, cnt = 0.0, 0
            for i in range(len(maps)):
                for j in range(i):
                    total += np.abs(maps[i] - maps[j]).sum()
                    cnt += C_
            diff = total / cnt if cnt else float("nan")
            if diff > best_diff:
                best_diff, best = diff, (float(dist), float(theta))
            theta = f(float(theta) + math.pi / 8)
        dist = f(dist + 25)
    return best, best_diff


# ---- particle initialisation, restated WITHOUT libstdc++: the Mersenne twister from numpy, the two
# distributions from their published libstdc++ algorithms (bits/random.tcc), logf / pow from glibc through ctypes
class LibstdcxxEngine:
    """std::mt19937(seed) plus generate_canonical<float, 24>: one 32-bit output -> float(x) / 2^32, and 1.0 (only
    reachable through the rounding of x >= 2^32 - 128) replaced by the largest float below 1"""

    def __init__(self, seed):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed))            # init_genrand: the same state std::mt19937(seed) starts from
        self.draws = 0

    def canonical(self):
        self.draws += 1
        r = np.float32(int(self.bg.random_raw())) / np.float32(4294967296.0)
        return np.float32(r) if r < 1 else np.nextafter(np.float32(1), np.float32(0))


class LibstdcxxNormal:
    """std::normal_distribution<float>(mean, stddev): Marsaglia's polar method, second value of a pair saved"""

    def __init__(self, mean=0.0, stddev=1.0):
        import ctypes
        self.mean, self.stddev, self.saved = np.float32(mean), np.float32(stddev), None
        self.logf = ctypes.CDLL("libm.so.6").logf
        self.logf.restype, self.logf.argtypes = ctypes.c_float, [ctypes.c_float]

    def __call__(self, eng):
        f = np.float32
        if self.saved is not None:
            ret, self.saved = self.saved, None
        else:
            while True:
                x = f(f(2) * eng.canonical() - 1.0)
                y = f(f(2) * eng.canonical() - 1.0)
                r2 = f(f(x * x) + f(y * y))
                if not (r2 > 1.0 or r2 == 0.0):
                    break
            mult = np.sqrt(f(f(f(-2) * f(self.logf(float(r2)))) / r2))
            self.saved = f(x * mult)
            ret = f(y * mult)
        return f(f(ret * self.stddev) + self.mean)


def init_particles(seed, layers, resolution, map_center, max_n, init_pos_px=(-1.0, -1.0), init_pos_px_cov=-1.0,
                   init_pos_m=(math.inf, math.inf), init_pos_deg_theta=math.inf, init_pos_deg_cov=10.0, fixed_scale=-1.0):
    """ParticleFilter::initializeParticles (particle_filter.cpp:19-84) with StateParticle's constructor
    (state_particle.cpp:3-49) on distance layers (C, cols, rows).  Returns (list of (init_x, init_y, theta, scale,
    have_init), engine outputs consumed)."""
    f = np.float32
    C, cols, rows = layers.shape
    eng = LibstdcxxEngine(seed)
    px, py, pcov = f(init_pos_px[0]), f(init_pos_px[1]), f(init_pos_px_cov)

    def on_road(x, y):                                 # getClassesAtPoint(Vector2i(x, y)) contains class 1
        cx, cy = int(f(int(x)) / f(resolution)), int(f(int(y)) / f(resolution))
        return 0 <= cx < cols and 0 <= cy < rows and layers[1, cx, cy] < 1

    def construct():
        normal = LibstdcxxNormal()
        mw, mh = f(f(cols) * f(resolution)), f(f(rows) * f(resolution))
        scale = f(math.pow(10, (float(eng.canonical()) - 0.5) * 2)) if fixed_scale < 0 else f(fixed_scale)
        while True:
            if px > 0:
                x = min(max(f(f(normal(eng) * pcov) + px), f(0)), mw)
                y = min(max(f(f(normal(eng) * pcov) + py), f(0)), mh)
            else:
                x = f(eng.canonical() * mw)
                y = f(eng.canonical() * mh)
            if on_road(x, y):
                break
        if init_pos_deg_theta != math.inf:
            th = f(f(normal(eng) * f(init_pos_deg_cov)) + f(init_pos_deg_theta))
            return (x, y, f(float(th) * (math.pi / 180)), scale, 1)
        return (x, y, f(0), scale, 0)

    num_at_scale = 10 if fixed_scale < 0 else 1
    if fixed_scale >= 0 and init_pos_m[0] != math.inf:
        px = f(f(f(init_pos_m[0]) * f(fixed_scale)) + f(map_center[0]))
        py = f(f(f(init_pos_m[1]) * f(fixed_scale)) + f(map_center[1]))
        if px < 0 or px >= cols or py < 0 or py >= rows:
            return [], 0
        if not any(on_road(f(px + f(dx)), f(py + f(dy))) for dx in range(-4, 5) for dy in range(-4, 5)):
            return [], 0
    out = []
    for _ in range(max_n // num_at_scale):
        proto = construct()
        scale = f(0)
        while scale < 1:
            part = construct()
            if fixed_scale < 0:
                part = proto[:3] + (f(math.pow(10.0, float(scale))), proto[4])
            out.append(part)
            construct()
            scale = f(float(scale) + 1.0 / num_at_scale)
    return out, eng.draws
