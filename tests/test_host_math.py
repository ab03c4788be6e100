"""Host build of the kernels' shared index math (csrc/tdr_math.cuh) against glibc and plain sequential loops.

The same inline functions are compiled into the CUDA kernels (with _rn intrinsics instead of
-ffp-contract=off), so these CPU tests pin the arithmetic the GPU parity tests then confirm."""
import ctypes as C
import math

import numpy as np
import pytest

import top_down_renderer_b200 as tdr

fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def hm():
    lib = C.CDLL(tdr.build_hostmath())
    lib.hm_atan2f.restype = C.c_float
    lib.hm_atan2f.argtypes = [C.c_float, C.c_float]
    lib.hm_atan2f_mismatches.restype = C.c_long
    lib.hm_round.restype = C.c_float
    lib.hm_round.argtypes = [C.c_float]
    lib.hm_f2i.argtypes = [C.c_float]
    lib.hm_dist_value.restype = C.c_float
    lib.hm_dist_value.argtypes = [C.c_uint, C.c_float]
    lib.hm_rot_to_shift.argtypes = [C.c_float, C.c_int]
    lib.hm_lattice_index.argtypes = [C.c_float] * 4
    lib.hm_polar_bin.argtypes = [C.c_float] * 4 + [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


def test_atan2f_is_glibc_bit_for_bit(hm):
    rng = np.random.default_rng(1)
    n = 2_000_000
    for scale in (1.0, 100.0, 1e-3, 1e6):
        y = (rng.standard_normal(n) * scale).astype(np.float32)
        x = (rng.standard_normal(n) * scale).astype(np.float32)
        assert hm.hm_atan2f_mismatches(y.ctypes.data_as(fp), x.ctypes.data_as(fp), C.c_long(n)) == 0
    # ratios that land on the argument-reduction thresholds, axes, zeros, infinities, NaN, denormals
    sp = np.array([0.0, -0.0, 1.0, -1.0, 0.4375, 0.6875, 1.1875, 2.4375, 1e-38, 1e-45, 3e38, np.inf, -np.inf, np.nan,
                   2.0 ** 25, 2.0 ** -29, 7.0 / 16, 11.0 / 16, 19.0 / 16, 39.0 / 16], dtype=np.float32)
    yy, xx = [a.reshape(-1).copy() for a in np.meshgrid(sp, sp)]
    assert hm.hm_atan2f_mismatches(yy.ctypes.data_as(fp), xx.ctypes.data_as(fp), C.c_long(len(yy))) == 0


def test_round_and_float_to_int(hm):
    libm = C.CDLL("libm.so.6")
    libm.roundf.restype = C.c_float
    libm.roundf.argtypes = [C.c_float]
    rng = np.random.default_rng(0)
    vals = [0.5, -0.5, 1.5, 2.5, -2.5, 0.49999997, -0.49999997, 8388607.5, 8388608.0, 1e10, -1e10, 3.0, -0.0]
    vals += list((rng.standard_normal(20000) * 100).astype(np.float32)) + list(rng.integers(-500, 500, 2000) + 0.5)
    for v in vals:
        assert hm.hm_round(float(v)) == libm.roundf(float(v)), v        # std::round: half away from zero
    assert hm.hm_f2i(float("nan")) == -2 ** 31 and hm.hm_f2i(3e9) == -2 ** 31 and hm.hm_f2i(-3e9) == -2 ** 31
    assert hm.hm_f2i(-7.9) == -7 and hm.hm_f2i(7.9) == 7


def test_polar_bin_matches_independent_numpy(hm):
    """scan_renderer_polar.cpp:95-102 restated with numpy float32 ops + glibc atan2f via ctypes libm"""
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(2)
    ang = np.float32(2 * math.pi / 100)
    ti, ri = C.c_int(), C.c_int()
    for res in (4.0, 0.5, 1.7):
        pts = (rng.standard_normal((4000, 2)) * 40).astype(np.float32)
        for x, y in pts:
            th = np.float32(libm.atan2f(x, y))
            r = np.sqrt(np.float32(np.float32(x * x) + np.float32(y * y)), dtype=np.float32)
            q = np.float32(th / ang)
            t = int(math.copysign(math.floor(abs(q) + np.float32(0.5)), q)) + 50
            rq = np.float32(r / np.float32(res))
            rr = int(math.floor(rq + np.float32(0.5)))
            ok = 0 <= t < 100 and 0 <= rr < 25
            got = hm.hm_polar_bin(x, y, res, ang, 100, 25, C.byref(ti), C.byref(ri))
            assert bool(got) == ok
            if ok:
                assert (ti.value, ri.value) == (t, rr)
    assert hm.hm_polar_bin(0.0, 0.0, 4.0, ang, 100, 25, C.byref(ti), C.byref(ri)) == 0       # skipped point (:95)
    assert hm.hm_polar_bin(float("nan"), 1.0, 4.0, ang, 100, 25, C.byref(ti), C.byref(ri)) == 0


def test_rot_to_shift_and_search_list(hm):
    from top_down_renderer_b200 import hostmath
    th, sh = hostmath.search_list(100)
    assert len(th) == 40                                  # the loop runs exactly 40 times
    assert list(sh) == [0, 3, 5, 8, 10, 13, 15, 17, 20, 22, 25, 27, 30, 32, 35, 37, 40, 42, 45, 47, 50, 53, 55, 58, 60,
                        63, 65, 68, 70, 73, 75, 78, 80, 83, 85, 88, 90, 93, 95, 98]
    for t, s in zip(th, sh):
        assert hm.hm_rot_to_shift(float(t), 100) == s
    for rot in (-0.1, -3.2, 7.0, 100.0, -100.0, 0.0314, 6.2831):
        assert hm.hm_rot_to_shift(rot, 100) == hostmath.rot_to_shift(rot, 100)
        assert 0 <= hm.hm_rot_to_shift(rot, 100) < 100


def test_dist_value_is_sqrt_mul_min(hm):
    for res in (1.0, 0.5, 2.0, 1.3):
        for d2 in (0, 1, 2, 5, 624, 625, 2499, 2500, 2501, 10000, 65536):
            want = min(np.float32(np.sqrt(np.float32(d2), dtype=np.float32) * np.float32(res)), np.float32(50.0))
            assert hm.hm_dist_value(d2, res) == want


def _seq_prefix(x, skip_nan=False):
    s = np.float32(0)
    out = np.empty(len(x), dtype=np.float32)
    rm = -np.inf
    for i, w in enumerate(x):
        if skip_nan and w != w:
            w = np.float32(0)
        s = np.float32(s + w)
        rm = max(rm, s) if s == s else rm
        out[i] = rm
    return out, s


def _cases(rng, n):
    w = rng.random(n).astype(np.float32)
    yield "uniform_normalised", (w / w.sum()).astype(np.float32)
    yield "raw_weights", (1.0 / (rng.random(n) * 2 + 0.7)).astype(np.float32)
    yield "dyadic_ties", (rng.integers(0, 8, n) * 2.0 ** -20).astype(np.float32)
    sp = np.full(n, 1e-9, dtype=np.float32)
    sp[rng.integers(0, n, max(1, n // 50))] = 1.0
    yield "spiky", (sp / sp.sum()).astype(np.float32)
    yield "denormals", np.where(rng.random(n) < 0.9, 2.93874e-39, 1e-3).astype(np.float32)
    z = rng.random(n).astype(np.float32)
    z[: n // 3] = 0
    yield "leading_zeros", z
    yield "equal_small", np.full(n, 1e-6, dtype=np.float32)


@pytest.mark.parametrize("n", [1, 33, 4096, 50_000])
def test_exact_prefix_equals_sequential_fp32(hm, n):
    """the binade / IncPair algebra of k_exact_seq reproduces s = RN(s + w) bit for bit"""
    rng = np.random.default_rng(n)
    for name, x in _cases(rng, n):
        want, tot = _seq_prefix(x)
        got = np.empty(n, dtype=np.float32)
        total = C.c_float()
        segs = C.c_long()
        for chunk in (4096, 64):
            hm.hm_exact_prefix(x.ctypes.data_as(fp), C.c_long(n), chunk, 0, got.ctypes.data_as(fp), C.byref(total),
                               C.byref(segs))
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, chunk)
            assert np.float32(total.value).view(np.uint32) == np.float32(tot).view(np.uint32), name


def test_exact_prefix_irregular_inputs(hm):
    rng = np.random.default_rng(5)
    n = 3000
    x = rng.random(n).astype(np.float32)
    x[rng.integers(0, n, 30)] = -0.5                     # negative weights: falls back to real adds
    x[rng.integers(0, n, 10)] = np.nan
    want, tot = _seq_prefix(x, skip_nan=True)
    got = np.empty(n, dtype=np.float32)
    total = C.c_float()
    hm.hm_exact_prefix(x.ctypes.data_as(fp), C.c_long(n), 256, 1, got.ctypes.data_as(fp), C.byref(total), None)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.float32(total.value).view(np.uint32) == np.float32(tot).view(np.uint32)


def test_pair_composition_is_associative(hm):
    rng = np.random.default_rng(6)
    for E in (-1, -5, -20, 0, 3, -127):
        for _ in range(20):
            n = int(rng.integers(1, 300))
            x = (rng.random(n) * 2.0 ** (E - rng.integers(0, 30, n))).astype(np.float32)
            assert hm.hm_pair_tree_equals_chain(x.ctypes.data_as(fp), n, E) == 1


def test_double_addend_chain_is_exact(hm):
    """s = (float)((double)s + d) with double addends (the lower-deviation sum, particle_filter.cpp:123) through
    inc_pair_d equals the sequential loop bit for bit"""
    hm.hm_exact_sum_double_addends.restype = C.c_long
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(11)
    cases = {
        "squares": (rng.random(200_000).astype(np.float32) * np.float32(1e-3)).astype(np.float64) ** 2,
        "float_valued": rng.random(100_000).astype(np.float32).astype(np.float64),
        "tiny_then_big": np.concatenate([np.full(1000, 1e-50), (rng.random(50_000).astype(np.float32) ** 2).astype(np.float64)]),
        "ties": (rng.integers(0, 16, 100_000) * 2.0 ** -30).astype(np.float64) + 2.0 ** -31,
        "zeros": np.zeros(1000),
    }
    for name, d in cases.items():
        d = np.ascontiguousarray(d, dtype=np.float64)
        a, b = C.c_float(), C.c_float()
        real = hm.hm_exact_sum_double_addends(d.ctypes.data_as(dp), C.c_long(len(d)), C.byref(a), C.byref(b))
        assert np.float32(a.value).view(np.uint32) == np.float32(b.value).view(np.uint32), (name, a.value, b.value)
        assert real < 2000 + 64, (name, real)      # only binade crossings (and the denormal head) take real adds


# ---- the reference's .eig map cache (top_down_map.h:29-50, top_down_map.cpp:226-286)
def test_eig_cache_format_and_round_trip(tmp_path):
    import struct
    from top_down_renderer_b200 import eigcache
    rng = np.random.default_rng(2)
    rows, cols, C = 5, 3, 2
    layers = rng.random((C, cols, rows)).astype(np.float32)
    geo = rng.random((2, cols, rows)).astype(np.float32)
    mask = (rng.random((cols, rows)) < 0.3).astype(np.uint8)
    d = str(tmp_path)
    eigcache.save_cache(d, "/maps/campus.svg", layers, geo, mask, 0.5)
    # byte layout: int64 rows, int64 cols, then COLUMN-major scalars (element (r, c) at c*rows + r)
    raw = open(f"{d}/class_map1.eig", "rb").read()
    assert struct.unpack("<qq", raw[:16]) == (rows, cols) and len(raw) == 16 + rows * cols * 4
    vals = np.frombuffer(raw[16:], dtype="<f4")
    assert vals[2 * rows + 4] == layers[1][2, 4]                      # (row 4, col 2)
    raw = open(f"{d}/class_mask.eig", "rb").read()
    assert len(raw) == 16 + rows * cols and raw[16 + 1 * rows + 3] == mask[1, 3]
    assert open(f"{d}/cached_data.txt").read() == "/maps/campus.svg\n2\n0.5\n"
    # validity rule (:233-239): same path, same class count, resolution within 0.01
    assert eigcache.cache_is_valid(d, "/maps/campus.svg", 2, 0.505)
    assert not eigcache.cache_is_valid(d, "/maps/campus.svg", 2, 0.52)
    assert not eigcache.cache_is_valid(d, "/maps/other.svg", 2, 0.5) and not eigcache.cache_is_valid(d, "/maps/campus.svg", 3, 0.5)
    assert not eigcache.cache_is_valid(str(tmp_path / "nowhere"), "/maps/campus.svg", 2, 0.5)
    l2, g2, m2 = eigcache.load_cache(d, C)
    assert np.array_equal(l2, layers) and np.array_equal(g2, geo) and np.array_equal(m2, mask)


# ---- the reference's raster cache (top_down_map.cpp:197-224): class<i>.png, 8-bit gray, flipped vertically
def test_raster_cache_png_codec_against_opencv(tmp_path):
    """the PNG files are the third-party part (cv::imwrite / cv::imread): what OpenCV writes this codec reads to the
    byte, what this codec writes OpenCV reads to the byte — for binary maps and for arbitrary gray images, whose scan
    lines make libpng pick every filter type"""
    import struct
    import zlib
    cv2 = pytest.importorskip("cv2")
    from top_down_renderer_b200 import rastercache as rc
    rng = np.random.default_rng(8)
    smooth = (np.add.outer(np.arange(97), np.arange(131)) % 256).astype(np.uint8)
    images = {"binary": (rng.random((60, 83)) < 0.4).astype(np.uint8) * 255, "noise": rng.integers(0, 256, (41, 67), dtype=np.uint8),
              "smooth": smooth, "one_pixel": np.uint8([[200]]), "blocks": np.kron(rng.integers(0, 256, (9, 11), dtype=np.uint8), np.ones((8, 8), np.uint8))}
    seen = set()
    for name, img in images.items():
        p_cv, p_own = str(tmp_path / f"{name}_cv.png"), str(tmp_path / f"{name}_own.png")
        assert cv2.imwrite(p_cv, img)
        assert np.array_equal(rc.read_png_gray(p_cv), img), name
        rc.write_png_gray(p_own, img)
        assert np.array_equal(cv2.imread(p_own, cv2.IMREAD_GRAYSCALE), img), name
        assert np.array_equal(rc.read_png_gray(p_own), img), name
        # which scan-line filters did libpng use?  (re-inflate the IDAT stream and look at the filter bytes)
        data, pos, idat = open(p_cv, "rb").read(), 8, b""
        while pos < len(data):
            n, tag = struct.unpack(">I4s", data[pos:pos + 8])
            if tag == b"IDAT":
                idat += data[pos + 8:pos + 8 + n]
            pos += 12 + n
        seen |= set(np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(img.shape[0], img.shape[1] + 1)[:, 0].tolist())
    assert 1 in seen, seen                                            # OpenCV's default: the Sub filter
    # the other filter types, from an encoder written out here (cv2 confirms the files are valid PNGs of the image)
    img = images["blocks"]
    h, w = img.shape
    for ft in (1, 2, 3, 4):
        raw = np.zeros((h, w + 1), dtype=np.uint8)
        raw[:, 0] = ft
        for y in range(h):
            for x in range(w):
                a = int(img[y, x - 1]) if x else 0
                b = int(img[y - 1, x]) if y else 0
                c = int(img[y - 1, x - 1]) if x and y else 0
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                pred = {1: a, 2: b, 3: (a + b) >> 1, 4: a if pa <= pb and pa <= pc else (b if pb <= pc else c)}[ft]
                raw[y, x + 1] = (int(img[y, x]) - pred) & 255
        path = str(tmp_path / f"filter{ft}.png")
        with open(path, "wb") as f:
            f.write(rc._SIG + rc._chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)))
            comp = zlib.compress(raw.tobytes())
            f.write(rc._chunk(b"IDAT", comp[:100]) + rc._chunk(b"IDAT", comp[100:]) + rc._chunk(b"IEND", b""))   # split IDAT
        assert np.array_equal(cv2.imread(path, cv2.IMREAD_GRAYSCALE), img), ft
        assert np.array_equal(rc.read_png_gray(path), img), ft
    cv2.imwrite(str(tmp_path / "rgb.png"), np.zeros((4, 4, 3), np.uint8))
    with pytest.raises(ValueError):
        rc.read_png_gray(str(tmp_path / "rgb.png"))
    data = bytearray(open(str(tmp_path / "noise_own.png"), "rb").read())
    data[60] ^= 1                                                     # a flipped bit inside IDAT: the CRC catches it
    open(str(tmp_path / "bad.png"), "wb").write(bytes(data))
    with pytest.raises(ValueError):
        rc.read_png_gray(str(tmp_path / "bad.png"))


def test_raster_cache_round_trip_and_orientation(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from top_down_renderer_b200 import rastercache as rc
    rng = np.random.default_rng(4)
    C, cols, rows = 3, 37, 22
    layers = (rng.random((C, cols, rows)) < 0.7).astype(np.float32)            # binary class maps, (cols, rows) = col-major rows x cols
    d = str(tmp_path / "campus_raster_cache")
    rc.save_rasterized_maps(d, layers)
    # the file looks like the input map: map row 0 (y = 0, the bottom) is the LAST image line, white = outside the class
    img = cv2.imread(f"{d}/class1.png", cv2.IMREAD_GRAYSCALE)
    assert img.shape == (rows, cols) and set(np.unique(img)) <= {0, 255}
    assert img[rows - 1, 5] == 255 * layers[1, 5, 0] and img[0, 7] == 255 * layers[1, 7, rows - 1]
    got = rc.load_rasterized_maps(d, C)
    assert got.dtype == np.float32 and np.array_equal(got, layers)            # 255 * float(1/255) == 1.0f
    # the literal OpenCV chain of loadRasterizedMaps on the same file: flip, convertTo(CV_32FC1, 1./255)
    want = (cv2.flip(img, 0).astype(np.float32) * np.float32(1.0 / 255)).T
    assert np.array_equal(got[1], want)


def test_lean_lattice_coordinate_equals_the_literal_form(hm):
    """tdr_math.cuh lattice_coord (interval test + trunc, what the tensor-core kernels run) against
    f2i_x86(round_half_away(v)) followed by 0 <= index < n (top_down_map_polar.cpp:31-37): every float near every
    rounding boundary of a 4000-px axis, the border values, huge / tiny / non-finite values, and 4e6 random ones"""
    hm.hm_lattice_coord_mismatches.restype = C.c_long
    hm.hm_lattice_coord_mismatches.argtypes = [C.POINTER(C.c_float), C.c_long, C.c_int]
    hm.hm_lattice_coord.argtypes = [C.c_float, C.c_int]
    hm.hm_lattice_fixed_mismatches.restype = C.c_long
    hm.hm_lattice_fixed_mismatches.argtypes = [C.POINTER(C.c_float), C.c_long, C.c_int]
    for limit in (1, 2, 1000, 4000, 4001):
        k = np.arange(-3, limit + 3, dtype=np.float64)
        near = []
        for off in (0.0, 0.5, -0.5):
            c = (k + off).astype(np.float32)
            for _ in range(3):                               # the boundary and three floats on either side of it
                near += [c]
                c = np.nextafter(c, np.float32(np.inf))
            c = (k + off).astype(np.float32)
            for _ in range(3):
                c = np.nextafter(c, np.float32(-np.inf))
                near += [c]
        special = np.float32([0.0, -0.0, np.nan, np.inf, -np.inf, 1e30, -1e30, 2147483648.0, -2147483648.0, 2147483520.0,
                              1e-38, -1e-38, 1e-45, -1e-45, 8388608.0, 16777216.0, -0.49999997, 0.49999997])
        rng = np.random.default_rng(limit)
        rnd = np.concatenate([rng.uniform(-10, limit + 10, 2_000_000), rng.normal(limit / 2, limit, 2_000_000)]).astype(np.float32)
        v = np.ascontiguousarray(np.concatenate(near + [special, rnd]), dtype=np.float32)
        assert hm.hm_lattice_coord_mismatches(v.ctypes.data_as(C.POINTER(C.c_float)), len(v), limit) == 0, limit
        # and the fixed-point form of the integer-record kernel (coordinates pre-multiplied by 4096)
        assert hm.hm_lattice_fixed_mismatches(v.ctypes.data_as(C.POINTER(C.c_float)), len(v), limit) == 0, limit
    # the documented boundary cases
    assert hm.hm_lattice_coord(-0.5, 10) == -1 and hm.hm_lattice_coord(np.nextafter(np.float32(-0.5), np.float32(0)), 10) == 0
    assert hm.hm_lattice_coord(9.5, 10) == -1 and hm.hm_lattice_coord(np.nextafter(np.float32(9.5), np.float32(0)), 10) == 9
    assert hm.hm_lattice_coord(0.5, 10) == 1 and hm.hm_lattice_coord(0.49999997, 10) == 0 and hm.hm_lattice_coord(float("nan"), 10) == -1


def test_integer_record_format_error_bound():
    """The 16-byte map record of csrc/score_mma_i8.cu in numpy: v_c = round(w_c dist_c / q) split into hi / lo bytes, u8 scan
    counts, exact integer sums  ->  cost = 0.01 q (256 Xhi + Xlo) / norm.  The claims the host relies on when it picks that
    kernel: |cost - exact cost| <= 0.005 q, hence a relative weight error <= 0.005 q / regularization, below 9e-6 whenever
    regularization >= 0.42 max(w) — and the hi / lo split loses nothing (256 hi + lo == v)."""
    rng = np.random.default_rng(5)
    C, P, S = 6, 2500, 40
    w = rng.uniform(0.2, 1.0, C)
    wmax = float(w.max())
    q = 50.0 * wmax / 65535.0
    worst_cost, worst_rel = 0.0, 0.0
    for trial in range(20):
        dist = np.minimum(rng.exponential(15.0, (P, C)), 50.0)              # class distances, capped like the map's
        x = w[None, :] * dist                                               # what the fp32 kernels multiply the counts with
        v = np.clip(np.rint(x / q), 0, 65535).astype(np.int64)
        hi, lo = v >> 8, v & 255
        assert np.array_equal(256 * hi + lo, v) and hi.max() <= 255
        known = (rng.random(P) > 0.1).astype(np.int64)
        counts = rng.integers(0, 256, (S, P, C)) * (rng.random((S, P, 1)) < 0.3)   # sparse scan, counts fit a byte
        tot = counts.sum(2)
        xhi = (counts * (hi * known[:, None])[None]).sum((1, 2))
        xlo = (counts * (lo * known[:, None])[None]).sum((1, 2))
        norm = (tot * known[None]).sum(1)
        ok = norm > 0
        cost_int = 0.01 * q * (256.0 * xhi[ok] + xlo[ok]) / norm[ok]
        cost_ref = 0.01 * (counts * (x * known[:, None])[None]).sum((1, 2))[ok] / norm[ok]
        worst_cost = max(worst_cost, float(np.abs(cost_int - cost_ref).max()))
        reg = 0.42 * wmax
        worst_rel = max(worst_rel, float((np.abs(1 / (cost_int + reg) - 1 / (cost_ref + reg)) * (cost_ref + reg)).max()))
    assert worst_cost <= 0.005 * q * (1 + 1e-9), (worst_cost, 0.005 * q)
    assert worst_rel <= 9e-6, worst_rel


def test_packed_edt_row_pass_model():
    """The arithmetic of k_edt_rows_dpx (csrc/map_build.cu) in numpy: 16-bit words g^2 with 0x7fff for "no seed", squared
    offsets up to (R + 7)^2 added WITHOUT saturation (the DPX add wraps at 2^16 — so nothing may reach it), minimum over
    the 2R + 8 taps a thread's eight pixels share, then min(d2, capcode).  Must equal the windowed scalar scan of
    k_edt_horizontal for every resolution the packed pass accepts."""
    rng = np.random.default_rng(8)
    for resolution in (2.0, 1.0, 0.5, 0.3):
        capcode = next(d2 for d2 in range(1 << 20) if np.float32(np.sqrt(np.float32(d2))) * np.float32(resolution) >= np.float32(50.0))
        rcap = int(np.ceil(np.sqrt(capcode))) + 1
        R = (rcap + 3) // 4 * 4
        assert (R + 7) ** 2 <= 0x7fff and capcode <= 0x7fff
        cols = 700
        g = rng.integers(0, rcap + 1, cols)
        g[rng.random(cols) < 0.85] = 255                                      # sparse seeds
        g2 = np.where(g == 255, 0x7fff, g * g).astype(np.int64)
        assert (g2.max() + (R + 7) ** 2) < (1 << 16)                          # the unsaturated 16-bit add cannot wrap
        pad = np.full(R + 8, 0x7fff, dtype=np.int64)
        row = np.concatenate([pad, g2, pad])
        got = np.empty(cols, dtype=np.int64)
        for xb in range(0, cols, 8):                                          # a thread's eight pixels and their shared taps
            taps = row[xb + 8: xb + 8 + 2 * R + 8]                            # source columns xb - R .. xb + R + 7
            src = np.arange(xb - R, xb + R + 8)
            for p in range(min(8, cols - xb)):
                d = src - (xb + p)
                got[xb + p] = min(int((taps + d * d).min()), capcode)
        want = np.empty(cols, dtype=np.int64)
        for x in range(cols):
            best = 0x3fffffff
            for dx in range(-rcap, rcap + 1):
                if 0 <= x + dx < cols and g[x + dx] != 255:
                    best = min(best, dx * dx + int(g[x + dx]) ** 2)
            want[x] = min(best, capcode)
        assert np.array_equal(got, want), resolution


def test_pair_constructions_agree_for_float_addends(hm):
    """tdr_math.cuh: inc_pair (32-bit arithmetic) and inc_pair_d (64-bit, for double addends) must give the same
    (increment-if-even, increment-if-odd) pair — and the same "irregular" verdict — for every float addend in every
    binade of the running sum: weights.cu picks the cheap form for chains of plain float weights."""
    hm.hm_pair_forms_mismatches.restype = C.c_long
    hm.hm_pair_forms_mismatches.argtypes = [C.POINTER(C.c_float), C.c_long]
    rng = np.random.default_rng(17)
    bits = rng.integers(0, 1 << 32, 40000, dtype=np.uint64).astype(np.uint32)          # every kind of float, NaN / inf / negative / denormal included
    special = np.array([0x00000000, 0x80000000, 0x00000001, 0x007fffff, 0x00800000, 0x3f800000, 0x3f800001, 0x7f7fffff, 0x7f800000,
                        0x7fc00000, 0xff800000, 0x33800000, 0x34000000, 0x4b000000, 0x4b800000], dtype=np.uint32)
    halves = (np.arange(1, 255, dtype=np.uint32) << 23)                                 # exact powers of two: the tie cases
    w = np.ascontiguousarray(np.concatenate([bits, special, halves, halves | 0x00400000]).view(np.float32))
    assert hm.hm_pair_forms_mismatches(w.ctypes.data_as(C.POINTER(C.c_float)), len(w)) == 0
