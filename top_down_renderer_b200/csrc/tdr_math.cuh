// tdr_math.cuh — index-producing arithmetic shared by every kernel.
//
// Everything here is written so that it evaluates to the same bits on the device
// (nvcc, -fmad=false, IEEE div/sqrt, no FTZ) and on the host (g++ -O2
// -ffp-contract=off) — the host build is what tests/test_host_math.py checks
// against glibc (atan2f) and against sequential fp32 loops (the exact-sum algebra).
#pragma once
#include <stdint.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define TDR_HD __host__ __device__ __forceinline__
#else
#define TDR_HD inline
#endif

// Non-contractable fp32 ops.  On the device these are the _rn intrinsics (never fused
// into FFMA); on the host plain ops under -ffp-contract=off.
#if defined(__CUDA_ARCH__)
#define TDR_FMUL(a, b) __fmul_rn((a), (b))
#define TDR_FADD(a, b) __fadd_rn((a), (b))
#define TDR_FSUB(a, b) __fsub_rn((a), (b))
#define TDR_FDIV(a, b) __fdiv_rn((a), (b))
#define TDR_FSQRT(a) __fsqrt_rn((a))
#else
#define TDR_FMUL(a, b) ((float)(a) * (float)(b))
#define TDR_FADD(a, b) ((float)(a) + (float)(b))
#define TDR_FSUB(a, b) ((float)(a) - (float)(b))
#define TDR_FDIV(a, b) ((float)(a) / (float)(b))
#define TDR_FSQRT(a) sqrtf((a))
#endif

namespace tdr {

TDR_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
TDR_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// x86-64 cvttss2si: NaN / out of range -> INT_MIN.  The reference's float->int
// assignments (scan_renderer_polar.cpp:100-103, top_down_map_polar.cpp:31) rely on it;
// CUDA's cvt.rzi saturates and maps NaN to 0, which would turn a NaN point into bin 0.
TDR_HD int f2i_x86(float v) {
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)v;
}

// std::round(float): half away from zero.  x - trunc(x) is exact in fp32.
TDR_HD float round_half_away(float x) {
  float t = truncf(x);
  float d = x - t;  // exact
  if (fabsf(d) >= 0.5f) t += (x < 0.f ? -1.f : 1.f);
  return t;
}

// ---------------------------------------------------------------------------------------
// glibc 2.39 atan2f == fdlibm fp32 e_atan2f.c + s_atanf.c evaluated with plain IEEE fp32
// operations and no FMA.  Verified bit-for-bit against this image's libm in
// tests/test_host_math.py.
// ---------------------------------------------------------------------------------------
TDR_HD float fdlibm_atanf(float x) {
  // decimal literals exactly as in fdlibm's s_atanf.c (its hex comments are off by one ulp for some
  // entries; the compiler sees the decimals)
  const float atanhi0 = (float)4.6364760399e-01, atanhi1 = (float)7.8539812565e-01,
              atanhi2 = (float)9.8279368877e-01, atanhi3 = (float)1.5707962513e+00;
  const float atanlo0 = (float)5.0121582440e-09, atanlo1 = (float)3.7748947079e-08,
              atanlo2 = (float)3.4473217170e-08, atanlo3 = (float)7.5497894159e-08;
  const float aT0 = (float)3.3333334327e-01, aT1 = (float)-2.0000000298e-01, aT2 = (float)1.4285714924e-01,
              aT3 = (float)-1.1111110449e-01, aT4 = (float)9.0908870101e-02, aT5 = (float)-7.6918758452e-02,
              aT6 = (float)6.6610731184e-02, aT7 = (float)-5.8335702866e-02, aT8 = (float)4.9768779427e-02,
              aT9 = (float)-3.6531571299e-02, aT10 = (float)1.6285819933e-02;
  const float one = 1.0f;
  int32_t hx = (int32_t)f2u(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  float hi = 0.f, lo = 0.f;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return TDR_FADD(x, x);  // NaN
    float r = TDR_FADD(atanhi3, atanlo3);
    return hx > 0 ? r : -r;
  }
  if (ix < 0x3ee00000) {   // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {        // |x| < 1.1875
      if (ix < 0x3f300000) {      // 7/16 <= |x| < 11/16
        id = 0; hi = atanhi0; lo = atanlo0;
        x = TDR_FDIV(TDR_FSUB(TDR_FMUL(2.0f, x), one), TDR_FADD(2.0f, x));
      } else {                    // 11/16 <= |x| < 19/16
        id = 1; hi = atanhi1; lo = atanlo1;
        x = TDR_FDIV(TDR_FSUB(x, one), TDR_FADD(x, one));
      }
    } else {
      if (ix < 0x401c0000) {      // |x| < 2.4375
        id = 2; hi = atanhi2; lo = atanlo2;
        x = TDR_FDIV(TDR_FSUB(x, 1.5f), TDR_FADD(one, TDR_FMUL(1.5f, x)));
      } else {                    // 2.4375 <= |x| < 2^25
        id = 3; hi = atanhi3; lo = atanlo3;
        x = TDR_FDIV(-1.0f, x);
      }
    }
  }
  float z = TDR_FMUL(x, x);
  float w = TDR_FMUL(z, z);
  // s1 = z*(aT[0]+w*(aT[2]+w*(aT[4]+w*(aT[6]+w*(aT[8]+w*aT[10])))))
  float s1 = TDR_FMUL(w, aT10);
  s1 = TDR_FMUL(w, TDR_FADD(aT8, s1));
  s1 = TDR_FMUL(w, TDR_FADD(aT6, s1));
  s1 = TDR_FMUL(w, TDR_FADD(aT4, s1));
  s1 = TDR_FMUL(w, TDR_FADD(aT2, s1));
  s1 = TDR_FMUL(z, TDR_FADD(aT0, s1));
  // s2 = w*(aT[1]+w*(aT[3]+w*(aT[5]+w*(aT[7]+w*aT[9]))))
  float s2 = TDR_FMUL(w, aT9);
  s2 = TDR_FMUL(w, TDR_FADD(aT7, s2));
  s2 = TDR_FMUL(w, TDR_FADD(aT5, s2));
  s2 = TDR_FMUL(w, TDR_FADD(aT3, s2));
  s2 = TDR_FMUL(w, TDR_FADD(aT1, s2));
  float xs = TDR_FMUL(x, TDR_FADD(s1, s2));
  if (id < 0) return TDR_FSUB(x, xs);
  z = TDR_FSUB(hi, TDR_FSUB(TDR_FSUB(xs, lo), x));
  return (hx < 0) ? -z : z;
}

TDR_HD float fdlibm_atan2f(float y, float x) {
  const float tiny = 1.0e-30f;
  const float pi_o_4 = (float)7.8539818525e-01, pi_o_2 = (float)1.5707963705e+00, pi = (float)3.1415927410e+00,
              pi_lo = (float)-8.7422776573e-08;
  int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return TDR_FADD(x, y);  // NaN
  if (hx == 0x3f800000) return fdlibm_atanf(y);                   // x = 1.0
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                    // 2*sign(x)+sign(y)
  if (iy == 0) {                                                  // y = 0
    switch (m) {
      case 0: case 1: return y;
      case 2: return TDR_FADD(pi, tiny);
      default: return TDR_FSUB(-pi, tiny);
    }
  }
  if (ix == 0) return (hy < 0) ? TDR_FSUB(-pi_o_2, tiny) : TDR_FADD(pi_o_2, tiny);  // x = 0
  if (ix == 0x7f800000) {                                         // x = INF
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return TDR_FADD(pi_o_4, tiny);
        case 1: return TDR_FSUB(-pi_o_4, tiny);
        case 2: return TDR_FADD(TDR_FMUL(3.0f, pi_o_4), tiny);
        default: return TDR_FSUB(TDR_FMUL(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return TDR_FADD(pi, tiny);
        default: return TDR_FSUB(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? TDR_FSUB(-pi_o_2, tiny) : TDR_FADD(pi_o_2, tiny);  // y = INF
  int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = TDR_FADD(pi_o_2, TDR_FMUL(0.5f, pi_lo));        // |y/x| > 2^60
  else if (hx < 0 && k < -60) z = 0.0f;                           // |y|/x < -2^60
  else z = fdlibm_atanf(fabsf(TDR_FDIV(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return TDR_FSUB(pi, TDR_FSUB(z, pi_lo));
    default: return TDR_FSUB(TDR_FSUB(z, pi_lo), pi);
  }
}

// ---------------------------------------------------------------------------------------
// a1 bin index   scan_renderer_polar.cpp:95-102.  Returns false when the point is dropped.
// ---------------------------------------------------------------------------------------
TDR_HD bool polar_bin(float x, float y, float res, float ang_res, int n_theta, int n_r, int* ti, int* ri) {
  if (x == 0.f && y == 0.f) return false;
  float theta = fdlibm_atan2f(x, y);                                       // atan2(pt.x, pt.y)
  float r = TDR_FSQRT(TDR_FADD(TDR_FMUL(x, x), TDR_FMUL(y, y)));
  int t = f2i_x86(TDR_FADD(round_half_away(TDR_FDIV(theta, ang_res)), (float)(n_theta / 2)));
  int rr = f2i_x86(round_half_away(TDR_FDIV(r, res)));
  if (t >= 0 && t < n_theta && rr >= 0 && rr < n_r) { *ti = t; *ri = rr; return true; }
  return false;
}
// a2 bin index   scan_renderer.cpp:67-71
TDR_HD bool cart_bin(float x, float y, float res, int rows, int cols, int* xi, int* yi) {
  if (x == 0.f && y == 0.f) return false;
  int xx = f2i_x86(TDR_FADD(round_half_away(TDR_FDIV(x, res)), (float)(cols / 2)));
  int yy = f2i_x86(TDR_FADD(round_half_away(TDR_FDIV(y, res)), (float)(rows / 2)));
  if (xx >= 0 && xx < cols && yy >= 0 && yy < rows) { *xi = xx; *yi = yy; return true; }
  return false;
}

// a7 lattice pixel   top_down_map_polar.cpp:28-31: round(((tab*scale)*res) + centre/resolution)
TDR_HD int lattice_index(float tab, float scale, float res, float off) {
  return f2i_x86(round_half_away(TDR_FADD(TDR_FMUL(TDR_FMUL(tab, scale), res), off)));
}

// the same decision without the detour through x86 conversion semantics: f2i_x86(round_half_away(v)) lies in [0, n)
// exactly when -0.5 < v < n - 0.5 (v = -0.5 rounds away to -1, v = n - 0.5 to n, NaN fails both comparisons), and inside
// that interval it is trunc(v) + (v - trunc(v) >= 0.5).  hi = n - 0.5 (exact in fp32 for n < 2^23).  Returns -1 off the
// range.  Proven equal to the literal form on the CPU (tests/test_host_math.py).
TDR_HD bool lattice_in(float v, float hi) { return v > -0.5f && v < hi; }
TDR_HD int lattice_round(float v) {                    // round_half_away(v) as an int, for v inside the interval only
  const float t = truncf(v);
  return (int)t + (TDR_FSUB(v, t) >= 0.5f ? 1 : 0);
}
TDR_HD int lattice_coord(float v, float hi) { return lattice_in(v, hi) ? lattice_round(v) : -1; }

// The same decision once more, in fixed point, for a coordinate that arrives pre-multiplied by 4096 (exact: scaling both
// addends of tab*scale*res + centre by a power of two commutes with the rounding of their sum).  With t = trunc(v4096) +
// 2047, -0.5 < v < n - 0.5 holds exactly when 0 <= t <= 4096 n - 2, and then round_half_away(v) = (t + 1) >> 12:
// FADD, F2I, IADD, ISETP, IADD, SHF per coordinate (score_mma_i8.cu is issue-bound on its gather threads).  The
// conversion saturates (cvt.rzi.s32.f32): huge and infinite values wrap past the limit and are off the map like in the
// reference; NaN converts to 0 — callers mark a hypothesis with a NaN centre inactive themselves.  n < 2^19.
TDR_HD int f2i_sat(float v) {
#if defined(__CUDA_ARCH__)
  return __float2int_rz(v);
#else
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT_MAX;
  if (v <= -2147483648.0f) return INT_MIN;
  return (int)v;
#endif
}
TDR_HD int lattice_fixed(float v4096, uint32_t lim /* 4096 n - 1 */) {
  const uint32_t t = (uint32_t)f2i_sat(v4096) + 2047u;
  return t < lim ? (int)((t + 1u) >> 12) : -1;
}

// a10 rot -> row shift   state_particle.cpp:123-128
TDR_HD int rot_to_shift(float rot, int n_theta) {
  double v = (double)TDR_FDIV(TDR_FMUL(rot, (float)n_theta), 2.0f) / 3.14159265358979323846;
  double r = round(v);
  if (!(r >= -2147483648.0 && r < 2147483648.0)) return 0;
  int s = (int)r % n_theta;
  if (s < 0) s += n_theta;
  return s;
}

// a4 value of a stored squared distance   top_down_map.cpp:312-315
TDR_HD float dist_value(uint32_t d2, float resolution) {
  float d = TDR_FMUL(TDR_FSQRT((float)d2), resolution);
  return d > 50.0f ? 50.0f : d;
}

// ---------------------------------------------------------------------------------------
// Order-exact emulation of a sequential fp32 accumulation  s = RN(s + w_j), w_j >= 0
// (particle_filter.cpp:110-116,135,142,175-183,195-197).
//
// While s stays inside one binade [2^E, 2^(E+1)) its ulp u = 2^(E-23) is fixed and
// s = m*u with integer m in [2^23, 2^24).  Adding w: w/u = q + f (q integer, f in [0,1)):
//     RN(s + w) = (m + q + rnd) * u,   rnd = 1 if f > 1/2, 0 if f < 1/2,
//                                      and on an exact tie rnd = (m + q) & 1   (ties-to-even).
// So each element is a map  m -> m + (m even ? a : b)  with
//     non-tie: a = b = q + [f > 1/2];   tie: a = q + (q & 1), b = q + 1 - (q & 1).
// These maps compose associatively (pair_compose), which turns the dependent FADD chain into a
// parallel scan.  A binade change (m reaches 2^24) or an irregular addend (negative, NaN, inf,
// larger than the binade) ends the segment; the single crossing add is done with a real fp32
// add and the scan restarts in the new binade.  The denormal range is the "binade" E = -127:
// u = 2^-149, m in [0, 2^23) — same algebra, every add exact.
// ---------------------------------------------------------------------------------------
struct IncPair { uint32_t a, b; };           // increments for even / odd m, saturating at SAT
static const uint32_t TDR_INC_SAT = 1u << 30;  // anything >= 2^24 means "left the binade"

TDR_HD uint32_t sat_add(uint32_t x, uint32_t y) { uint32_t s = x + y; return s > TDR_INC_SAT ? TDR_INC_SAT : s; }

// apply first p then q
TDR_HD IncPair pair_compose(IncPair p, IncPair q) {
  IncPair r;
  r.a = sat_add(p.a, (p.a & 1u) ? q.b : q.a);        // m even -> m + p.a has parity of p.a
  r.b = sat_add(p.b, (p.b & 1u) ? q.a : q.b);        // m odd  -> m + p.b has parity 1 ^ (p.b & 1)
  return r;
}

// binade exponent of a non-negative finite float, -127 for zero/denormals; and its integer m
TDR_HD int binade_of(float s) { uint32_t e = (f2u(s) >> 23) & 0xff; return e == 0 ? -127 : (int)e - 127; }
TDR_HD uint32_t mant_of(float s) {
  uint32_t u = f2u(s); uint32_t e = (u >> 23) & 0xff; uint32_t m = u & 0x7fffffu;
  return e == 0 ? m : (m | 0x800000u);
}
// rebuild the float from (E, m): valid for m < 2^24 (E > -127) or m < 2^23 (E == -127)
TDR_HD float from_binade(int E, uint32_t m) {
  if (E == -127) return u2f(m);                       // denormal / zero (m may reach 2^23 = smallest normal: still right)
  return u2f(((uint32_t)(E + 127) << 23) + (m - 0x800000u));  // m == 2^24 carries into the exponent correctly
}

// per-element pair for binade E.  `irregular` is set when the element cannot be handled inside
// the binade algebra (negative / NaN / inf); such elements always force a real add.
TDR_HD IncPair inc_pair(float w, int E, bool* irregular) {
  IncPair r; r.a = r.b = 0; *irregular = false;
  uint32_t u = f2u(w);
  if (u == 0u) return r;                              // +0
  if (u == 0x80000000u) return r;                     // -0: s + (-0) = s for s >= +0
  if ((u >> 31) || ((u >> 23) & 0xff) == 0xff) { *irregular = true; r.a = r.b = TDR_INC_SAT; return r; }
  int ew = binade_of(w);
  uint32_t mw = mant_of(w);                           // w = mw * 2^(ew_eff - 23), ew_eff = max(ew,-126)
  int ew_eff = ew == -127 ? -126 : ew;
  int E_eff = E == -127 ? -126 : E;
  int sh = E_eff - ew_eff;                            // w / u = mw * 2^(-sh)
  if (sh < 0) {                                       // w >= 2 * 2^E: certainly leaves the binade
    r.a = r.b = TDR_INC_SAT; return r;
  }
  if (sh == 0) { r.a = r.b = mw; return r; }          // exact, no rounding
  if (sh > 25) return r;                              // f < 1/2, q = 0
  uint32_t q = (sh >= 32) ? 0u : (mw >> sh);
  uint32_t rem = mw & ((1u << sh) - 1u);
  uint32_t half = 1u << (sh - 1);
  if (rem > half) { r.a = r.b = q + 1u; }
  else if (rem < half) { r.a = r.b = q; }
  else { r.a = q + (q & 1u); r.b = q + 1u - (q & 1u); }
  return r;
}

// The same pair for an accumulation whose addend is a DOUBLE:  s = (float)((double)s + d), d >= 0
// (particle_filter.cpp:123: bottom_stddev += std::pow(w - mean, 2) with a float accumulator).  The sum is first
// rounded to double: inside binade E that grid is 2^(E-52), and because s is a multiple of 2^(E-23) the rounded
// addend d' does not depend on s; the float rounding of s + d' then follows the usual rule.  For a float-valued
// d this reduces to inc_pair.  strict_denormal: in the denormal binade a double addend can carry bits below the
// double grid assumed here, so every nonzero addend forces a real add (never more than a handful of elements).
TDR_HD IncPair inc_pair_d(double d, int E, bool strict_denormal, bool* irregular) {
  IncPair r; r.a = r.b = 0; *irregular = false;
  uint64_t u;
#if defined(__CUDA_ARCH__)
  u = (uint64_t)__double_as_longlong(d);
#else
  memcpy(&u, &d, 8);
#endif
  if ((u << 1) == 0) return r;                                   // +-0
  const uint32_t ex = (uint32_t)((u >> 52) & 0x7ff);
  if ((u >> 63) || ex == 0x7ff) { *irregular = true; r.a = r.b = TDR_INC_SAT; return r; }
  if (E == -127 && strict_denormal) { r.a = r.b = TDR_INC_SAT; return r; }
  const int E_eff = E == -127 ? -126 : E;
  // d = md * 2^ed
  const uint64_t md = ex ? ((u & 0xfffffffffffffull) | (1ull << 52)) : (u & 0xfffffffffffffull);
  const int ed = (ex ? (int)ex : 1) - 1075;
  // d >= 2^(E_eff+1) certainly leaves the binade
  const int msb = 63 - (int)
#if defined(__CUDA_ARCH__)
      __clzll((long long)md);
#else
      __builtin_clzll(md);
#endif
  if (msb + ed >= E_eff + 1) { r.a = r.b = TDR_INC_SAT; return r; }
  // D = RN_even(d / 2^(E_eff - 52)), the double-rounded addend in units of the double grid
  const int shift = (E_eff - 52) - ed;
  uint64_t D;
  if (shift <= 0) D = md << (-shift);                            // fits: d < 2^(E_eff+1)  ->  D < 2^53
  else if (shift > 54) D = 0;
  else {
    const uint64_t q0 = md >> shift, rem0 = md & ((1ull << shift) - 1ull), half0 = 1ull << (shift - 1);
    D = q0 + ((rem0 > half0 || (rem0 == half0 && (q0 & 1ull))) ? 1ull : 0ull);
  }
  // float grid = 2^29 double-grid units
  const uint32_t q = (uint32_t)(D >> 29);
  const uint32_t rem = (uint32_t)(D & ((1u << 29) - 1u)), half = 1u << 28;
  if (rem > half) { r.a = r.b = q + 1u; }
  else if (rem < half) { r.a = r.b = q; }
  else { r.a = q + (q & 1u); r.b = q + 1u - (q & 1u); }
  return r;
}

// upper limit of m inside binade E (exclusive)
TDR_HD uint32_t binade_limit(int E) { return E == -127 ? 0x800000u : 0x1000000u; }

}  // namespace tdr
