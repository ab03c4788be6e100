// hostmath_test.cpp — host build of tdr_math.cuh for CPU unit tests (tests/test_host_math.py).
// It exercises the SAME inline functions the kernels use (index arithmetic, fdlibm atan2f, the
// exact-sum pair algebra).  Not part of the product path: the product library is libtdr_b200.so.
#include <vector>
#include "tdr_math.cuh"

using namespace tdr;

extern "C" {

float hm_atan2f(float y, float x) { return fdlibm_atan2f(y, x); }

// bulk compare against libm: returns the number of mismatching bit patterns
long hm_atan2f_mismatches(const float* y, const float* x, long n) {
  long bad = 0;
  for (long i = 0; i < n; i++) {
    float a = fdlibm_atan2f(y[i], x[i]);
    float b = atan2f(y[i], x[i]);
    if (f2u(a) != f2u(b) && !(a != a && b != b)) bad++;
  }
  return bad;
}

int hm_polar_bin(float x, float y, float res, float ang_res, int n_theta, int n_r, int* ti, int* ri) {
  return polar_bin(x, y, res, ang_res, n_theta, n_r, ti, ri) ? 1 : 0;
}
int hm_cart_bin(float x, float y, float res, int rows, int cols, int* xi, int* yi) {
  return cart_bin(x, y, res, rows, cols, xi, yi) ? 1 : 0;
}
int hm_lattice_index(float tab, float scale, float res, float off) { return lattice_index(tab, scale, res, off); }
// the lean lattice coordinate of the tensor-core kernels against the literal form, in bulk: number of disagreements
long hm_lattice_coord_mismatches(const float* v, long n, int limit) {
  long bad = 0;
  const float hi = (float)limit - 0.5f;
  for (long i = 0; i < n; i++) {
    const int lit = f2i_x86(round_half_away(v[i]));
    const int want = (lit >= 0 && lit < limit) ? lit : -1;
    if (lattice_coord(v[i], hi) != want) bad++;
  }
  return bad;
}
// the fixed-point form of the integer-record kernel (score_mma_i8.cu), fed with v * 4096 (exact)
long hm_lattice_fixed_mismatches(const float* v, long n, int limit) {
  long bad = 0;
  for (long i = 0; i < n; i++) {
    if (v[i] != v[i]) continue;                 // NaN centres are switched off by the caller
    const int lit = f2i_x86(round_half_away(v[i]));
    const int want = (lit >= 0 && lit < limit) ? lit : -1;
    if (lattice_fixed(v[i] * 4096.0f, 4096u * (uint32_t)limit - 1u) != want) bad++;
  }
  return bad;
}
int hm_lattice_coord(float v, int limit) { return lattice_coord(v, (float)limit - 0.5f); }
int hm_rot_to_shift(float rot, int n_theta) { return rot_to_shift(rot, n_theta); }
float hm_round(float x) { return round_half_away(x); }
int hm_f2i(float x) { return f2i_x86(x); }
float hm_dist_value(unsigned d2, float res) { return dist_value(d2, res); }

// CPU emulation of k_exact_seq: identical control flow (binade segments, crossing adds, irregular
// tail), with the block scan replaced by a left-to-right composition of the same IncPairs.
void hm_exact_prefix(const float* x, long n, int chunk, int skip_nan, float* runmax_out, float* total_out,
                     long* n_segments) {
  float S = 0.f, rmax = -INFINITY;
  long pos = 0, segs = 0;
  int mode = 0;
  std::vector<IncPair> pref(chunk);
  while (pos < n) {
    if (mode == 1) {
      for (long j = pos; j < n; j++) {
        float w = x[j];
        if (skip_nan && w != w) w = 0.f;
        S = S + w;
        if (S > rmax) rmax = S;
        if (runmax_out) runmax_out[j] = rmax;
      }
      pos = n;
      break;
    }
    segs++;
    const int E = binade_of(S);
    const uint32_t m_in = mant_of(S), limit = binade_limit(E);
    const int len = (int)((n - pos) < chunk ? (n - pos) : chunk);
    IncPair agg; agg.a = agg.b = 0;
    for (int j = 0; j < len; j++) {
      float w = x[pos + j];
      if (skip_nan && w != w) w = 0.f;
      bool irr;
      agg = pair_compose(agg, inc_pair(w, E, &irr));
      pref[j] = agg;
    }
    const bool odd = (m_in & 1u) != 0;
    int cross = len;
    uint32_t mprev = m_in;
    for (int j = 0; j < len; j++) {
      uint32_t m = m_in + (odd ? pref[j].b : pref[j].a);
      if (m >= limit) { cross = j; break; }
      float Sj = from_binade(E, m);
      if (runmax_out) runmax_out[pos + j] = Sj > rmax ? Sj : rmax;
      mprev = m;
    }
    if (cross > 0) { S = from_binade(E, mprev); if (S > rmax) rmax = S; }
    pos += cross;
    if (cross < len) {
      float w = x[pos];
      if (skip_nan && w != w) w = 0.f;
      S = S + w;
      if (S > rmax) rmax = S;
      if (runmax_out) runmax_out[pos] = rmax;
      pos++;
      if (!(S >= 0.f) || S == INFINITY) mode = 1;
    }
  }
  if (total_out) *total_out = S;
  if (n_segments) *n_segments = segs;
}

// sequential  s = (float)((double)s + d[i])  vs the inc_pair_d algebra (one binade segment at a time)
long hm_exact_sum_double_addends(const double* d, long n, float* total_seq, float* total_alg) {
  float s = 0.f;
  for (long i = 0; i < n; i++) s = (float)((double)s + d[i]);
  *total_seq = s;
  float S = 0.f;
  long real_adds = 0;
  long i = 0;
  while (i < n) {
    const int E = binade_of(S);
    uint32_t m = mant_of(S);
    const uint32_t limit = binade_limit(E);
    bool crossed = false;
    while (i < n) {
      bool irr;
      IncPair pr = inc_pair_d(d[i], E, true, &irr);
      uint32_t inc = (m & 1u) ? pr.b : pr.a;
      if (m + inc >= limit || inc >= TDR_INC_SAT) { crossed = true; break; }
      m += inc; i++;
    }
    S = from_binade(E, m);
    if (crossed) { S = (float)((double)S + d[i]); i++; real_adds++; }
  }
  *total_alg = S;
  return real_adds;
}

// associativity probe: compose pairs in a balanced tree instead of left-to-right and compare
int hm_pair_tree_equals_chain(const float* x, int n, int E) {
  std::vector<IncPair> v(n);
  bool irr;
  for (int i = 0; i < n; i++) v[i] = inc_pair(x[i], E, &irr);
  IncPair chain; chain.a = chain.b = 0;
  for (int i = 0; i < n; i++) chain = pair_compose(chain, v[i]);
  std::vector<IncPair> t = v;
  int len = n;
  while (len > 1) {
    int o = 0;
    for (int i = 0; i + 1 < len; i += 2) t[o++] = pair_compose(t[i], t[i + 1]);
    if (len & 1) t[o++] = t[len - 1];
    len = o;
  }
  return (n == 0) || (t[0].a == chain.a && t[0].b == chain.b);
}

// inc_pair (32-bit) against inc_pair_d (64-bit) for float-valued addends: the order-exact sums use the former for chains of
// plain float weights and the latter where the addends are doubles; for a float they must be the same pair in every binade
long hm_pair_forms_mismatches(const float* w, long n) {
  long bad = 0;
  for (long i = 0; i < n; i++) {
    for (int E = -127; E <= 127; E++) {
      bool i1, i2;
      const IncPair a = inc_pair(w[i], E, &i1);
      const IncPair b = inc_pair_d((double)w[i], E, false, &i2);
      if (a.a != b.a || a.b != b.b || i1 != i2) bad++;
    }
  }
  return bad;
}

}  // extern "C"
