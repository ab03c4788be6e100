// weights.cu — a11/a12: weight normalisation and systematic resampling.
//
// Reference: ParticleFilter::update (src/particle_filter.cpp:107-147 weights, :172-187 resample).
//
// The reference's sums are SEQUENTIAL fp32 accumulations and its resampler compares samples with
// a sequential fp32 prefix, so bit-exact indices need the same prefix VALUES.  k_exact_seq
// reproduces a sequential fp32 accumulation exactly, in parallel, with the binade algebra of
// tdr_math.cuh (IncPair scan), restarting at the ~20-30 binade crossings of a run.
#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

// One job = one sequential accumulation chain  s_j = RN(s_{j-1} + x[start + j*stride]), j < count.
struct SeqJob {
  const float* x; long long start, stride, count;
  float* runmax_out;   // optional: max_{k<=j} s_k per element (contiguous, length count)
  float* total_out;    // final s
  int skip_nan;        // NaN elements are skipped (== add +0), particle_filter.cpp:112
  int mode;            // 0: addend = x (float).  1: addend = (double)(x - param)^2 for non-NaN x < param, else 0
                       //    (lower-half squared deviations, particle_filter.cpp:120-125; float accumulator, double adds)
  const float* param;  // device scalar (mode 1: the mean, scal[SC_MEAN])
  // the whole chain is skipped (every kernel returns at once) when *skip_if == skip_val: the lower-half deviation is
  // only ever used to replace NaN weights, so an update whose weights are all valid does not compute it
  const unsigned long long* skip_if; unsigned long long skip_val;
};
__device__ __forceinline__ bool seq_skipped(const SeqJob& job) { return job.skip_if && *job.skip_if == job.skip_val; }
#define TDR_MAX_JOBS 9
struct SeqJobs { SeqJob j[TDR_MAX_JOBS]; };

static const int SEQ_THREADS = 1024;
static const int SEQ_ITEMS = 4;
static const int SEQ_CHUNK = SEQ_THREADS * SEQ_ITEMS;

// addend j of a chain as a double (exact for float chains)
__device__ __forceinline__ double seq_elem(const SeqJob& job, long long j) {
  float w = job.x[job.start + j * job.stride];
  if (job.mode == 0) {
    if (job.skip_nan && w != w) w = 0.f;
    return (double)w;
  }
  const float mean = __ldg(job.param);
  if (w == w && w < mean) { float dv = TDR_FSUB(w, mean); return (double)dv * (double)dv; }
  return 0.0;
}
__device__ __forceinline__ float seq_add(float S, double d) { return (float)((double)S + d); }
// the IncPair of one addend.  Mode 0 chains (plain float weights, `strict` false) hold float values in their double
// slots: the 32-bit form gives the same pair (tdr_math.cuh: "for a float-valued d this reduces to inc_pair") at a third
// of the instructions of the 64-bit one — k_seq_tile_aggs is the second largest cost of a normalisation.
__device__ __forceinline__ IncPair elem_pair(double x, int E, bool double_addends, bool* irregular) {
  return double_addends ? inc_pair_d(x, E, true, irregular) : inc_pair((float)x, E, irregular);
}

__device__ __forceinline__ IncPair shfl_up_pair(IncPair v, int d) {
  IncPair r;
  r.a = __shfl_up_sync(0xffffffffu, v.a, d);
  r.b = __shfl_up_sync(0xffffffffu, v.b, d);
  return r;
}

// One CTA walks a chain chunk by chunk (the order-exact emulation described at the top of this file).
// elem(j) returns addend j as a double; strict: see inc_pair_d.  runmax_out (may alias the storage elem reads
// from: every element is read before its slot is written) receives max_{k<=j} s_k.  Returns the final sum to
// every thread.  All SEQ_THREADS threads must call it.
struct CtaSeqShared {
  IncPair warp[32];
  int cross;            // first crossing position inside the chunk (relative), or chunk_len
  uint32_t mprev;       // m of the element just before the crossing
  float S, rmax;        // chain state
  long long pos;
  int mode;             // 0 = binade scan, 1 = plain sequential tail (irregular state)
};

// GT = threads cooperating on the chain (SEQ_THREADS: the whole CTA; 128: one of 8 groups running 8 chains side by
// side, each with its own named barrier and CtaSeqShared).  Every thread of the group must call it.
template <int GT> __device__ __forceinline__ void group_sync() {
  if (GT == SEQ_THREADS) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / GT)), "r"(GT) : "memory");
}
static const int SEQ_HEAD = 96;   // elements added one by one by a single thread before the scan machinery starts:
                                  // the running sum doubles (changes binade) at elements 1, 2, 4, ... so the head
                                  // absorbs the first ~7 binade crossings, each of which would cost a whole chunk pass

template <int GT, class Elem>
__device__ float cta_exact_chain(Elem elem, long long count, bool strict, float* runmax_out, CtaSeqShared& sh) {
  constexpr int CHUNK = GT * SEQ_ITEMS;
  const int tid = threadIdx.x % GT, lane = tid & 31, warp = tid >> 5;
  group_sync<GT>();
  if (tid == 0) {
    float S = 0.f, rm = -INFINITY;
    const long long head = count < SEQ_HEAD ? count : SEQ_HEAD;
    for (long long j = 0; j < head; j++) {
      S = seq_add(S, elem(j));
      if (S > rm) rm = S;
      if (runmax_out) runmax_out[j] = rm;
    }
    sh.S = S; sh.rmax = rm; sh.pos = head;
    sh.mode = (!(S >= 0.f) || S == INFINITY) ? 1 : 0;
  }
  group_sync<GT>();
  while (true) {
    const long long pos = sh.pos;
    if (pos >= count) break;
    if (sh.mode == 1) {
      // irregular accumulator (negative / inf / NaN): finish with real adds, one thread.
      if (tid == 0) {
        float S = sh.S, rm = sh.rmax;
        for (long long j = pos; j < count; j++) {
          S = seq_add(S, elem(j));
          if (S > rm) rm = S;
          if (runmax_out) runmax_out[j] = rm;
        }
        sh.S = S; sh.rmax = rm; sh.pos = count;
      }
      group_sync<GT>();
      continue;
    }
    const float S_in = sh.S;
    const float rmax_in = sh.rmax;
    const int E = binade_of(S_in);
    const uint32_t m_in = mant_of(S_in);
    const uint32_t limit = binade_limit(E);
    long long remaining = count - pos;
    const int chunk_len = remaining < CHUNK ? (int)remaining : CHUNK;
    if (tid == 0) sh.cross = chunk_len;

    // ---- per-thread pairs (blocked: thread owns SEQ_ITEMS consecutive elements)
    IncPair loc[SEQ_ITEMS];
    IncPair agg; agg.a = agg.b = 0;
#pragma unroll
    for (int k = 0; k < SEQ_ITEMS; k++) {
      int j = tid * SEQ_ITEMS + k;
      double w = 0.0;
      if (j < chunk_len) w = elem(pos + j);
      bool irr;
      IncPair pr = elem_pair(w, E, strict, &irr);
      agg = pair_compose(agg, pr);
      loc[k] = agg;                       // inclusive within the thread
    }
    // ---- group exclusive scan of thread aggregates (pair_compose is associative, not commutative)
    IncPair incl = agg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      IncPair o = shfl_up_pair(incl, d);
      if (lane >= d) incl = pair_compose(o, incl);
    }
    if (lane == 31) sh.warp[warp] = incl;
    group_sync<GT>();
    if (warp == 0) {
      IncPair v; v.a = v.b = 0;
      if (lane < GT / 32) v = sh.warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        IncPair o = shfl_up_pair(v, d);
        if (lane >= d) v = pair_compose(o, v);
      }
      if (lane < GT / 32) sh.warp[lane] = v;    // inclusive over warps
    }
    group_sync<GT>();
    IncPair excl_thread = shfl_up_pair(incl, 1);
    if (lane == 0) { excl_thread.a = 0; excl_thread.b = 0; }
    if (warp > 0) excl_thread = pair_compose(sh.warp[warp - 1], excl_thread);

    // ---- element values and the first crossing
    const bool odd = (m_in & 1u) != 0;
    uint32_t mv[SEQ_ITEMS];
    int my_cross = chunk_len;
#pragma unroll
    for (int k = 0; k < SEQ_ITEMS; k++) {
      int j = tid * SEQ_ITEMS + k;
      IncPair t = pair_compose(excl_thread, loc[k]);
      uint32_t inc = odd ? t.b : t.a;
      uint32_t m = m_in + inc;            // <= 2^24 + 2^30, no overflow
      mv[k] = m;
      if (j < chunk_len && m >= limit && j < my_cross) my_cross = j;
    }
    if (my_cross < chunk_len) atomicMin(&sh.cross, my_cross);
    group_sync<GT>();
    const int cross = sh.cross;
    // the crossing element is read BEFORE anything is emitted (runmax_out may alias the element storage)
    double w_cross = 0.0;
    if (tid == 0 && cross < chunk_len) w_cross = elem(pos + cross);
    // ---- emit the valid part [0, cross)
#pragma unroll
    for (int k = 0; k < SEQ_ITEMS; k++) {
      int j = tid * SEQ_ITEMS + k;
      if (j < cross) {
        if (runmax_out) {
          float S = from_binade(E, mv[k]);
          runmax_out[pos + j] = S > rmax_in ? S : rmax_in;   // non-decreasing inside a segment
        }
        if (j == cross - 1) sh.mprev = mv[k];
      }
    }
    group_sync<GT>();
    // ---- advance the chain state (one thread; the crossing add is a real fp32 add)
    if (tid == 0) {
      float S = (cross > 0) ? from_binade(E, sh.mprev) : S_in;
      float rm = rmax_in;
      if (cross > 0 && S > rm) rm = S;
      long long np = pos + cross;
      if (cross < chunk_len) {
        S = seq_add(S, w_cross);
        if (S > rm) rm = S;
        if (runmax_out) runmax_out[pos + cross] = rm;
        np += 1;
        if (!(S >= 0.f) || S == INFINITY) sh.mode = 1;   // negative / NaN / inf accumulator
      }
      sh.S = S; sh.rmax = rm; sh.pos = np;
    }
    group_sync<GT>();
  }
  const float total = sh.S;
  group_sync<GT>();
  return total;
}

// grid.x = number of jobs.  only_flagged != nullptr: run a job only when its flag is set (fallback of the tiled
// path for chains with negative / NaN / inf elements)
__global__ void __launch_bounds__(SEQ_THREADS) k_exact_seq(SeqJobs jobs, const int* __restrict__ only_flagged) {
  const SeqJob job = jobs.j[blockIdx.x];
  if (seq_skipped(job)) return;
  if (only_flagged && only_flagged[blockIdx.x] == 0) return;
  __shared__ CtaSeqShared sh;
  const float total = cta_exact_chain<SEQ_THREADS>([&](long long j) { return seq_elem(job, j); }, job.count, job.mode != 0,
                                      job.runmax_out, sh);
  if (threadIdx.x == 0 && job.total_out) *job.total_out = total;
}

// ================================================================================================
// Tiled (multi-CTA) order-exact accumulation for long chains.
//
// The chain is cut into tiles of SQ_TILE elements.  (0) per-tile double sums give an approximate prefix, hence
// for every tile a small window of binades the exact running sum can be in; (1) every tile reduces its
// elements to one IncPair per candidate binade (grid-parallel); (2) ONE CTA per chain walks the tiles: a block
// scan over the tile pairs of the current binade finds the tile in which the sum leaves the binade (or whose
// window does not contain it), that tile alone is resolved element-wise, and the walk continues in the new
// binade — ~25 iterations for 1e6 normalised weights instead of n/4096 dependent chunk steps; (3) with the
// exact sum at every tile start known, all tiles emit their prefix values in parallel.  The window only affects
// speed: a tile outside its window is resolved element-wise, so the result is exact regardless.
// Chains with negative / NaN / inf elements fall back to k_exact_seq (flag set by step 0).
// ================================================================================================
static const int SQ_THREADS = 512;
static const int SQ_ITEMS = 4;
static const int SQ_TILE = SQ_THREADS * SQ_ITEMS;     // 2048
static const int SQ_HEAD = 128;         // elements at the start of a chain that one thread adds sequentially
static const int SQ_KMAX = 4;                         // candidate binades per tile

struct SeqWs {            // per-job workspace (device pointers)
  double* tile_sum;       // [n_tiles]   approximate sums, then exclusive approximate prefix
  int* win_lo;            // [n_tiles]   first candidate binade
  IncPair* agg;           // [n_tiles][SQ_KMAX]
  float* tile_start;      // [n_tiles]   exact running sum at the tile start
  int* flag;              // [1] irregular elements seen
};
struct SeqWsAll { SeqWs w[TDR_MAX_JOBS]; };


// ordered block reduction / scan of IncPairs (SQ_THREADS threads).  Returns the inclusive scan value of this
// thread; *total = composition of all threads.
__device__ __forceinline__ IncPair block_scan_pairs(IncPair v, IncPair* s_warp, IncPair* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  IncPair incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    IncPair o = shfl_up_pair(incl, d);
    if (lane >= d) incl = pair_compose(o, incl);
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    IncPair w; w.a = w.b = 0;
    if (lane < SQ_THREADS / 32) w = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      IncPair o = shfl_up_pair(w, d);
      if (lane >= d) w = pair_compose(o, w);
    }
    if (lane < SQ_THREADS / 32) s_warp[lane] = w;
  }
  __syncthreads();
  if (warp > 0) incl = pair_compose(s_warp[warp - 1], incl);
  if (total) *total = s_warp[SQ_THREADS / 32 - 1];
  __syncthreads();
  return incl;
}

// (0) approximate tile sums + irregular flag
__global__ void __launch_bounds__(SQ_THREADS) k_seq_tile_sums(SeqJobs jobs, SeqWsAll ws) {
  const SeqJob job = jobs.j[blockIdx.y];
  if (seq_skipped(job)) return;
  const long long base = (long long)blockIdx.x * SQ_TILE;
  if (base >= job.count) return;
  __shared__ double s_d[SQ_THREADS / 32];
  double acc = 0.0; bool irr = false;
#pragma unroll
  for (int k = 0; k < SQ_ITEMS; k++) {
    long long j = base + threadIdx.x * SQ_ITEMS + k;
    if (j < job.count) {
      double w = seq_elem(job, j);
      if (!(w >= 0.0) || w > 1e30) irr = true;              // negative / NaN / inf / huge: sequential fallback
      acc += w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_d[threadIdx.x >> 5] = acc;
  if (irr) atomicOr(ws.w[blockIdx.y].flag, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < SQ_THREADS / 32; w++) t += s_d[w];
    ws.w[blockIdx.y].tile_sum[blockIdx.x] = t;
  }
}

// (0b) one CTA per job: exclusive scan of the tile sums (double) -> candidate binade window per tile
__global__ void __launch_bounds__(SQ_THREADS) k_seq_plan(SeqJobs jobs, SeqWsAll ws) {
  const SeqJob job = jobs.j[blockIdx.x];
  if (seq_skipped(job)) return;
  SeqWs w = ws.w[blockIdx.x];
  if (*w.flag) return;
  const int n_tiles = (int)((job.count + SQ_TILE - 1) / SQ_TILE);
  __shared__ double s_w[SQ_THREADS / 32];
  __shared__ double s_carry;
  if (threadIdx.x == 0) s_carry = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += SQ_THREADS) {
    const int t = base + threadIdx.x;
    const double v = t < n_tiles ? w.tile_sum[t] : 0.0;
    double inc = v;
    for (int d = 1; d < 32; d <<= 1) { double o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    double pre = s_carry;
    for (int k = 0; k < warp; k++) pre += s_w[k];
    const double start = pre + inc - v, end = pre + inc;
    if (t < n_tiles) {
      // the exact fp32 running sum drifts from the real sum by random rounding (~sqrt(n) 2^-24 relative) or, for long
      // runs of near-identical addends a few ulps large, systematically by a few per cent (8e6 equal weights: 4.6 %).
      // +-10 % covers both and still leaves 7 tiles in 10 with ONE candidate binade (the earlier 0.7 / 1.4 always
      // spanned two: k_seq_tile_aggs, the second-largest cost of a normalisation, did twice the work).  A wrong
      // window is only slower, never wrong: the walk resolves such a tile element-wise.
      int e_lo = binade_of((float)(start * 0.9));
      int e_hi = binade_of((float)(end * 1.1));
      if (start <= 0.0) e_lo = -127;
      if (e_hi - e_lo + 1 > SQ_KMAX) e_lo = 1000;          // too many binades inside the tile: resolve element-wise
      int cnt = e_hi - e_lo + 1;                            // candidates actually needed (usually 2 of SQ_KMAX)
      if (cnt < 1 || cnt > SQ_KMAX || start <= 0.0) cnt = SQ_KMAX;
      w.win_lo[t] = (e_lo + 200) | (cnt << 16);
    }
    __syncthreads();
    if (threadIdx.x == SQ_THREADS - 1) s_carry = end;
    __syncthreads();
  }
}

// (1) per tile: one IncPair per candidate binade
__global__ void __launch_bounds__(SQ_THREADS) k_seq_tile_aggs(SeqJobs jobs, SeqWsAll ws) {
  const SeqJob job = jobs.j[blockIdx.y];
  if (seq_skipped(job)) return;
  SeqWs w = ws.w[blockIdx.y];
  const long long base = (long long)blockIdx.x * SQ_TILE;
  if (base >= job.count || *w.flag) return;
  const int packed = w.win_lo[blockIdx.x];
  const int e_lo = (packed & 0xffff) - 200, cnt = packed >> 16;
  if (e_lo > 500) return;
  __shared__ IncPair s_warp[SQ_THREADS / 32];
  double x[SQ_ITEMS];
#pragma unroll
  for (int k = 0; k < SQ_ITEMS; k++) {
    long long j = base + threadIdx.x * SQ_ITEMS + k;
    x[k] = j < job.count ? seq_elem(job, j) : 0.0;
  }
  for (int c = 0; c < SQ_KMAX; c++) {
    const int E = e_lo + c;
    if (c >= cnt) {                                        // outside the planned window: the walk resolves element-wise
      if (threadIdx.x == 0) { IncPair sat; sat.a = sat.b = TDR_INC_SAT; w.agg[(size_t)blockIdx.x * SQ_KMAX + c] = sat; }
      continue;
    }
    IncPair agg; agg.a = agg.b = 0;
    if (E <= 127) {
#pragma unroll
      for (int k = 0; k < SQ_ITEMS; k++) { bool irr; agg = pair_compose(agg, elem_pair(x[k], E, job.mode != 0, &irr)); }
    }
    IncPair total;
    block_scan_pairs(agg, s_warp, &total);
    if (threadIdx.x == 0) w.agg[(size_t)blockIdx.x * SQ_KMAX + c] = total;
  }
}

// element-wise exact processing of ONE tile from the exact running sum S at its start (any number of binade
// crossings inside).  Elements live in x[] (thread owns SQ_ITEMS consecutive ones, tile_len valid in total).
// Optionally emits the prefix values.  Returns the exact running sum at the tile end (all threads).
__device__ float seq_resolve_tile(const double (&x)[SQ_ITEMS], bool strict, int tile_len, float S, float* __restrict__ out,
                                  IncPair* s_warp, int* s_cross, uint32_t* s_mprev, float* s_S) {
  int first = 0;                       // elements [0, first) are already consumed
  if (S == 0.f) {
    // the start of a chain: the sum climbs a binade every element or two, and every crossing costs a round of the
    // block-wide loop below (~2.5 us).  One thread adds the first SQ_HEAD elements one by one instead (~7 crossings).
    __shared__ double s_head[SQ_HEAD];
    const int H = tile_len < SQ_HEAD ? tile_len : SQ_HEAD;
#pragma unroll
    for (int k = 0; k < SQ_ITEMS; k++) {
      const int j = threadIdx.x * SQ_ITEMS + k;
      if (j < H) s_head[j] = x[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int j = 0; j < H; j++) { s = seq_add(s, s_head[j]); if (out) out[j] = s; }
      *s_S = s;
    }
    __syncthreads();
    S = *s_S; first = H;
    __syncthreads();
    if (H >= tile_len) return S;
  }
  while (true) {
    const int E = binade_of(S);
    const uint32_t m_in = mant_of(S), limit = binade_limit(E);
    const bool odd = (m_in & 1u) != 0;
    if (threadIdx.x == 0) *s_cross = tile_len;
    IncPair loc[SQ_ITEMS];
    IncPair agg; agg.a = agg.b = 0;
#pragma unroll
    for (int k = 0; k < SQ_ITEMS; k++) {
      const int j = threadIdx.x * SQ_ITEMS + k;
      IncPair pr; pr.a = pr.b = 0;
      if (j >= first && j < tile_len) { bool irr; pr = elem_pair(x[k], E, strict, &irr); }
      agg = pair_compose(agg, pr);
      loc[k] = agg;
    }
    IncPair incl = block_scan_pairs(agg, s_warp, nullptr);      // contains a __syncthreads after s_cross init
    // exclusive value of this thread = inclusive of the previous thread
    IncPair excl = shfl_up_pair(incl, 1);
    if ((threadIdx.x & 31) == 0) {
      if (threadIdx.x == 0) { excl.a = 0; excl.b = 0; }
      else excl = s_warp[(threadIdx.x >> 5) - 1];               // s_warp holds inclusive warp aggregates
    }
    uint32_t mv[SQ_ITEMS];
    int my_cross = tile_len;
#pragma unroll
    for (int k = 0; k < SQ_ITEMS; k++) {
      const int j = threadIdx.x * SQ_ITEMS + k;
      IncPair t = pair_compose(excl, loc[k]);
      mv[k] = m_in + (odd ? t.b : t.a);
      if (j >= first && j < tile_len && mv[k] >= limit && j < my_cross) my_cross = j;
    }
    if (my_cross < tile_len) atomicMin(s_cross, my_cross);
    __syncthreads();
    const int cross = *s_cross;
#pragma unroll
    for (int k = 0; k < SQ_ITEMS; k++) {
      const int j = threadIdx.x * SQ_ITEMS + k;
      if (j >= first && j < cross) {
        if (out) out[j] = from_binade(E, mv[k]);
        if (j == cross - 1) *s_mprev = mv[k];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float Sn = (cross > first) ? from_binade(E, *s_mprev) : S;
      if (cross < tile_len) {
        // the crossing add is a real fp32 add; x of element `cross` is fetched through shared memory below
        *s_S = Sn;
      } else *s_S = Sn;
    }
    __syncthreads();
    if (cross >= tile_len) return *s_S;
    // the thread owning element `cross` performs the real add and publishes it
    if (threadIdx.x == cross / SQ_ITEMS) {
      double wv = 0.0;
#pragma unroll
      for (int k = 0; k < SQ_ITEMS; k++) if (k == cross % SQ_ITEMS) wv = x[k];
      float Sn = seq_add(*s_S, wv);
      if (out) out[cross] = Sn;
      *s_S = Sn;
    }
    __syncthreads();
    S = *s_S;
    first = cross + 1;
    __syncthreads();
    if (first >= tile_len) return S;
  }
}

// (2) one CTA per job walks the tiles
__global__ void __launch_bounds__(SQ_THREADS) k_seq_walk(SeqJobs jobs, SeqWsAll ws) {
  const SeqJob job = jobs.j[blockIdx.x];
  if (seq_skipped(job)) return;
  SeqWs w = ws.w[blockIdx.x];
  if (*w.flag) return;
  const int n_tiles = (int)((job.count + SQ_TILE - 1) / SQ_TILE);
  __shared__ IncPair s_warp[SQ_THREADS / 32];
  __shared__ int s_cross, s_first_bad;
  __shared__ uint32_t s_mprev;
  __shared__ float s_S;
  __shared__ uint32_t s_mlast;
  float S = 0.f;
  int t0 = 0;
  while (t0 < n_tiles) {
    const int E = binade_of(S);
    const uint32_t m_in = mant_of(S), limit = binade_limit(E);
    const bool odd = (m_in & 1u) != 0;
    const int t = t0 + threadIdx.x;
    IncPair pr; pr.a = pr.b = 0;
    if (t < n_tiles) {
      const int e_lo = (w.win_lo[t] & 0xffff) - 200;
      if (E >= e_lo && E < e_lo + SQ_KMAX) pr = w.agg[(size_t)t * SQ_KMAX + (E - e_lo)];
      else pr.a = pr.b = TDR_INC_SAT;                       // outside the window: resolve this tile element-wise
    }
    if (threadIdx.x == 0) s_first_bad = SQ_THREADS;
    IncPair incl = block_scan_pairs(pr, s_warp, nullptr);
    IncPair excl = shfl_up_pair(incl, 1);
    if ((threadIdx.x & 31) == 0) {
      if (threadIdx.x == 0) { excl.a = 0; excl.b = 0; }
      else excl = s_warp[(threadIdx.x >> 5) - 1];
    }
    const uint32_t m_start = m_in + (odd ? excl.b : excl.a);   // running mantissa at the start of tile t
    const uint32_t m_end = m_in + (odd ? incl.b : incl.a);
    const bool in_range = t < n_tiles;
    const bool bad = in_range && m_end >= limit;               // the sum leaves the binade inside (or before) tile t
    if (bad) atomicMin(&s_first_bad, (int)threadIdx.x);
    __syncthreads();
    const int fb = s_first_bad;
    // tiles before the first bad one (and the bad one itself) start inside binade E
    if (in_range && (int)threadIdx.x <= fb) w.tile_start[t] = from_binade(E, m_start);
    if ((int)threadIdx.x == (fb < SQ_THREADS ? fb : SQ_THREADS - 1)) { s_mlast = (fb < SQ_THREADS) ? m_start : m_end; }
    __syncthreads();
    if (fb >= SQ_THREADS || t0 + fb >= n_tiles) {
      // no crossing in this window of tiles
      const int adv = (n_tiles - t0) < SQ_THREADS ? (n_tiles - t0) : SQ_THREADS;
      // m after the last valid tile: thread adv-1 holds it
      if ((int)threadIdx.x == adv - 1) s_mlast = m_end;
      __syncthreads();
      S = from_binade(E, s_mlast);
      t0 += adv;
      __syncthreads();
      continue;
    }
    // resolve tile t0 + fb element-wise from its exact start
    const int tb = t0 + fb;
    float Sb = from_binade(E, s_mlast);
    const long long base = (long long)tb * SQ_TILE;
    const long long rem = job.count - base;
    const int tile_len = rem < SQ_TILE ? (int)rem : SQ_TILE;
    double x[SQ_ITEMS];
#pragma unroll
    for (int k = 0; k < SQ_ITEMS; k++) {
      const int j = threadIdx.x * SQ_ITEMS + k;
      x[k] = j < tile_len ? seq_elem(job, base + j) : 0.0;
    }
    __syncthreads();
    S = seq_resolve_tile(x, job.mode != 0, tile_len, Sb, nullptr, s_warp, &s_cross, &s_mprev, &s_S);
    t0 = tb + 1;
    __syncthreads();
  }
  if (threadIdx.x == 0 && job.total_out) *job.total_out = S;
}

// (3) all tiles emit their prefix values from their exact start
__global__ void __launch_bounds__(SQ_THREADS) k_seq_emit(SeqJobs jobs, SeqWsAll ws) {
  const SeqJob job = jobs.j[blockIdx.y];
  if (seq_skipped(job)) return;
  SeqWs w = ws.w[blockIdx.y];
  const long long base = (long long)blockIdx.x * SQ_TILE;
  if (base >= job.count || *w.flag || !job.runmax_out) return;
  __shared__ IncPair s_warp[SQ_THREADS / 32];
  __shared__ int s_cross;
  __shared__ uint32_t s_mprev;
  __shared__ float s_S;
  const long long rem = job.count - base;
  const int tile_len = rem < SQ_TILE ? (int)rem : SQ_TILE;
  double x[SQ_ITEMS];
#pragma unroll
  for (int k = 0; k < SQ_ITEMS; k++) {
    const int j = threadIdx.x * SQ_ITEMS + k;
    x[k] = j < tile_len ? seq_elem(job, base + j) : 0.0;
  }
  seq_resolve_tile(x, job.mode != 0, tile_len, w.tile_start[blockIdx.x], job.runmax_out + base, s_warp, &s_cross, &s_mprev, &s_S);
}

// ------------------------------------------------------------------------------------------------
// count of non-NaN weights
__global__ void k_count_valid(const float* __restrict__ w, long long n, unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += (w[i] == w[i]) ? 1ull : 0ull;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void k_mean(float* scal, const unsigned long long* nvalid) {
  // mean = sum / num_valid   (int -> float conversion of num_valid, particle_filter.cpp:117)
  scal[SC_NVALID] = (float)(long long)*nvalid;
  scal[SC_MEAN] = TDR_FDIV(scal[SC_SUM], (float)(int)*nvalid);
}

// number of weights below the mean (particle_filter.cpp:120-125); the squared deviations themselves are an
// order-exact chain of mode 1 (float accumulator, double addends)
__global__ void k_under(const float* __restrict__ w, long long n, const float* __restrict__ scal,
                        unsigned long long* nunder) {
  const float mean = scal[SC_MEAN];
  unsigned long long c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = w[i];
    if (v == v && v < mean) c++;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(nunder, c);
}

// u64: [1] = num_valid, [2] = num_under, [4] <- "k_fill_nan will change a weight" (a NaN to replace, or the all-ones
// fallback): the sums over the filled weights that normalize() started speculatively on the raw ones are redone
__global__ void k_stats(float* scal, unsigned long long* u64, unsigned long long n) {
  int nu = (int)u64[2];
  float bs = scal[SC_BSRAW];
  bs = TDR_FSQRT(TDR_FDIV(bs, (float)nu));                       // :126
  scal[SC_BS] = bs; scal[SC_NUNDER] = (float)nu;
  bool fallback = (scal[SC_SUM] == 0.f) || (nu < 1);             // :129
  scal[SC_FALLBACK] = fallback ? 1.f : 0.f;
  scal[SC_REP] = TDR_FSUB(scal[SC_MEAN], bs);                    // :133
  u64[4] = (fallback || u64[1] != n) ? 1ull : 0ull;
}

__global__ void k_fill_nan(float* __restrict__ w, long long n, const float* __restrict__ scal) {
  const bool fallback = scal[SC_FALLBACK] != 0.f;
  const float rep = scal[SC_REP];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = w[i];
    if (fallback) w[i] = 1.f;               // :130
    else if (v != v) w[i] = rep;            // :133
  }
}

// Eigen's vectorised weights_.sum() (SSE2 packets, 2x unrolled; Eigen/src/Core/Redux.h): combine the 8
// exact chain totals in Eigen's order.  Small n is summed directly.
__device__ float eigen_sum_finish_dev(const float* x, long long n, const float* chain) {
  const long long ps = 4;
  const long long aS2 = (n / (2 * ps)) * (2 * ps), aS = (n / ps) * ps;
  float res;
  if (n == 0) return 0.f;
  if (aS == 0) { res = x[0]; for (long long i = 1; i < n; i++) res = TDR_FADD(res, x[i]); return res; }
  float p0[4], p1[4];
  if (aS2 >= 2 * ps) {
    for (int k = 0; k < 4; k++) { p0[k] = chain[k]; p1[k] = chain[4 + k]; }
    for (int k = 0; k < 4; k++) p0[k] = TDR_FADD(p0[k], p1[k]);
    if (aS > aS2) for (int k = 0; k < 4; k++) p0[k] = TDR_FADD(p0[k], x[aS2 + k]);
  } else {
    for (int k = 0; k < 4; k++) p0[k] = x[k];      // exactly one packet (4 <= n < 8)
  }
  res = TDR_FADD(TDR_FADD(p0[0], p0[2]), TDR_FADD(p0[1], p0[3]));
  for (long long i = aS; i < n; i++) res = TDR_FADD(res, x[i]);
  return res;
}
__global__ void k_eigen_sum_finish(const float* __restrict__ x, long long n, const float* __restrict__ chain,
                                   float* __restrict__ out) {
  *out = eigen_sum_finish_dev(x, n, chain);
}

// w /= s1 ; w = d*w + (1-d)/N   (:135-141)
__global__ void k_regularize(float* __restrict__ w, const float* __restrict__ last_dist, long long n,
                             const float* __restrict__ scal) {
  const float s1 = scal[SC_S1];
  const float fn = (float)(unsigned long long)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = TDR_FDIV(w[i], s1);
    float d = TDR_FMUL(last_dist[i], 5.0f);
    d = (1.0f < d) ? 1.0f : d;               // std::min<float>(a, 1) == (1 < a) ? 1 : a  (a NaN stays NaN)
    w[i] = TDR_FADD(TDR_FMUL(d, v), TDR_FDIV(TDR_FSUB(1.0f, d), fn));
  }
}

// w /= s2 and first arg-max (:142-147)
__global__ void k_final_div_argmax(float* __restrict__ w, long long n, const float* __restrict__ scal,
                                   unsigned long long* __restrict__ best) {
  const float s2 = scal[SC_S2];
  unsigned long long loc = 0ull;   // key: (ordered float bits << 32) | (0xffffffff - index): max key = max value, then smallest index
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = TDR_FDIV(w[i], s2);
    w[i] = v;
    if (v == v) {
      uint32_t u = __float_as_uint(v);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
      if (key > loc) loc = key;
    }
  }
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, loc, o); if (t > loc) loc = t; }
  if ((threadIdx.x & 31) == 0 && loc) atomicMax(best, loc);
}

__global__ void k_argmax_store(const unsigned long long* best, float* scal) {
  unsigned long long k = *best;
  int idx = k ? (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)) : 0;
  reinterpret_cast<int*>(scal)[SC_ARGMAX] = idx;
}

// ------------------------------------------------------------------------------------------------
// resample (:172-185): idx[i] = first j with runmax[j] > sample_i, else n-1;  new state i = old state idx[i]
struct GatherPtrs {
  const float *ix, *iy, *dx, *dy, *th, *sc, *ld; const uint8_t* hi;
  float *oix, *oiy, *odx, *ody, *oth, *osc, *old; uint8_t* ohi;
};

__global__ void k_resample(const float* __restrict__ runmax, long long n, float u, long long M, long long i0,
                           long long i1, int32_t* __restrict__ idx, GatherPtrs g, long long src_n) {
  const float fM = (float)(int)M;
  for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
    float sample = TDR_FDIV(TDR_FADD((float)(int)i, u), fM);    // (static_cast<float>(i)+shift)/num_particles_
    long long lo = 0, hi = n - 1;                               // first j in [0, n-1] with runmax[j] > sample, default n-1
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (runmax[mid] > sample) hi = mid; else lo = mid + 1;
    }
    idx[i - i0] = (int32_t)lo;
    if (g.ix && lo < src_n) {
      long long o = i - i0;
      g.oix[o] = g.ix[lo]; g.oiy[o] = g.iy[lo]; g.odx[o] = g.dx[lo]; g.ody[o] = g.dy[lo];
      g.oth[o] = g.th[lo]; g.osc[o] = g.sc[lo]; g.ohi[o] = g.hi[lo];
      g.old[o] = g.ld[lo];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Small particle sets (tracking): the whole of particle_filter.cpp:107-187 in ONE single-CTA kernel with the
// weights in shared memory — the same arithmetic as the kernel sequence of normalize() + resample(), without the
// ~20 launches and their dependent round trips (the 10 k-particle update is latency-bound, not bandwidth-bound).
// ------------------------------------------------------------------------------------------------
static const int SMALL_MAX = 32768;

__device__ __forceinline__ float block_bcast(float v, float* slot) {
  __syncthreads();
  if (threadIdx.x == 0) *slot = v;
  __syncthreads();
  return *slot;
}

__global__ void __launch_bounds__(SEQ_THREADS) k_small_update(float* __restrict__ w_g, const float* __restrict__ last_dist,
                                                             int n, float* __restrict__ scal, float u, int M,
                                                             int32_t* __restrict__ idx, GatherPtrs g, int do_resample) {
  extern __shared__ float w_s[];
  __shared__ CtaSeqShared sh;
  __shared__ CtaSeqShared sh8[8];
  __shared__ float s_chain[8];
  __shared__ float s_slot;
  __shared__ int s_cnt[2];
  __shared__ unsigned long long s_key;
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += SEQ_THREADS) w_s[i] = w_g[i];
  if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_key = 0ull; }
  // ---- sum / num_valid / mean (:108-117)
  const float sum = cta_exact_chain<SEQ_THREADS>([&](long long j) { float w = w_s[j]; return (double)(w != w ? 0.f : w); }, n, false, nullptr, sh);
  int c = 0;
  for (int i = tid; i < n; i += SEQ_THREADS) c += (w_s[i] == w_s[i]) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[0], c);
  __syncthreads();
  const int nvalid = s_cnt[0];
  const float mean = TDR_FDIV(sum, (float)nvalid);
  // ---- lower-half deviation (:120-126)
  c = 0;
  for (int i = tid; i < n; i += SEQ_THREADS) { float v = w_s[i]; c += (v == v && v < mean) ? 1 : 0; }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((tid & 31) == 0 && c) atomicAdd(&s_cnt[1], c);
  // only ever used to replace NaN weights (:133): an update whose weights are all valid skips the chain (CTA-uniform)
  float bsraw = 0.f;
  if (nvalid != n) bsraw = cta_exact_chain<SEQ_THREADS>([&](long long j) {
    float w = w_s[j];
    if (w == w && w < mean) { float dv = TDR_FSUB(w, mean); return (double)dv * (double)dv; }
    return 0.0; }, n, true, nullptr, sh);
  __syncthreads();
  const int nunder = s_cnt[1];
  const float bs = TDR_FSQRT(TDR_FDIV(bsraw, (float)nunder));
  const bool fallback = (sum == 0.f) || (nunder < 1);                  // :129
  const float rep = TDR_FSUB(mean, bs);                                // :133
  if (tid == 0) {
    scal[SC_SUM] = sum; scal[SC_NVALID] = (float)nvalid; scal[SC_MEAN] = mean; scal[SC_BS] = bs; scal[SC_BSRAW] = bsraw;
    scal[SC_NUNDER] = (float)nunder; scal[SC_FALLBACK] = fallback ? 1.f : 0.f; scal[SC_REP] = rep;
  }
  for (int i = tid; i < n; i += SEQ_THREADS) {
    float v = w_s[i];
    if (fallback) w_s[i] = 1.f; else if (v != v) w_s[i] = rep;
  }
  // ---- weights_.sum() (Eigen order), regularisation, second sum (:135-142)
  const long long aS2 = ((long long)n / 8) * 8;
  for (int pass = 0; pass < 2; pass++) {
    if (aS2 >= 8) {
      // the 8 interleaved chains of Eigen's packet sum side by side: 128 threads and one named barrier each
      __syncthreads();
      const int k = tid >> 7;
      const float t = cta_exact_chain<128>([&](long long j) { return (double)w_s[k + 8 * j]; }, aS2 / 8, false, nullptr, sh8[k]);
      if ((tid & 127) == 0) s_chain[k] = t;
    }
    __syncthreads();
    float tot = 0.f;
    if (tid == 0) tot = eigen_sum_finish_dev(w_s, n, s_chain);
    tot = block_bcast(tot, &s_slot);
    if (pass == 0) {
      if (tid == 0) scal[SC_S1] = tot;
      const float fn = (float)(unsigned long long)n;
      for (int i = tid; i < n; i += SEQ_THREADS) {
        float v = TDR_FDIV(w_s[i], tot);
        float d = TDR_FMUL(last_dist[i], 5.0f);
        d = (1.0f < d) ? 1.0f : d;
        w_s[i] = TDR_FADD(TDR_FMUL(d, v), TDR_FDIV(TDR_FSUB(1.0f, d), fn));
      }
    } else {
      if (tid == 0) scal[SC_S2] = tot;
      unsigned long long loc = 0ull;
      for (int i = tid; i < n; i += SEQ_THREADS) {
        float v = TDR_FDIV(w_s[i], tot);
        w_s[i] = v; w_g[i] = v;
        if (v == v) {
          uint32_t ub = __float_as_uint(v);
          ub = (ub & 0x80000000u) ? ~ub : (ub | 0x80000000u);
          unsigned long long key = ((unsigned long long)ub << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
          if (key > loc) loc = key;
        }
      }
      for (int o = 16; o > 0; o >>= 1) { unsigned long long t2 = __shfl_xor_sync(0xffffffffu, loc, o); if (t2 > loc) loc = t2; }
      if ((tid & 31) == 0 && loc) atomicMax(&s_key, loc);
    }
    __syncthreads();
  }
  if (tid == 0) {
    const unsigned long long k = s_key;
    reinterpret_cast<int*>(scal)[SC_ARGMAX] = k ? (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)) : 0;
  }
  if (!do_resample) return;
  // ---- prefix in place, then the systematic draw (:172-187)
  cta_exact_chain<SEQ_THREADS>([&](long long j) { return (double)w_s[j]; }, n, false, w_s, sh);
  const float fM = (float)M;
  for (int i = tid; i < M; i += SEQ_THREADS) {
    const float sample = TDR_FDIV(TDR_FADD((float)i, u), fM);
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (w_s[mid] > sample) hi = mid; else lo = mid + 1; }
    idx[i] = lo;
    if (g.ix) {
      g.oix[i] = g.ix[lo]; g.oiy[i] = g.iy[lo]; g.odx[i] = g.dx[lo]; g.ody[i] = g.dy[lo];
      g.oth[i] = g.th[lo]; g.osc[i] = g.sc[lo]; g.ohi[i] = g.hi[lo]; g.old[i] = g.ld[lo];
    }
  }
}

// ------------------------------------------------------------------------------------------------
static const long long SQ_MIN_COUNT = 32768;    // shorter chains: the single-CTA kernel (fewer launches)

static int launch_seq(tdr_ctx* ctx, const SeqJobs& jobs, int njobs) {
  long long maxc = 0;
  for (int k = 0; k < njobs; k++) maxc = jobs.j[k].count > maxc ? jobs.j[k].count : maxc;
  if (maxc < SQ_MIN_COUNT || ctx->seq_impl == 1) {
    k_exact_seq<<<njobs, SEQ_THREADS, 0, ctx->stream>>>(jobs, nullptr);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    return TDR_OK;
  }
  const int n_tiles = (int)((maxc + SQ_TILE - 1) / SQ_TILE);
  // workspace: per job  tile_sum (8) + win_lo (4) + agg (8*KMAX) + tile_start (4) per tile, + one flag word
  const size_t per_job = (size_t)n_tiles * (8 + 4 + 8 * SQ_KMAX + 4) + 256;
  if (int e = ctx->seq_ws.reserve(per_job * njobs + 256)) return e;
  SeqWsAll ws;
  unsigned char* p = ctx->seq_ws.as<unsigned char>();
  TDR_CUDA(cudaMemsetAsync(p, 0, 256, ctx->stream));              // the flags live in the first 256 bytes
  int* flags = reinterpret_cast<int*>(p);
  p += 256;
  for (int k = 0; k < njobs; k++) {
    ws.w[k].flag = flags + k;
    ws.w[k].tile_sum = reinterpret_cast<double*>(p); p += (size_t)n_tiles * 8;
    ws.w[k].agg = reinterpret_cast<IncPair*>(p); p += (size_t)n_tiles * 8 * SQ_KMAX;
    ws.w[k].win_lo = reinterpret_cast<int*>(p); p += (size_t)n_tiles * 4;
    ws.w[k].tile_start = reinterpret_cast<float*>(p); p += (size_t)n_tiles * 4;
    p += (256 - ((size_t)n_tiles * 8 * (1 + SQ_KMAX) + (size_t)n_tiles * 8) % 256) % 256;
  }
  bool any_out = false;
  for (int k = 0; k < njobs; k++) any_out |= jobs.j[k].runmax_out != nullptr;
  dim3 grid(n_tiles, njobs);
  k_seq_tile_sums<<<grid, SQ_THREADS, 0, ctx->stream>>>(jobs, ws);
  k_seq_plan<<<njobs, SQ_THREADS, 0, ctx->stream>>>(jobs, ws);
  k_seq_tile_aggs<<<grid, SQ_THREADS, 0, ctx->stream>>>(jobs, ws);
  k_seq_walk<<<njobs, SQ_THREADS, 0, ctx->stream>>>(jobs, ws);
  if (any_out) { k_seq_emit<<<grid, SQ_THREADS, 0, ctx->stream>>>(jobs, ws); count_launch(ctx); }
  k_exact_seq<<<njobs, SEQ_THREADS, 0, ctx->stream>>>(jobs, flags);   // only chains flagged irregular
  count_launch(ctx, 5);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

// Eigen linear-vectorised sum of x[0..n) into scal[slot]
// the 8 interleaved chains of Eigen's packet sum as jobs [at, at + 8)
static void eigen_chain_jobs(SeqJobs& jobs, int at, const float* x, long long n, float* scal, int skip_nan) {
  const long long aS2 = (n / 8) * 8;
  for (int k = 0; k < 8; k++) {
    SeqJob& j = jobs.j[at + k];
    j.x = x; j.start = k; j.stride = 8; j.count = aS2 / 8;
    j.runmax_out = nullptr; j.total_out = scal + SC_CHAIN + k; j.skip_nan = skip_nan;
  }
}
// skip_if_zero: the chain totals in scal[SC_CHAIN..] are already those of x (computed speculatively) unless this device
// word is non-zero
static int eigen_sum(tdr_ctx* ctx, const float* x, long long n, int slot, const unsigned long long* skip_if_zero = nullptr) {
  float* scal = ctx->scal.as<float>();
  const long long aS2 = (n / 8) * 8;
  if (aS2 >= 8) {
    SeqJobs jobs; memset(&jobs, 0, sizeof(jobs));
    eigen_chain_jobs(jobs, 0, x, n, scal, 0);
    if (skip_if_zero) for (int k = 0; k < 8; k++) { jobs.j[k].skip_if = skip_if_zero; jobs.j[k].skip_val = 0ull; }
    if (int e = launch_seq(ctx, jobs, 8)) return e;
  }
  k_eigen_sum_finish<<<1, 1, 0, ctx->stream>>>(x, n, scal + SC_CHAIN, scal + slot);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int normalize(tdr_ctx* ctx, bool lazy_stddev) {
  const long long n = ctx->n_weights;
  TDR_REQUIRE(n > 0, TDR_ESTATE, "no weights");
  TDR_REQUIRE(n < (1ll << 31), TDR_EUNSUPPORTED, "too many weights");
  float* w = ctx->weights.as<float>();
  const float* ld = ctx->ld_override ? ctx->ld_override : ctx->part[ctx->cur].last_dist.as<float>();
  if (int e = ctx->scal.reserve(SC_TOTAL * 4)) return e;
  float* scal = ctx->scal.as<float>();
  double* dbl = reinterpret_cast<double*>(scal + SC_DBL);
  unsigned long long* u64 = reinterpret_cast<unsigned long long*>(dbl);
  TDR_CUDA(cudaMemsetAsync(dbl, 0, 8 * 8, ctx->stream));   // [0]=sumsq [1]=nvalid [2]=nunder [3]=argmax key
  const int blocks = (int)((n + 1023) / 1024 < ctx->sm_count * 4 ? (n + 1023) / 1024 : ctx->sm_count * 4);
  // sum / num_valid (:108-116): sequential fp32 over non-NaN weights
  SeqJobs jobs; memset(&jobs, 0, sizeof(jobs));
  jobs.j[0].x = w; jobs.j[0].start = 0; jobs.j[0].stride = 1; jobs.j[0].count = n; jobs.j[0].runmax_out = nullptr;
  jobs.j[0].total_out = scal + SC_SUM; jobs.j[0].skip_nan = 1;
  // In the steady state no weight is NaN and the all-ones fallback (:129) does not fire, so k_fill_nan changes nothing
  // and the sum AFTER it (:135) runs over the same numbers as this one: its 8 Eigen chains ride along in the same
  // launches (their walks in parallel CTAs) instead of costing a second dependent pass.  k_stats decides on the device
  // whether that was right (u64[4]); if not, eigen_sum below recomputes them.
  const bool spec = n >= 8 && n >= SQ_MIN_COUNT;
  if (spec) eigen_chain_jobs(jobs, 1, w, n, scal, 1);
  if (int e = launch_seq(ctx, jobs, spec ? 9 : 1)) return e;
  if (spec) memset(&jobs.j[1], 0, 8 * sizeof(SeqJob));
  k_count_valid<<<blocks, 256, 0, ctx->stream>>>(w, n, u64 + 1);
  k_mean<<<1, 1, 0, ctx->stream>>>(scal, u64 + 1);
  k_under<<<blocks, 256, 0, ctx->stream>>>(w, n, scal, u64 + 2);
  // bottom_stddev (:120-125): float accumulator, double addends, sequential -> order-exact chain of mode 1
  jobs.j[0].total_out = scal + SC_BSRAW; jobs.j[0].skip_nan = 0; jobs.j[0].mode = 1; jobs.j[0].param = scal + SC_MEAN;
  if (lazy_stddev) { jobs.j[0].skip_if = u64 + 1; jobs.j[0].skip_val = (unsigned long long)n; }     // num_valid == n: no NaN to replace
  if (int e = launch_seq(ctx, jobs, 1)) return e;
  jobs.j[0].skip_if = nullptr;
  k_stats<<<1, 1, 0, ctx->stream>>>(scal, u64, (unsigned long long)n);
  k_fill_nan<<<blocks, 256, 0, ctx->stream>>>(w, n, scal);
  count_launch(ctx, 5);
  if (int e = eigen_sum(ctx, w, n, SC_S1, spec ? u64 + 4 : nullptr)) return e;
  k_regularize<<<blocks, 256, 0, ctx->stream>>>(w, ld, n, scal);
  count_launch(ctx);
  if (int e = eigen_sum(ctx, w, n, SC_S2)) return e;
  k_final_div_argmax<<<blocks, 256, 0, ctx->stream>>>(w, n, scal, u64 + 3);
  k_argmax_store<<<1, 1, 0, ctx->stream>>>(u64 + 3, scal);
  count_launch(ctx, 2);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

// order-exact sequential fp32 totals of up to 8 columns (pose sums, particle_filter.cpp:195-197)
int exact_sums(tdr_ctx* ctx, const float* const* cols, long long n, int ncols, float* totals_dev) {
  TDR_REQUIRE(ncols >= 1 && ncols <= TDR_MAX_JOBS, TDR_EINVAL, "bad column count");
  SeqJobs jobs; memset(&jobs, 0, sizeof(jobs));
  for (int k = 0; k < ncols; k++) {
    jobs.j[k].x = cols[k]; jobs.j[k].start = 0; jobs.j[k].stride = 1; jobs.j[k].count = n;
    jobs.j[k].runmax_out = nullptr; jobs.j[k].total_out = totals_dev + k; jobs.j[k].skip_nan = 0;
  }
  return launch_seq(ctx, jobs, ncols);
}

// prefix (running max of the exact sequential prefix) over the resident weights
int build_prefix(tdr_ctx* ctx) {
  const long long n = ctx->n_weights;
  TDR_REQUIRE(n > 0, TDR_ESTATE, "no weights");
  if (int e = ctx->prefix.reserve((size_t)n * 4)) return e;
  if (int e = ctx->scal.reserve(SC_TOTAL * 4)) return e;
  SeqJobs jobs; memset(&jobs, 0, sizeof(jobs));
  jobs.j[0].x = ctx->weights.as<float>(); jobs.j[0].start = 0; jobs.j[0].stride = 1; jobs.j[0].count = n;
  jobs.j[0].runmax_out = ctx->prefix.as<float>(); jobs.j[0].total_out = ctx->scal.as<float>() + SC_CHAIN + 8;
  jobs.j[0].skip_nan = 0;
  return launch_seq(ctx, jobs, 1);
}

// have_init == 0 over a particle set: warp ballot, one atomic per warp that saw any (none in the steady state)
__global__ void k_count_uninit(const uint8_t* __restrict__ have_init, long long n, int* __restrict__ out) {
  int mine = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mine += have_init[i] ? 0 : 1;
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}
int recount_uninit(tdr_ctx* ctx, Particles& pt) {
  int* d = ctx->uninit_dev.as<int>();
  TDR_CUDA(cudaMemsetAsync(d, 0, 4, ctx->stream));
  const long long n = pt.n;
  const int blocks = (int)((n + 1023) / 1024 < ctx->sm_count * 4 ? (n + 1023) / 1024 : ctx->sm_count * 4);
  k_count_uninit<<<blocks < 1 ? 1 : blocks, 256, 0, ctx->stream>>>(pt.have_init.as<uint8_t>(), n, d);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(ctx->uninit_pin, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaEventRecord(ctx->uninit_ev, ctx->stream));
  ctx->uninit_pending = true;
  return TDR_OK;
}
int sync_uninit(tdr_ctx* ctx) {
  if (!ctx->uninit_pending) return TDR_OK;
  TDR_CUDA(cudaEventSynchronize(ctx->uninit_ev));      // recorded an update ago: normally complete already
  ctx->n_uninit = *ctx->uninit_pin;
  ctx->uninit_pending = false;
  return TDR_OK;
}

// outputs [i0, i1) of the M systematic samples over the resident weights; when src/dst are given the
// states of the drawn particles are gathered from *src into *dst (dst->n = i1 - i0).
// Single GPU: i0 = 0, i1 = M, src = particles_, dst = new_particles_ (particle_filter.cpp:185-187).
int resample(tdr_ctx* ctx, float u, long long M, long long i0, long long i1, Particles* src, Particles* dst) {
  const long long n = ctx->n_weights;
  TDR_REQUIRE(M > 0 && M < (1ll << 31) && i0 >= 0 && i1 <= M && i0 < i1, TDR_EINVAL, "bad resample range");
  if (int e = build_prefix(ctx)) return e;
  const long long cnt = i1 - i0;
  if (int e = ctx->idx.reserve((size_t)cnt * 4)) return e;
  GatherPtrs g; memset(&g, 0, sizeof(g));
  long long src_n = 0;
  if (src && dst) {
    TDR_REQUIRE(src->n > 0, TDR_ESTATE, "no particles to gather");
    if (int e = dst->reserve(cnt)) return e;
    g.ix = src->init_x.as<float>(); g.iy = src->init_y.as<float>(); g.dx = src->dx.as<float>(); g.dy = src->dy.as<float>();
    g.th = src->theta.as<float>(); g.sc = src->scale.as<float>(); g.ld = src->last_dist.as<float>(); g.hi = src->have_init.as<uint8_t>();
    g.oix = dst->init_x.as<float>(); g.oiy = dst->init_y.as<float>(); g.odx = dst->dx.as<float>(); g.ody = dst->dy.as<float>();
    g.oth = dst->theta.as<float>(); g.osc = dst->scale.as<float>(); g.old = dst->last_dist.as<float>(); g.ohi = dst->have_init.as<uint8_t>();
    src_n = src->n;
  }
  int blocks = (int)((cnt + 255) / 256 < ctx->sm_count * 8 ? (cnt + 255) / 256 : ctx->sm_count * 8);
  k_resample<<<blocks, 256, 0, ctx->stream>>>(ctx->prefix.as<float>(), n, u, M, i0, i1, ctx->idx.as<int32_t>(), g, src_n);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (src && dst) {
    dst->n = cnt;
    // un-initialised (gated) particles may have been multiplied or dropped: recount, asynchronously
    if (ctx->uninit_pending || ctx->n_uninit > 0) { if (int e = recount_uninit(ctx, *dst)) return e; }
  }
  return TDR_OK;
}

// a11 + a12 fused for small particle sets; returns *used = false when the set is too large (caller falls back)
int small_update(tdr_ctx* ctx, float u, long long M, bool do_resample, bool* used) {
  *used = false;
  const long long n = ctx->n_weights;
  tdr::Particles& src = ctx->part[ctx->cur];
  tdr::Particles& dst = ctx->part[ctx->cur ^ 1];
  if (n <= 0 || n > SMALL_MAX || ctx->ld_override || src.n != n || ctx->seq_impl == 1) return TDR_OK;
  if (do_resample && (M <= 0 || M >= (1ll << 31))) return TDR_OK;
  if (int e = ctx->scal.reserve(SC_TOTAL * 4)) return e;
  GatherPtrs g; memset(&g, 0, sizeof(g));
  if (do_resample) {
    if (int e = dst.reserve(M)) return e;
    if (int e = ctx->idx.reserve((size_t)M * 4)) return e;
    g.ix = src.init_x.as<float>(); g.iy = src.init_y.as<float>(); g.dx = src.dx.as<float>(); g.dy = src.dy.as<float>();
    g.th = src.theta.as<float>(); g.sc = src.scale.as<float>(); g.ld = src.last_dist.as<float>(); g.hi = src.have_init.as<uint8_t>();
    g.oix = dst.init_x.as<float>(); g.oiy = dst.init_y.as<float>(); g.odx = dst.dx.as<float>(); g.ody = dst.dy.as<float>();
    g.oth = dst.theta.as<float>(); g.osc = dst.scale.as<float>(); g.old = dst.last_dist.as<float>(); g.ohi = dst.have_init.as<uint8_t>();
  }
  TDR_SMEM_OPTIN(ctx, OPTIN_SMALL_UPDATE, k_small_update, SMALL_MAX * 4);
  k_small_update<<<1, SEQ_THREADS, (size_t)n * 4, ctx->stream>>>(ctx->weights.as<float>(), src.last_dist.as<float>(), (int)n,
                                                                ctx->scal.as<float>(), u, (int)M, ctx->idx.as<int32_t>(), g,
                                                                do_resample ? 1 : 0);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (do_resample) {
    dst.n = M;
    if (ctx->uninit_pending || ctx->n_uninit > 0) { if (int e = recount_uninit(ctx, dst)) return e; }
  }
  *used = true;
  return TDR_OK;
}

}  // namespace tdr
