// score_mma_i8.cu — a7/a9/a10, the 40-shift theta search, as an INTEGER tcgen05 gather-GEMM on 16-byte map records.
//
// Reference: TopDownMapPolar::getLocalMap (src/top_down_map_polar.cpp:21-53), StateParticle::getCostForRot
// (src/state_particle.cpp:112-155), StateParticle::computeWeight (:157-219).
//
// Why a second operand format (profiles/r02_gather_microbench.txt, profiles/r02_score_list_wavefront_counters.csv): the
// fp16 hi/lo kernel (score_mma_list.cu) reads one 32-byte record per (hypothesis, cell) and is bound by L1 wavefronts —
// 24.9 per warp load measured, one per distinct 128-byte line inside a half warp (x ~1.2 for partly used lines), at
// most one per clock and SM.  A 128-byte line of that layout holds 4 pixels.  Here a pixel is 16 bytes:
//     bytes 0..6  : hi bytes of v_c = round(w_c * dist_c / q), q = 50 max(w) / 65535   (c < 7)     byte 7 : known (0 / 1)
//     bytes 8..14 : lo bytes of v_c                                                               byte 15: 0
// so a line holds 8 pixels (a 4 x 2 block, or 8 px of a row) and neighbouring hypotheses share lines twice as often; one
// 128-bit load per (hypothesis, cell) — 1.9 records / clk / SM measured for warps whose lanes lie within 8 x 8 px
// (tools/tex_bench.cu), against 1.5 for 256-bit loads of pixel pairs and 1.0 for the 32-byte records.  The kernel was
// ISSUE-bound on its gather threads (78 instructions per record, profiles/r02_SUMMARY.md), so the index arithmetic is
// fixed point (tdr_math.cuh lattice_fixed) on a table pre-multiplied by scale * res * 4096 in constant memory — which
// needs ONE scale for all hypotheses of a launch (FilterParams::fixed_scale; checked on the device).
// FOUR cells make one K = 32 step of tcgen05.mma.kind::i8 (u8 x u8 -> s32, exact), 8 K slots per cell (class 0..6 and
// the known / zero byte).  The hi halves of the four records are one A tile, the lo halves a second one (tensor memory),
// and both meet the SAME scan block B of 96 rows:
//     B rows n = s        : u8 class counts of the scan cell under candidate shift s, slots 0..6     (s < S <= 40)
//            n = 48 + s   : their sum at slot 7
//            n = 88       : 1 at slot 7
//     hi MMA (N = 96, all rows)   : Xhi[s], norm[s] = sum of known x total count, number of known cells
//     lo MMA (N = 48, rows 0..47) : Xlo[s]                                  (rows 40..47 are zero)
//     cost[s] = 0.01 q (256 Xhi[s] + Xlo[s]) / norm[s]
// (The first version kept a class's hi and lo byte next to each other — two cells per K step — and needed separate hi
// and lo count rows: 128 rows x 32 B per TWO cells.  With that, almost half of the bytes reaching an SM were the scan
// operand streamed by cp.async.bulk, and a probe that streamed half of it ran 14 % faster; sharing one block between
// the halves of four cells cut the operand to 768 B per cell: 4.70 -> 4.13 ms, same sums, same bits.)
// Integer sums are exact; the only error is the 16-bit quantisation of w_c * dist_c: |d cost| <= 0.01 q / 2, i.e. a
// relative weight error <= 0.005 q / regularization.  The host takes this path only where that bound is below 9e-6
// (regularization >= 0.42 max(w): the launch file's 0.7 qualifies, the code default 0.15 does not) and the scan
// counts fit a byte (checked on the device; the guarded CUDA-core launch behind the kernel takes over otherwise).
#include "mma_common.cuh"

namespace tdr {

static const int I8_S_MAX = 40;           // candidate shifts per launch
static const int I8_NH = 96;              // rows of the scan block = accumulator columns of the hi MMA: [0, 40) class counts, [48, 88) class-summed counts, 88 the probe
static const int I8_NL = 48;              // accumulator columns of the lo MMA (rows [0, 48) of the same block: counts, then zeros)
static const int I8_NORM0 = 48, I8_PROBE = 88;
static const int I8_ACC = I8_NH + I8_NL;  // accumulator columns per tile
static const int I8_CELLS = 4;            // lattice cells per pipeline stage = per K = 32 step (8 bytes each)
static const int I8_MAX_COUNT = 255;
static const int PLAN_HDR = 4;            // ints in front of the cell lists of a scan plan (k_plan_cells)

__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// unfiltered texel fetch with integer coordinates; outside the texture the border colour (zeros) comes back
__device__ __forceinline__ uint4 tex_fetch(unsigned long long tex, int x, int y) {
  uint4 v;
  asm("tex.2d.v4.u32.s32 {%0, %1, %2, %3}, [%4, {%5, %6}];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(tex), "r"(x), "r"(y));
  return v;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// ------------------------------------------------------------------------------------------------
// the 16-byte map copy
// ------------------------------------------------------------------------------------------------
// (Splitting the record into a hi and a lo plane of 8-byte records, 16 px per line, read with two LDG.64, was measured
// SLOWER: 4.17 ms against 3.87 ms in Morton order, 4.8-5.6 ms in row order — profiles/r02_sweep_i8.txt.)
// Two layouts: row-major (`pitch` records per map row: a 128-byte line = 8 px of one row) or 4 x 2-px blocks (a line =
// 4 px of two adjacent rows, `pitch` blocks per block row) — with hypotheses in Morton order the compact 2-D clusters of
// a warp touch fewer lines of the blocked layout (12.1 against 14.6 simulated on cfg3).  One all-zero record past the
// map is what cells off the map read.
struct Geom8 {
  uint32_t pitch, zero_rec;
  uint32_t lim_y, lim_x;           // 4096 rows - 1, 4096 cols - 1 (lattice_fixed)
};
template <bool BLK> __host__ __device__ __forceinline__ uint32_t map8_index(uint32_t r, uint32_t c, uint32_t pitch) {
  if (!BLK) return r * pitch + c;
  return (((r >> 1) * pitch + (c >> 2)) << 3) | ((r & 1u) << 2) | (c & 3u);
}

static __global__ void k_build_map8(const MapPixel* __restrict__ map, size_t n, int cols, int C, const float* __restrict__ cw,
                                    float inv_q, int blocked, uint32_t pitch, uint4* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* src = reinterpret_cast<const float4*>(map + i);
  const float4 a = src[0], b = src[1];
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    uint32_t q = 0;
    if (k < C) {
      float x = TDR_FMUL(TDR_FMUL(cw[k], v[k]), inv_q);
      x = fminf(fmaxf(x, 0.f), 65535.f);
      q = __float2uint_rn(x);
    }
    // byte k: hi, byte 8 + k: lo
    w[k >> 2] |= (q >> 8) << ((k & 3) * 8);
    w[2 + (k >> 2)] |= (q & 0xffu) << ((k & 3) * 8);
  }
  if (v[7] != 0.f) w[1] |= 1u << 24;            // byte 7: known
  const uint32_t r = (uint32_t)(i / cols), c = (uint32_t)(i % cols);
  out[blocked ? map8_index<true>(r, c, pitch) : i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// one scale for every hypothesis of the launch?  (min, max) of the scale bits over the particles still to be searched
static __global__ void k_scale_range(const float* __restrict__ scale, const uint8_t* __restrict__ have_init, long long n,
                                     uint32_t* __restrict__ range /* [min, max] of the (positive) float bits */, int want_init) {
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if ((have_init[i] != 0) == (want_init != 0)) { const uint32_t b = __float_as_uint(scale[i]); lo = min(lo, b); hi = max(hi, b); }
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(range, lo); atomicMax(range + 1, hi); }
}
// table of lattice offsets for that scale, ((tab * scale) * res) * 4096, in PLAN order: the gathered cells (padding: far
// off every map), then from entry 2 x stages on the skipped ones
static __global__ void k_scale_tab4096(const float2* __restrict__ tab, int P, const int* __restrict__ plan, const uint32_t* __restrict__ range,
                                       float res, float2* __restrict__ out, int* __restrict__ bailed) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int P_cap = (P + 3) & ~3;
  if (j == 0 && range[0] != range[1]) *bailed = 1;           // mixed scales (or negative / NaN bits): not this kernel's case
  const int n_gather = I8_CELLS * plan[2];
  if (j >= n_gather + plan[1]) return;
  const float scale = __uint_as_float(range[0]);
  const int p = j < n_gather ? plan[PLAN_HDR + j] : plan[PLAN_HDR + P_cap + (j - n_gather)];
  out[j] = p >= 0 ? make_float2(TDR_FMUL(TDR_FMUL(TDR_FMUL(tab[p].x, scale), res), 4096.f), TDR_FMUL(TDR_FMUL(TDR_FMUL(tab[p].y, scale), res), 4096.f))
                  : make_float2(1e30f, 1e30f);
}

// Which lattice cells does this scan need?  A lattice cell (theta, r) meets the scan cells (theta + s, r) of the S
// candidate shifts; if none of them holds a single return in any class, the cell contributes nothing to any cost or
// normalisation and is not gathered at all (top_down_map_polar.cpp:21-53 per cell is the dominant cost) — whole rings
// beyond the sensor's range, and most of the sparse outer ones.  Only the "mostly unknown" test
// (state_particle.cpp:117-120) still needs the known flags of the skipped cells: the kernel's epilogue counts them
// exactly, and only for the hypotheses whose answer the gathered cells leave open.
// plan: [0] gathered cells  [1] skipped cells  [2] stages (4 cells each, >= 1)  [3] unused | gathered[P_cap] | skipped[P_cap]
static __global__ void __launch_bounds__(1024) k_plan_cells(const float* __restrict__ img, int C, int n_theta, int n_r,
                                                            const int32_t* __restrict__ shifts, int S, int skip_on, int* __restrict__ plan) {
  __shared__ float s_tot[MMA_TAB_MAX];
  __shared__ int s_w[32];
  __shared__ int s_shift[TDR_MAX_SHIFTS];
  const int P = n_theta * n_r, P_cap = (P + 3) & ~3;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float t = 0.f;
    for (int c = 0; c < C; c++) t += img[(size_t)c * P + p];
    s_tot[p] = t;
  }
  for (int k = threadIdx.x; k < S; k += blockDim.x) { int v = shifts[k] % n_theta; s_shift[k] = v < 0 ? v + n_theta : v; }
  __syncthreads();
  // this thread's cells: 4 consecutive ones (P <= 4096)
  int live[4], n_live = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int p = threadIdx.x * 4 + q;
    live[q] = 0;
    if (p < P) {
      if (!skip_on) live[q] = 1;
      else {
        const int r = p / n_theta, th = p - r * n_theta;
        for (int k = 0; k < S && !live[q]; k++) {
          int t2 = th + s_shift[k]; if (t2 >= n_theta) t2 -= n_theta;
          if (s_tot[r * n_theta + t2] != 0.f) live[q] = 1;              // counts are >= 0: a zero sum means every class is 0
        }
      }
      n_live += live[q];
    }
  }
  int total;
  int pos = block_excl_scan(n_live, s_w, &total);
  int* act = plan + PLAN_HDR;
  int* skp = act + P_cap;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int p = threadIdx.x * 4 + q;
    if (p < P) {
      if (live[q]) act[pos++] = p;
    }
  }
  // skipped cells: the same scan on the complement
  int n_dead = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) { const int p = threadIdx.x * 4 + q; if (p < P && !live[q]) n_dead++; }
  int total_dead;
  int dpos = block_excl_scan(n_dead, s_w, &total_dead);
#pragma unroll
  for (int q = 0; q < 4; q++) { const int p = threadIdx.x * 4 + q; if (p < P && !live[q]) skp[dpos++] = p; }
  for (int j = total + threadIdx.x; j < P_cap; j += blockDim.x) act[j] = -1;            // padding of the last stage
  if (threadIdx.x == 0) { plan[0] = total; plan[1] = total_dead; plan[2] = total > 0 ? (total + I8_CELLS - 1) / I8_CELLS : 1; plan[3] = 0; }
}

// scan operand, per stage k (cells act[4k .. 4k + 3]): [kc = 2][n = I8_NH][16 B] (K-major canonical layout, LBO = I8_NH*16,
// SBO = 128); chunk kc holds cells 2 kc and 2 kc + 1 of the stage, 8 K slots (bytes) each — the slots of a record HALF:
// class 0..6, then the known byte (hi half) / zero (lo half).  Row s < S: the class counts of the scan cell the lattice
// cell meets under candidate shift s (slots 0..6); row 48 + s: their sum at slot 7; row 88: 1 at slot 7.  The hi MMA
// multiplies all 96 rows with the hi halves (-> Xhi, norm, number of known cells), the lo MMA rows [0, 48) with the lo
// halves (-> Xlo; rows 40..47 are zero).  One thread per (stage, n).
static __global__ void k_build_scan_operand_i8(const float* __restrict__ img, int C, int n_theta, int P, const int* __restrict__ plan,
                                               const int32_t* __restrict__ shifts, int S, uint4* __restrict__ out,
                                               int* __restrict__ maxcount) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_stages = plan[2];
  if (id >= (long long)n_stages * I8_NH) return;
  const int k = (int)(id / I8_NH), n = (int)(id - (long long)k * I8_NH);
  const int* act = plan + PLAN_HDR;
  uint4* stage = out + (size_t)k * I8_NH * 2;
  const int kind = n < I8_S_MAX ? 0 : (n >= I8_NORM0 && n < I8_NORM0 + I8_S_MAX) ? 1 : n == I8_PROBE ? 2 : 3;
  const int s = kind == 0 ? n : n - I8_NORM0;
#pragma unroll
  for (int kc = 0; kc < 2; kc++) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int p = act[I8_CELLS * k + 2 * kc + h];
      if (p < 0 || kind == 3 || (kind < 2 && s >= S)) continue;
      if (kind == 2) { w[2 * h + 1] = 1u << 24; continue; }             // probe row: counts the known cells
      const int r = p / n_theta, th = p - r * n_theta;
      int t2 = th + shifts[s];
      t2 %= n_theta; if (t2 < 0) t2 += n_theta;
      const int cell = r * n_theta + t2;                                // scan row (theta + shift) pairs with map row theta
      float tot = 0.f;
      for (int c = 0; c < C; c++) {
        const float v = img[(size_t)c * P + cell];
        tot += v;
        const uint32_t u = (uint32_t)fminf(fmaxf(v, 0.f), 255.f);
        if (kind == 0) w[2 * h + (c >> 2)] |= u << ((c & 3) * 8);       // slot c
      }
      if (kind == 1) {
        w[2 * h + 1] = (uint32_t)fminf(fmaxf(tot, 0.f), 255.f) << 24;   // slot 7
        if (s == 0) atomicMax(maxcount, (int)fminf(tot, 1e9f));         // every non-empty scan cell X is met at shift 0 by the gathered cell X - shift[0]: the max over all counts in play
      }
    }
    stage[kc * I8_NH + n] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

struct I8Params {
  const uint4* map8; Geom8 geom; float resolution;
  const int* plan; int P, P_cap;      // the scan plan (k_plan_cells)
  const float2* tab_g;                // the constant table's global twin (divergent indices: the epilogue's lookups): [0] gathered cells [1] skipped cells [2] stages
  const uint4* bop;
  const int* perm; long long n_work;
  const int* n_work_dev;              // != nullptr: the number of listed hypotheses lives on the device (tracking passes)
  int track, track_lo, n_theta;       // track: every particle keeps the cost of ITS heading's row shift, list position shift - track_lo
  const float *init_x, *init_y, *dx, *dy; float* theta; const float* scale; uint8_t* have_init; float* weights;
  int force_on_map; float map_w, map_h; int scale_gate; double scale_lo, scale_hi; float regularization;
  const float* thetas; int n_shifts;
  unsigned long long tex;           // the same records as a pitch-linear 2-D texture (border = zero record); 0: not used
  float q001;                       // 0.01 * q: accumulator units -> cost
  const int* maxcount; int* bailed; // device-side preconditions: scan counts fit a byte, one scale for all hypotheses
};

// T = 128-hypothesis tiles per CTA, R = gather threads per hypothesis row (they take turns stage by stage),
// F = stages every gather thread keeps in flight (2 cells = two 128-bit loads each)
template <int T, int R, int F> struct I8Cfg {
  static const int kThreads = 128 * T * R + 64;
  static const int kBBytes = I8_NH * 32;                   // scan operand of one stage (4 cells)
  static const int kACols = T * 16;                        // TMEM columns of one stage of A: per tile 8 (hi halves of 4 cells) + 8 (lo halves)
  static const int kCtasPerSm = T == 1 ? 2 : 1;            // one tile per CTA leaves room for two CTAs (two pipelines) per SM
  static const int kTmemCols = 512 / kCtasPerSm;
  static const int kStagesMax = (kTmemCols - T * I8_ACC) / kACols;
  static const int kStages = kStagesMax > 16 ? 16 : kStagesMax;
  static const int kSmem = kStages * kBBytes + 512;        // + barriers (2 * kStages + 1) and the TMEM base word
  static_assert(kStages >= R * F, "every stage a thread has in flight needs its own slot");
};

// TEX: the second cell of every stage comes through the texture pipe instead of the LSU — the kernel is bound by L1
// data-pipe wavefronts of its 128-bit loads (84 % busy, profiles/r02_SUMMARY.md) and the two front ends together deliver
// more records per clock than either alone (2.27 against 1.9, tools/tex_bench.cu)
template <int T, int R, int F, bool TEX, bool BLK>
__global__ void __launch_bounds__(128 * T * R + 64, I8Cfg<T, R, F>::kCtasPerSm) k_score_mma_i8(I8Params sp) {
  using Cfg = I8Cfg<T, R, F>;
  constexpr int GW = 4 * T * R;        // gather warps
  constexpr int NS = Cfg::kStages;
  if (*sp.maxcount > I8_MAX_COUNT || *sp.bailed) {   // grid-uniform, before anything is allocated
    if (blockIdx.x == 0 && threadIdx.x == 0) *sp.bailed = 1;
    return;
  }
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sB = smem;                                                   // [NS][kc 2][I8_NH][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * Cfg::kBBytes);   // full[NS] empty[NS] accum
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NS), bar_accum = smem_u32(bars + 2 * NS);

  if (warp == GW + 1) tmem_alloc(smem_u32(s_tmem), Cfg::kTmemCols);
  if (tid == 0) {
    for (int s = 0; s < NS; s++) { mbar_init(bar_full + 8 * s, 4 * T + 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_accum, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const long long per_batch = 128 * T;
  const long long n_work = sp.n_work_dev ? (long long)*sp.n_work_dev : sp.n_work;
  const long long n_batches = (n_work + per_batch - 1) / per_batch;
  const int K_ITERS = sp.plan[2];
  uint32_t local_batch = 0;

  if (warp < GW) {
    // =========================== gather + epilogue ===========================
    const int sub = warp / (4 * T);                      // which of the R threads of a row this is
    const int t = (warp % (4 * T)) >> 2;
    const uint4* map8 = sp.map8;
    uint32_t it = 0;                                      // pipeline iteration of the batch's first stage (same sequence in every role)
    const uint32_t tcol0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(T * I8_ACC + t * 16);
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, local_batch++) {
      const long long slot = batch * per_batch + (tid % (128 * T));
      long long i = -1;
      if (slot < n_work) i = sp.perm ? (long long)sp.perm[slot] : slot;
      float cx = 0.f, cy = 0.f;
      bool active = false, gated = false;
      if (i >= 0) {
        const float sc = sp.scale[i];
        cx = TDR_FADD(TDR_FMUL(sp.dx[i], sc), sp.init_x[i]);
        cy = TDR_FADD(TDR_FMUL(sp.dy[i], sc), sp.init_y[i]);
        if (sp.force_on_map && (cx < 0.f || cy < 0.f || cx > sp.map_w || cy > sp.map_h)) gated = true;      // :163-168
        if (sp.scale_gate && ((double)sc < sp.scale_lo || (double)sc > sp.scale_hi)) gated = true;          // :169-176
        active = !gated;
      }
      // centre / resolution, times 4096 (exact); a NaN centre or an idle row reads the zero record everywhere
      const float oy = TDR_FMUL(TDR_FDIV(cy, sp.resolution), 4096.f), ox = TDR_FMUL(TDR_FDIV(cx, sp.resolution), 4096.f);
      if (!(oy == oy) || !(ox == ox)) active = false;
      const uint32_t lim_y = active ? sp.geom.lim_y : 0u, lim_x = sp.geom.lim_x;
      // slot / parity of THIS thread's next store: iteration it + sub, then + R per stage
      uint32_t st = (it + (uint32_t)sub) % NS, ph = ((it + (uint32_t)sub) / NS) & 1u;

      // one 128-bit load per cell: the 16-byte record of this hypothesis' lattice pixel (top_down_map_polar.cpp:28-37)
      auto load_stage = [&](int k, uint4 (&rec)[I8_CELLS]) {
#pragma unroll
        for (int g = 0; g < I8_CELLS; g++) {
          const float2 tb = c_tab[I8_CELLS * k + g];       // ((tab * scale) * res) * 4096; padding cells: far off the map
          const uint32_t ty = (uint32_t)__float2int_rz(TDR_FADD(tb.x, oy)) + 2047u;
          const uint32_t tx = (uint32_t)__float2int_rz(TDR_FADD(tb.y, ox)) + 2047u;
          const bool ok = ty < lim_y && tx < lim_x;
          if (TEX && (g & 1)) rec[g] = tex_fetch(sp.tex, ok ? (int)((tx + 1u) >> 12) : -1, (int)((ty + 1u) >> 12));
          else {
            uint32_t r = map8_index<BLK>((ty + 1u) >> 12, (tx + 1u) >> 12, sp.geom.pitch);
            if (!ok) r = sp.geom.zero_rec;
            rec[g] = __ldg(map8 + r);
          }
        }
      };
      auto store_stage = [&](const uint4 (&rec)[I8_CELLS]) {
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        // row m of the A tiles = TMEM lane m (this warp's quarter); 8 columns = 32 K slots = the hi halves (.x .y) of the
        // stage's four records, the next 8 columns their lo halves (.z .w)
        tmem_st8(tcol0 + st * Cfg::kACols, make_uint4(rec[0].x, rec[0].y, rec[1].x, rec[1].y), make_uint4(rec[2].x, rec[2].y, rec[3].x, rec[3].y));
        tmem_st8(tcol0 + st * Cfg::kACols + 8u, make_uint4(rec[0].z, rec[0].w, rec[1].z, rec[1].w), make_uint4(rec[2].z, rec[2].w, rec[3].z, rec[3].w));
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
        st += R; if (st >= NS) { st -= NS; ph ^= 1u; }
      };

      // software pipeline of depth F over this thread's stages sub, sub + R, ...: F - 1 loads ahead of every store
      uint4 buf[F][I8_CELLS];
      const int J = (K_ITERS - sub + R - 1) / R;         // this thread's stages
#pragma unroll
      for (int f = 0; f < F - 1; f++) if (f < J) load_stage(sub + f * R, buf[f]);
#pragma unroll 1
      for (int j = 0; j < J; j += F) {
#pragma unroll
        for (int f = 0; f < F; f++) {
          const int jj = j + f;
          if (jj + F - 1 < J) load_stage(sub + (jj + F - 1) * R, buf[(f + F - 1) % F]);
          if (jj < J) store_stage(buf[f]);
        }
      }
      it += (uint32_t)K_ITERS;
      if (sub != 0) continue;                            // the first thread of each row owns the epilogue

      // ---- epilogue: this thread's accumulator row (s32)
      mbar_wait(bar_accum, local_batch & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * I8_ACC);
      const int S = sp.n_shifts;
      const uint32_t kraw = tmem_ld1(trow + (uint32_t)I8_PROBE);
      tmem_wait_ld();
      // known cells: the gathered ones were counted by the probe row; the cells of the skipped (empty) rings only matter
      // when they can still tip the "less than half known" test, and are looked up one by one then
      int known = (int)kraw;
      const int n_skip = sp.plan[1];
      const bool open_q = i >= 0 && !gated && n_skip > 0 && 2 * known < sp.P && 2 * (known + n_skip) >= sp.P;
      // the warp settles its open rows one after the other, 32 skipped cells at a time (6 % of cfg3's particles)
      unsigned need = __ballot_sync(0xffffffffu, open_q);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const float soy = __shfl_sync(0xffffffffu, oy, src), sox = __shfl_sync(0xffffffffu, ox, src);
        int cnt = 0;
        for (int m = lane; m < n_skip; m += 32) {
          const float2 tb = sp.tab_g[I8_CELLS * K_ITERS + m];
          const uint32_t ty = (uint32_t)__float2int_rz(TDR_FADD(tb.x, soy)) + 2047u;
          const uint32_t tx = (uint32_t)__float2int_rz(TDR_FADD(tb.y, sox)) + 2047u;
          if (ty < sp.geom.lim_y && tx < lim_x) cnt += (int)((__ldg(map8 + map8_index<BLK>((ty + 1u) >> 12, (tx + 1u) >> 12, sp.geom.pitch)).y >> 24) & 1u);
        }
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == src) known += cnt;
      }
      const bool unknown = (double)TDR_FDIV((float)known, (float)sp.P) < 0.5;                    // :117-120
      float best = 3.402823466e+38f, best_theta = 0.f;                                           // :193-204
      // tracking: the one column of this particle's heading (state_particle.cpp:207-210); NaN stays NaN there
      const int my_col = (sp.track && i >= 0) ? rot_to_shift(sp.theta[i], sp.n_theta) - sp.track_lo : -1;
      float my_cost = 0.f;
      uint32_t vh[8], vl[8], vn[8];
#pragma unroll 1
      for (int ch = 0; ch * 8 < S; ch++) {
        tmem_ld8(trow + (uint32_t)(ch * 8), vh);
        tmem_ld8(trow + (uint32_t)(I8_NH + ch * 8), vl);
        tmem_ld8(trow + (uint32_t)(I8_NORM0 + ch * 8), vn);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int s = ch * 8 + j;
          if (s < S) {
            const float num = fmaf((float)(int)vh[j], 256.f, (float)(int)vl[j]);
            const float cost = unknown ? __int_as_float(0x7fc00000) : TDR_FDIV(TDR_FMUL(num, sp.q001), (float)(int)vn[j]);   // :137,154
            if (sp.track) { if (s == my_col) my_cost = cost; }
            else if (cost < best) { best = cost; best_theta = sp.thetas[s]; }
          }
        }
      }
      if (i >= 0 && sp.track) {
        sp.weights[i] = gated ? 0.f : (float)(1.0 / (double)TDR_FADD(my_cost, sp.regularization));   // :212
      } else if (i >= 0) {
        if (gated) sp.weights[i] = 0.f;
        else {
          sp.theta[i] = best_theta;
          sp.have_init[i] = 1;
          sp.weights[i] = (float)(1.0 / (double)TDR_FADD(best, sp.regularization));             // :212
        }
      }
      tc_fence_before();           // TMEM reads are done before the next batch's first full-barrier arrive
    }
  } else if (warp == GW) {
    // =========================== scan-operand loader ===========================
    const bool leader = elect_one();
    const uint32_t sB_u = smem_u32(sB);
    uint32_t st = 0, ph = 0;
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(sp.bop);
      for (int k = 0; k < K_ITERS; k++, src += Cfg::kBBytes) {
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        if (leader) {
          mbar_expect_tx(bar_full + 8 * st, Cfg::kBBytes);
          bulk_g2s(sB_u + st * Cfg::kBBytes, src, Cfg::kBBytes, bar_full + 8 * st);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1u; }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    // instruction descriptors: D = s32, A = B = u8, both K-major, M = 128 (K = 32); N = 96 (hi halves) and 48 (lo halves)
    const bool leader = elect_one();
    const uint32_t idesc_hi = (2u << 4) | ((uint32_t)(I8_NH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_lo = (2u << 4) | ((uint32_t)(I8_NL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sB_u = smem_u32(sB);
    uint32_t st = 0, ph = 0;
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
      for (int k = 0; k < K_ITERS; k++) {
        mbar_wait(bar_full + 8 * st, ph);
        tc_fence_after();
        if (leader) {
          const uint64_t bdesc = umma_desc(sB_u + st * Cfg::kBBytes, I8_NH * 16, 128);    // the lo MMA reads rows [0, 48) of the same block
#pragma unroll
          for (int tt = 0; tt < T; tt++) {
            const uint32_t a_cols = tmem_base + (uint32_t)(T * I8_ACC + tt * 16) + st * Cfg::kACols;
            umma_i8_ts(tmem_base + (uint32_t)(tt * I8_ACC), a_cols, bdesc, idesc_hi, k > 0 ? 1u : 0u);
            umma_i8_ts(tmem_base + (uint32_t)(tt * I8_ACC + I8_NH), a_cols + 8u, bdesc, idesc_lo, k > 0 ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * st);        // implies tcgen05.fence::before_thread_sync
          if (k == K_ITERS - 1) umma_commit(bar_accum);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GW + 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// worst-case relative weight error of the 16-bit records: 0.005 q / regularization, q = 50 max(w) / 65535
static bool i8_usable(tdr_ctx* ctx, int n_shifts, float* q_out) {
  if (ctx->score_impl == 1 || ctx->mma_i8 == 0) return false;
  if (n_shifts < 1 || n_shifts > I8_S_MAX || ctx->C > 7) return false;
  float wmax = 0.f;
  for (int c = 0; c < ctx->C; c++) {
    const float w = ctx->fp.class_weights[c];
    if (!(w >= 0.f) || w > 1e4f) return false;
    if (w > wmax) wmax = w;
  }
  if (!(wmax > 0.f) || !(ctx->fp.regularization > 0.f)) return false;
  const float q = 50.f * wmax / 65535.f;
  *q_out = q;
  if (ctx->mma_i8 == 2) return true;                       // forced (tests of the kernel itself)
  return 0.005 * (double)q / (double)ctx->fp.regularization <= 9e-6;
}

static int build_map8(tdr_ctx* ctx, float q, bool blocked, const uint4** out, Geom8* geom) {
  const uint32_t bx = ((uint32_t)ctx->cols + 3u) / 4u, by = ((uint32_t)ctx->rows + 1u) / 2u;
  const size_t n_px = (size_t)ctx->rows * ctx->cols;
  const size_t n_rec = blocked ? (size_t)bx * by * 8 : n_px;
  TDR_REQUIRE(n_rec < (1ull << 31) && ctx->rows < (1 << 19) && ctx->cols < (1 << 19), TDR_EUNSUPPORTED,
              "map too large for the 16-byte copy (%zu records)", n_rec);
  geom->pitch = blocked ? bx : (uint32_t)ctx->cols; geom->zero_rec = (uint32_t)n_rec;
  geom->lim_y = 4096u * (uint32_t)ctx->rows - 1u; geom->lim_x = 4096u * (uint32_t)ctx->cols - 1u;
  if (!ctx->map8_valid || ctx->map8_q != q || ctx->map8_blocked != (blocked ? 1 : 0)) {
    if (int e = ctx->map8.reserve((n_rec + 1) * 16)) return e;
    if (blocked) TDR_CUDA(cudaMemsetAsync(ctx->map8.p, 0, (n_rec + 1) * 16, ctx->stream));           // padding pixels of the edge blocks
    else TDR_CUDA(cudaMemsetAsync(ctx->map8.as<unsigned char>() + n_rec * 16, 0, 16, ctx->stream));   // the all-zero record
    k_build_map8<<<(unsigned)((n_px + 255) / 256), 256, 0, ctx->stream>>>(ctx->map_px.as<MapPixel>(), n_px, ctx->cols, ctx->C,
                                                                           ctx->d_cw.as<float>(), 1.f / q, blocked ? 1 : 0, geom->pitch, ctx->map8.as<uint4>());
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    ctx->map8_valid = true; ctx->map8_q = q; ctx->map8_blocked = blocked ? 1 : 0;
    // the same buffer as a pitch-linear 2-D texture of uint4 texels; outside it reads the border colour = the zero record
    if (ctx->map8_tex) { cudaDestroyTextureObject((cudaTextureObject_t)ctx->map8_tex); ctx->map8_tex = 0; }
    if (ctx->mma_tex && !blocked && ((size_t)ctx->cols * 16) % 32 == 0) {
      cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
      rd.resType = cudaResourceTypePitch2D;
      rd.res.pitch2D.devPtr = ctx->map8.p; rd.res.pitch2D.desc = cudaCreateChannelDesc<uint4>();
      rd.res.pitch2D.width = (size_t)ctx->cols; rd.res.pitch2D.height = (size_t)ctx->rows; rd.res.pitch2D.pitchInBytes = (size_t)ctx->cols * 16;
      cudaTextureDesc td; memset(&td, 0, sizeof(td));
      td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint;
      td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
      cudaTextureObject_t t = 0;
      if (cudaCreateTextureObject(&t, &rd, &td, nullptr) == cudaSuccess) ctx->map8_tex = (unsigned long long)t;
      else (void)cudaGetLastError();                  // size beyond the linear-texture limits: loads only
    }
  }
  *out = ctx->map8.as<uint4>();
  return TDR_OK;
}

// returns TDR_OK and sets *used = true when the integer tensor-core path was launched (it may still leave the work to
// the guarded CUDA-core launch behind it: scan counts above 255, decided on the device)
// (min, max) of the scale bits over the particles of a launch into the context's SC_MMA_SCALE_RANGE words
static int scale_range(tdr_ctx* ctx, tdr::Particles& pt, bool track) {
  uint32_t* d_range = reinterpret_cast<uint32_t*>(ctx->scal.as<float>() + SC_MMA_SCALE_RANGE);
  static const uint32_t init_range[2] = {0xffffffffu, 0u};
  TDR_CUDA(cudaMemcpyAsync(d_range, init_range, 8, cudaMemcpyHostToDevice, ctx->stream));
  const int blocks = (int)((pt.n + 1023) / 1024 < ctx->sm_count * 4 ? (pt.n + 1023) / 1024 : ctx->sm_count * 4);
  k_scale_range<<<blocks < 1 ? 1 : blocks, 256, 0, ctx->stream>>>(pt.scale.as<float>(), pt.have_init.as<uint8_t>(), pt.n, d_range, track ? 1 : 0);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

// Head start for a theta search: everything that depends on the PARTICLES only (the spatial sort, 5 small kernels, and the
// scale check) runs on a side stream while the main stream rasterises the scan and builds the scan-dependent operand —
// ~0.1 ms of dependent small launches that used to sit between the rasteriser and the score kernel.  Only for sets in
// which every particle is searched (a mixed set sorts twice into the same buffer); the decision whether the integer
// kernel runs at all is taken later as before — a head start that is not used is simply joined.
int score_i8_prepare_async(tdr_ctx* ctx) {
  tdr::Particles& pt = ctx->part[ctx->cur];
  float q = 0.f;
  if (int e = score_i8_prepare_join(ctx)) return e;           // a head start nobody consumed (an error in between) is stale: drop it
  if (ctx->uninit_pending || pt.n <= 0 || ctx->n_uninit != pt.n) return TDR_OK;
  if (!(ctx->mma_kernel == 0 || ctx->mma_kernel == 3) || ctx->score_impl == 1) return TDR_OK;
  if (ctx->score_impl == 0 && (long long)ctx->n_uninit * ctx->count_scale < 4096) return TDR_OK;
  const int n_shifts = (int)ctx->search_shifts.size();
  if (!i8_usable(ctx, n_shifts, &q)) return TDR_OK;
  if (!ctx->prep_stream) {
    TDR_CUDA(cudaStreamCreateWithFlags(&ctx->prep_stream, cudaStreamNonBlocking));
    TDR_CUDA(cudaEventCreateWithFlags(&ctx->prep_fork, cudaEventDisableTiming));
    TDR_CUDA(cudaEventCreateWithFlags(&ctx->prep_done, cudaEventDisableTiming));
  }
  TDR_CUDA(cudaEventRecord(ctx->prep_fork, ctx->stream));          // after everything that last read the sort buffers
  TDR_CUDA(cudaStreamWaitEvent(ctx->prep_stream, ctx->prep_fork, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->prep_stream;                                   // the helpers launch on ctx->stream
  int e = build_perm(ctx, false, pt.n, ctx->mma_sort == 1);
  if (!e) e = scale_range(ctx, pt, false);
  ctx->stream = main_stream;
  if (e) { cudaStreamSynchronize(ctx->prep_stream); return e; }
  TDR_CUDA(cudaEventRecord(ctx->prep_done, ctx->prep_stream));
  ctx->prep_pending = true;
  return TDR_OK;
}
int score_i8_prepare_join(tdr_ctx* ctx) {
  if (!ctx->prep_pending) return TDR_OK;
  TDR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->prep_done, 0));
  ctx->prep_pending = false;
  return TDR_OK;
}

// One launch of the integer kernel.  Search (track_pass < 0): the particles without a heading against the candidate
// list.  Tracking pass p >= 0: the particles whose heading maps to a row shift in [shift_lo, shift_lo + n_shifts), against
// exactly those shifts (dev_shifts = shift_lo, shift_lo + 1, ...); each keeps the column of its own shift.
static int launch_i8(tdr_ctx* ctx, float res, const int32_t* dev_shifts, int n_shifts, int track_pass, int shift_lo, bool* used) {
  *used = false;
  const bool track = track_pass >= 0;
  float q = 0.f;
  if (!i8_usable(ctx, n_shifts, &q)) return TDR_OK;
  const int P = ctx->n_theta * ctx->n_r;
  if (((P + 3) & ~3) + 4 > MMA_TAB_MAX) return TDR_OK;         // gathered (+ padding) and skipped cells share the constant table
  // the previous scan's maximum count predicts whether this one fits a byte (the device check decides)
  // (later tracking passes do not look again: the passes stand or fall together, the device guard is the same for all)
  if (track_pass <= 0) {
    if (ctx->scan_max_pending && cudaEventQuery(ctx->scan_max_ev) == cudaSuccess) { ctx->scan_max_seen = *ctx->scan_max_pin; ctx->scan_max_pending = false; }
    if (ctx->mma_i8 != 2 && ctx->scan_max_seen > I8_MAX_COUNT) return TDR_OK;
  }
  const int P_cap = (P + 3) & ~3, max_stages = P_cap / I8_CELLS;
  if (int e = ctx->scan_op.reserve((size_t)max_stages * I8_NH * 32 + (size_t)(PLAN_HDR + 2 * P_cap) * 4)) return e;
  int* d_plan = reinterpret_cast<int*>(ctx->scan_op.as<unsigned char>() + (size_t)max_stages * I8_NH * 32);
  int* d_max = reinterpret_cast<int*>(ctx->scal.as<float>() + SC_MMA_MAXCOUNT);
  if (track_pass <= 0) TDR_CUDA(cudaMemsetAsync(d_max, 0, 8, ctx->stream));       // max count, "tensor-core kernel bailed out" flag (one verdict for all passes)
  {
    k_plan_cells<<<1, 1024, 0, ctx->stream>>>(ctx->scan_img.as<float>(), ctx->C, ctx->n_theta, ctx->n_r, dev_shifts, n_shifts,
                                              ctx->mma_skip_rings, d_plan);
    const long long total = (long long)max_stages * I8_NH;
    k_build_scan_operand_i8<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->scan_img.as<float>(), ctx->C, ctx->n_theta, P,
                                                                                       d_plan, dev_shifts, n_shifts, ctx->scan_op.as<uint4>(), d_max);
    count_launch(ctx, 2);
    TDR_CUDA(cudaGetLastError());
    TDR_CUDA(cudaMemcpyAsync(ctx->scan_max_pin, d_max, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaEventRecord(ctx->scan_max_ev, ctx->stream));
    ctx->scan_max_pending = true;
  }
  const int* count_dev = nullptr;
  const bool prepared = !track && ctx->prep_pending;          // sort + scale check already under way on the side stream
  if (prepared) { if (int e = score_i8_prepare_join(ctx)) return e; }
  else if (track) {
    if (int e = score_i8_prepare_join(ctx)) return e;         // (a stale head start must not race with this sort)
    if (int e = build_perm(ctx, false, ctx->part[ctx->cur].n, ctx->mma_sort == 1, true, shift_lo, shift_lo + n_shifts, &count_dev)) return e;
  }
  else if (int e = build_perm(ctx, false, ctx->part[ctx->cur].n, ctx->mma_sort == 1)) return e;
  tdr::Particles& pt = ctx->part[ctx->cur];
  // the lattice offsets for this launch's scale and radial resolution, x 4096, into constant memory; the scale comes
  // from the particles themselves (one value for all of them, else the kernel leaves the search to the CUDA cores)
  {
    uint32_t* d_range = reinterpret_cast<uint32_t*>(ctx->scal.as<float>() + SC_MMA_SCALE_RANGE);
    if (!prepared) { if (int e = scale_range(ctx, pt, track)) return e; }
    if (int e = ctx->tab_scaled.reserve((size_t)(P_cap + 4) * 8)) return e;
    k_scale_tab4096<<<(P_cap + 4 + 255) / 256, 256, 0, ctx->stream>>>(ctx->tab.as<float2>(), P, d_plan, d_range, res,
                                                                       ctx->tab_scaled.as<float2>(), d_max + 1);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    TDR_CUDA(cudaMemcpyToSymbolAsync(c_tab, ctx->tab_scaled.p, (size_t)(P_cap + 4) * 8, 0, cudaMemcpyDeviceToDevice, ctx->stream));
    g_tab_on_device[ctx->device % MMA_MAX_DEVICES] = 0;
  }
  I8Params sp; memset(&sp, 0, sizeof(sp));
  const bool blocked = ctx->mma_sort == 1;              // Morton order goes with the 4 x 2-px block layout
  if (int e = build_map8(ctx, q, blocked, &sp.map8, &sp.geom)) return e;
  sp.resolution = ctx->resolution; sp.P = P; sp.P_cap = P_cap; sp.plan = d_plan;
  sp.bop = ctx->scan_op.as<uint4>();
  sp.perm = ctx->perm.as<int>();
  sp.n_shifts = n_shifts;
  sp.q001 = 0.01f * q;
  sp.tab_g = ctx->tab_scaled.as<float2>();
  sp.tex = ctx->map8_tex;
  sp.maxcount = d_max; sp.bailed = d_max + 1;
  sp.n_work = track ? pt.n - ctx->n_uninit : ctx->n_uninit;      // tracking: an upper bound for the grid size; the kernel reads the real count
  sp.n_work_dev = track ? count_dev : nullptr;
  sp.track = track ? 1 : 0; sp.track_lo = shift_lo; sp.n_theta = ctx->n_theta;
  sp.init_x = pt.init_x.as<float>(); sp.init_y = pt.init_y.as<float>(); sp.dx = pt.dx.as<float>(); sp.dy = pt.dy.as<float>();
  sp.theta = pt.theta.as<float>(); sp.scale = pt.scale.as<float>(); sp.have_init = pt.have_init.as<uint8_t>();
  sp.weights = ctx->weights.as<float>();
  sp.force_on_map = ctx->fp.force_on_map;
  sp.map_w = (float)ctx->cols * ctx->resolution; sp.map_h = (float)ctx->rows * ctx->resolution;
  sp.scale_gate = ctx->fp.fixed_scale < 0 ? 1 : 0;
  sp.scale_lo = pow(10.0, (double)ctx->fp.scale_log_min); sp.scale_hi = pow(10.0, (double)ctx->fp.scale_log_max);
  sp.regularization = ctx->fp.regularization;
  sp.thetas = ctx->d_search_thetas.as<float>();
#define I8_OPTIN(BIT, KERNEL)                                                                                         \
  do {                                                                                                                \
    if (!(ctx->smem_optin_i8 & (1ull << (BIT)))) {                                                                    \
      TDR_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));                \
      ctx->smem_optin_i8 |= 1ull << (BIT);                                                                            \
    }                                                                                                                 \
  } while (0)
#define TDR_LAUNCH_I8(IDX, TT, RR, FF)                                                                                \
  do {                                                                                                                \
    using Cfg = I8Cfg<TT, RR, FF>;                                                                                    \
    if (sp.tex) I8_OPTIN(IDX, (k_score_mma_i8<TT, RR, FF, true, false>));                                             \
    else if (blocked) I8_OPTIN(16 + IDX, (k_score_mma_i8<TT, RR, FF, false, true>));                                  \
    else I8_OPTIN(32 + IDX, (k_score_mma_i8<TT, RR, FF, false, false>));                                              \
    const long long nb = (sp.n_work + 128 * TT - 1) / (128 * TT);                                                     \
    long long cap = (long long)ctx->sm_count * Cfg::kCtasPerSm;                                                       \
    if (ctx->mma_grid_cap > 0 && ctx->mma_grid_cap < cap) cap = ctx->mma_grid_cap;                                    \
    const int grid = (int)(nb < cap ? nb : cap);                                                                      \
    if (sp.tex) k_score_mma_i8<TT, RR, FF, true, false><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);        \
    else if (blocked) k_score_mma_i8<TT, RR, FF, false, true><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);  \
    else k_score_mma_i8<TT, RR, FF, false, false><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);              \
  } while (0)
  switch (ctx->mma_i8_cfg) {                     // tiles * 100 + threads per row * 10 + stages in flight (TDR_MMA_I8_CFG)
    case 131: TDR_LAUNCH_I8(0, 1, 3, 1); break;
    case 222: TDR_LAUNCH_I8(1, 2, 2, 2); break;
    case 223: TDR_LAUNCH_I8(2, 2, 2, 3); break;
    case 141: TDR_LAUNCH_I8(5, 1, 4, 1); break;
    case 151: TDR_LAUNCH_I8(8, 1, 5, 1); break;
    case 132: TDR_LAUNCH_I8(3, 1, 3, 2); break;
    case 122: TDR_LAUNCH_I8(4, 1, 2, 2); break;
    case 123: TDR_LAUNCH_I8(10, 1, 2, 3); break;
    case 121: TDR_LAUNCH_I8(11, 1, 2, 1); break;
    case 161: TDR_LAUNCH_I8(9, 1, 6, 1); break;
    case 231: TDR_LAUNCH_I8(6, 2, 3, 1); break;
    default: TDR_LAUNCH_I8(7, 2, 3, 2); break;
  }
#undef TDR_LAUNCH_I8
#undef I8_OPTIN
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  *used = true;
  return TDR_OK;
}

int score_mma_i8(tdr_ctx* ctx, float res, const int32_t* dev_shifts, int n_shifts, bool* used) {
  return launch_i8(ctx, res, dev_shifts, n_shifts, -1, 0, used);
}

// Tracking LARGE particle sets (steady-state global localisation: 1e6 particles that all have a heading): the heading of
// a particle selects ONE of the n_theta row shifts, a different one per particle, so a tile of 128 particles needs a
// whole window of shifts.  The particles are binned by their shift into windows of 40 (the operand's candidate rows)
// and every window is one launch of the search kernel in which each particle keeps its own column: together the passes
// gather every particle's cells once — half the time of the fp16 ring kernel's all-shift pass
// (state_particle.cpp:207-212).  *used = false: not this kernel's case, nothing was launched.
int score_mma_i8_track(tdr_ctx* ctx, float res, bool* used) {
  *used = false;
  float q = 0.f;
  const int n_theta = ctx->n_theta;
  if (n_theta < 1 || !i8_usable(ctx, n_theta < I8_S_MAX ? n_theta : I8_S_MAX, &q)) return TDR_OK;
  if (ctx->ident_shifts_n != n_theta) {
    std::vector<int32_t> h((size_t)n_theta);
    for (int k = 0; k < n_theta; k++) h[(size_t)k] = k;
    if (int e = ctx->ident_shifts.reserve((size_t)n_theta * 4)) return e;
    TDR_CUDA(cudaMemcpyAsync(ctx->ident_shifts.p, h.data(), (size_t)n_theta * 4, cudaMemcpyHostToDevice, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));           // h goes out of scope; once per table
    ctx->ident_shifts_n = n_theta;
  }
  int pass = 0;
  for (int lo = 0; lo < n_theta; lo += I8_S_MAX, pass++) {
    const int S = n_theta - lo < I8_S_MAX ? n_theta - lo : I8_S_MAX;
    bool u = false;
    if (int e = launch_i8(ctx, res, ctx->ident_shifts.as<int32_t>() + lo, S, pass, lo, &u)) return e;
    if (!u) {
      // a pass that declines after an earlier one ran would leave some particles unscored: the preconditions are the same
      // for every pass (operand format, table size, predicted scan maximum), so only the first can decline
      TDR_REQUIRE(pass == 0, TDR_ESTATE, "integer tracking pass %d declined after earlier passes ran", pass);
      return TDR_OK;
    }
  }
  *used = true;
  return TDR_OK;
}

}  // namespace tdr
