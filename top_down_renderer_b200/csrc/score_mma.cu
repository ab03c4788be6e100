// score_mma.cu — a7/a9/a10 as a tcgen05 gather-GEMM (theta search and exhaustive grid).
//
// Reference: TopDownMapPolar::getLocalMap (src/top_down_map_polar.cpp:21-53), StateParticle::getCostForRot
// (src/state_particle.cpp:112-155), StateParticle::computeWeight (:157-219).
//
// For a tile of 128 hypotheses the costs of all candidate row shifts are one matrix product
//     D[m, n] = sum_k A[m, k] * B[k, n],      k = (lattice cell p, slot j)
//   A[m, (p, j)]  = the map record of hypothesis m at lattice cell p: 16 fp16 = 8 hi + 8 lo halves of
//                   w_c * dist_c (class weight folded into the map copy; hi + lo carries 22 mantissa bits),
//                   slot 7 hi = known.  One 32-byte record = one L2 sector = one K = 16 step.
//   B[(p, j), n]  = the scan circulant: n < S      : class count at the shifted angle (both hi and lo slots)
//                                       n == S     : 1 at slot 7 -> D = number of known cells
//                                       n = S_pad+s: class-summed count at slot 7 -> the normalisation
// fp32 accumulation in TMEM.  cost[s] = 0.01 * D[m, s] / D[m, S_pad + s]  (:136-154).
//
// Warp roles (T tiles of 128 hypotheses per CTA, one CTA per SM, persistent over batches):
//   warps [0, 4T)  gather: thread = hypothesis; computes its lattice pixel per cell (exact index math of
//                  tdr_math.cuh), loads the 32-byte record and stores it straight into the K-major UMMA operand
//                  layout; afterwards the same threads run the epilogue on their TMEM row.
//   warp 4T        one lane streams the precomputed scan operand with cp.async.bulk (mbarrier complete_tx).
//   warp 4T+1      one lane issues tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 2*S_pad, K = 16) per
//                  (cell, tile); tcgen05.commit frees the stage / publishes the accumulators.
#include <cuda_fp16.h>

#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
      : "l"(p));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle (cute UMMA "INTERLEAVE"): core matrix = 8 rows x 16 B contiguous;
// LBO = byte step between the two K chunks, SBO = byte step between 8-row groups; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}

// ------------------------------------------------------------------------------------------------
// operand builders
// ------------------------------------------------------------------------------------------------
// MapPixel (8 fp32) -> 16 fp16: hi[0..7] | lo[0..7];  value_c = w_c * dist_c, slot 7 = known (hi only)
__global__ void k_build_map16(const MapPixel* __restrict__ map, size_t n, int C, const float* __restrict__ cw,
                              uint4* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* src = reinterpret_cast<const float4*>(map + i);
  float4 a = src[0], b = src[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __half hi[8], lo[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    float x = 0.f;
    if (c < C) x = TDR_FMUL(cw[c], v[c]);
    if (c == 7) x = v[7];
    hi[c] = __float2half_rn(x);
    lo[c] = (c == 7) ? __float2half_rn(0.f) : __float2half_rn(TDR_FSUB(x, __half2float(hi[c])));
  }
  uint4 h, l;
  h.x = (uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16);
  h.y = (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16);
  h.z = (uint32_t)__half_as_ushort(hi[4]) | ((uint32_t)__half_as_ushort(hi[5]) << 16);
  h.w = (uint32_t)__half_as_ushort(hi[6]) | ((uint32_t)__half_as_ushort(hi[7]) << 16);
  l.x = (uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16);
  l.y = (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16);
  l.z = (uint32_t)__half_as_ushort(lo[4]) | ((uint32_t)__half_as_ushort(lo[5]) << 16);
  l.w = (uint32_t)__half_as_ushort(lo[6]) | ((uint32_t)__half_as_ushort(lo[7]) << 16);
  out[2 * i] = h; out[2 * i + 1] = l;
}

// scan operand, per cell p: [kc = 2][n = N][8 halfs]  (K-major canonical layout, LBO = N*16 B, SBO = 128 B)
// one thread per (p, n).  maxcount: device int, max class-summed count seen (fp16 integers are exact up to 2048).
__global__ void k_build_scan_operand(const float* __restrict__ img, int C, int n_theta, int n_r, int P, int P_pad,
                                     const int32_t* __restrict__ shifts, int S, int S_pad, int N,
                                     uint4* __restrict__ out, int* __restrict__ maxcount) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)P_pad * N) return;
  const int p = (int)(id / N), n = (int)(id - (long long)p * N);
  unsigned short h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool lo_copy = false;
  if (p < P) {
    const int r = p / n_theta, th = p - r * n_theta;
    if (n < S || (n >= S_pad && n < S_pad + S)) {
      const int s = n < S ? n : n - S_pad;
      int t2 = th + shifts[s];
      t2 %= n_theta; if (t2 < 0) t2 += n_theta;
      const int cell = r * n_theta + t2;          // scan row (theta + shift) pairs with map row theta
      float tot = 0.f;
      for (int c = 0; c < C; c++) {
        float v = img[(size_t)c * P + cell];
        tot += v;
        if (n < S) h[c] = __half_as_ushort(__float2half_rn(v));
      }
      if (n < S) lo_copy = true;
      else h[7] = __half_as_ushort(__float2half_rn(tot));
      if (n == S_pad) atomicMax(maxcount, (int)tot);     // shift[0] is a bijection of the cells: global max of tot
    } else if (n == S) {
      h[7] = 0x3C00;                              // 1.0: counts the known cells
    }
  }
  uint4 q;
  q.x = h[0] | ((uint32_t)h[1] << 16); q.y = h[2] | ((uint32_t)h[3] << 16);
  q.z = h[4] | ((uint32_t)h[5] << 16); q.w = h[6] | ((uint32_t)h[7] << 16);
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4* cellbase = out + (size_t)p * N * 2;      // N*32 B per cell = 2N uint4
  cellbase[n] = q;                                // kc = 0: hi slots
  cellbase[N + n] = lo_copy ? q : z;              // kc = 1: lo slots see the same counts (slot 7 lo is never set)
}

// ------------------------------------------------------------------------------------------------
// spatial binning of the hypotheses (L2 locality): counting sort by coarse map tile
// ------------------------------------------------------------------------------------------------
struct BinParams {
  const float *init_x, *init_y, *dx, *dy, *scale; const uint8_t* have_init;   // particle mode
  const float* centers;                                                       // grid mode
  long long n; float resolution; int rows, cols, st_shift, seg_shift, super_x, per_super, n_bins;
};
__device__ __forceinline__ int bin_of(const BinParams& b, long long i) {
  float x, y;
  if (b.centers) { x = b.centers[2 * i]; y = b.centers[2 * i + 1]; }
  else {
    if (b.have_init[i]) return -1;                // tracked by k_score_track
    float s = b.scale[i];
    x = TDR_FADD(TDR_FMUL(b.dx[i], s), b.init_x[i]); y = TDR_FADD(TDR_FMUL(b.dy[i], s), b.init_y[i]);
  }
  int c = f2i_x86(TDR_FDIV(x, b.resolution)), r = f2i_x86(TDR_FDIV(y, b.resolution));
  if (c < 0 || r < 0 || c >= b.cols || r >= b.rows) return b.n_bins - 1;      // off-map: last bin
  // super-tile (2^st_shift px square, row-major over the map), then pixel row, then 32-px column segment
  const int S = 1 << b.st_shift, m = S - 1;
  const int sup = (r >> b.st_shift) * b.super_x + (c >> b.st_shift);
  return sup * b.per_super + (r & m) * (S >> b.seg_shift) + ((c & m) >> b.seg_shift);
}
__global__ void k_bin_count(BinParams b, int* __restrict__ counts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (long long)gridDim.x * blockDim.x) {
    int k = bin_of(b, i);
    if (k >= 0) atomicAdd(counts + k, 1);
  }
}
// exclusive scan of the bin counts (in place): per-block local scan + block sums, scan of the sums, add back
static const int SCAN_ITEMS = 4, SCAN_BLOCK = 1024, SCAN_TILE = SCAN_ITEMS * SCAN_BLOCK;
__device__ __forceinline__ int block_excl_scan(int v, int* s_w, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_w[lane];
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += o; }
    s_w[lane] = w;
  }
  __syncthreads();
  const int excl = inc - v + (warp > 0 ? s_w[warp - 1] : 0);
  if (total) *total = s_w[31];
  __syncthreads();
  return excl;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_local(int* __restrict__ a, int n, int* __restrict__ sums) {
  __shared__ int s_w[32];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], t = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? a[base + k] : 0; t += v[k]; }
  int total;
  int excl = block_excl_scan(t, s_w, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) a[base + k] = excl; excl += v[k]; }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int* __restrict__ sums, int nb) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int total;
    const int excl = block_excl_scan(v, s_w, &total) + s_carry;
    if (i < nb) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(int* __restrict__ a, int n, const int* __restrict__ sums) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  const int add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) a[base + k] += add;
}
__global__ void k_bin_scatter(BinParams b, int* __restrict__ cursor, int* __restrict__ perm) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (long long)gridDim.x * blockDim.x) {
    int k = bin_of(b, i);
    if (k >= 0) perm[atomicAdd(cursor + k, 1)] = (int)i;
  }
}

// ------------------------------------------------------------------------------------------------
// the gather-GEMM
// ------------------------------------------------------------------------------------------------
// The polar offset table is read with a warp-uniform index once per cell by every gather thread.  As a global
// load it queued behind the record loads in the L1 pipe and a quarter of all stall samples sat on the first
// FMUL of the index math; from constant memory (uniform LDC through the constant cache) it is off that path.
static const int MMA_TAB_MAX = 4096;
__constant__ float2 c_tab[MMA_TAB_MAX];
static const void* g_tab_owner = nullptr;     // device table currently mirrored in c_tab

struct MmaParams {
  const uint4* map16; int rows, cols; float resolution;
  const float2* tab; int P, P_pad; float res;
  const uint4* bop;
  const int* perm; long long n_work;
  // particle mode
  const float *init_x, *init_y, *dx, *dy; float* theta; const float* scale; uint8_t* have_init; float* weights;
  int force_on_map; float map_w, map_h; int scale_gate; double scale_lo, scale_hi; float regularization;
  const float* thetas; int n_shifts;
  // grid mode
  const float* centers; float grid_scale; float* costs;
};

static const int MMA_G = 2;        // lattice cells per pipeline stage
// A tile (128 hypotheses x 16 fp16, K-major, no swizzle): K chunk 0 at [0, 2048), K chunk 1 at [A_LBO, A_LBO + 2048).
// A_LBO is 64 bytes past a multiple of 128 so that a quarter-warp's 4 rows x 2 chunks hit 8 different 16-byte
// bank groups (conflict-free STS.128).
static const int A_LBO = 2048 + 64;
static const int A_TILE = 4224;
// T = 128-hypothesis tiles per CTA; R = gather threads per hypothesis row (the R threads of a row take turns
// stage by stage, so the loads in flight per SM double without doubling the hypotheses — and their map
// footprint — that are in flight together)
template <int N, int T, int R> struct MmaCfg {
  static const int kThreads = 128 * T * R + 64;
  static const int kTmemCols = T * N <= 128 ? 128 : (T * N <= 256 ? 256 : 512);
  static const int kByTmem = 512 / kTmemCols, kByRegs = 65536 / (kThreads * 88) < 1 ? 1 : 65536 / (kThreads * 88);
  static const int kCtasPerSm = kByTmem < kByRegs ? kByTmem : kByRegs;
  static const int kABytes = MMA_G * T * A_TILE;          // per stage
  static const int kBBytes = MMA_G * N * 32;              // per stage
  static const int kStageBytes = kABytes + kBBytes;
  static const int kBudget = (216 * 1024) / kCtasPerSm - 1280;
  static const int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static const int kSmem = kStages * kStageBytes + 256;
};

template <int N, int T, int R>
__global__ void __launch_bounds__(128 * T * R + 64, MmaCfg<N, T, R>::kCtasPerSm) k_score_mma(MmaParams sp) {
  using Cfg = MmaCfg<N, T, R>;
  constexpr int GW = 4 * T * R;        // gather warps
  constexpr int NS = Cfg::kStages;
  constexpr int S_PAD = N / 2;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                                  // [NS][G][T] tiles of A_TILE bytes
  unsigned char* sB = smem + (size_t)NS * Cfg::kABytes;       // [NS][G][kc 2][N][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * Cfg::kStageBytes);   // full[NS] empty[NS] accum
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
  const long long n_batches = (sp.n_work + per_batch - 1) / per_batch;
  const int K_ITERS = sp.P_pad / MMA_G;
  uint32_t it = 0;                 // pipeline iteration counter, continues across batches (same sequence in every role)
  uint32_t local_batch = 0;

  if (warp < GW) {
    // =========================== gather + epilogue ===========================
    const int sub = warp / (4 * T);                      // which of the R threads of a row this is
    const int t = (warp % (4 * T)) >> 2, m = tid & 127;
    const unsigned char* map_bytes = reinterpret_cast<const unsigned char*>(sp.map16);
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, local_batch++) {
      const long long slot = batch * per_batch + (tid % (128 * T));
      long long i = -1;
      if (slot < sp.n_work) i = sp.perm ? (long long)sp.perm[slot] : slot;
      float cx = 0.f, cy = 0.f, sc = 1.f;
      bool active = false, gated = false;
      if (i >= 0) {
        if (sp.centers) { cx = sp.centers[2 * i]; cy = sp.centers[2 * i + 1]; sc = sp.grid_scale; active = true; }
        else {
          sc = sp.scale[i];
          cx = TDR_FADD(TDR_FMUL(sp.dx[i], sc), sp.init_x[i]);
          cy = TDR_FADD(TDR_FMUL(sp.dy[i], sc), sp.init_y[i]);
          if (sp.force_on_map && (cx < 0.f || cy < 0.f || cx > sp.map_w || cy > sp.map_h)) gated = true;      // :163-168
          if (sp.scale_gate && ((double)sc < sp.scale_lo || (double)sc > sp.scale_hi)) gated = true;          // :169-176
          active = !gated;
        }
      }
      const float oy = TDR_FDIV(cy, sp.resolution), ox = TDR_FDIV(cx, sp.resolution);

      // Each thread pulls the whole 32-byte record of ITS hypothesis with one 256-bit load (one sector, one L1
      // wavefront; measured 0.95 records/clk/SM from L2 against 0.42 for 2 x LDG.128 — tools/gather_bench.cu).
      auto load_stage = [&](int k, uint4 (&rec)[MMA_G][2]) {
#pragma unroll
        for (int g = 0; g < MMA_G; g++) {
          const int p = k * MMA_G + g;
          rec[g][0] = make_uint4(0, 0, 0, 0); rec[g][1] = rec[g][0];
          if (active && p < sp.P) {
            const float2 tb = c_tab[p];
            const int r = lattice_index(tb.x, sc, sp.res, oy);
            const int c = lattice_index(tb.y, sc, sp.res, ox);
            if (r >= 0 && r < sp.rows && c >= 0 && c < sp.cols)
              ldg256(map_bytes + ((size_t)r * sp.cols + c) * 32, rec[g][0], rec[g][1]);
          }
        }
      };
      auto store_stage = [&](uint32_t iter, const uint4 (&rec)[MMA_G][2]) {
        const uint32_t st = iter % NS, ph = (iter / NS) & 1u;
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        unsigned char* base = sA + (size_t)st * Cfg::kABytes + (size_t)t * A_TILE + (size_t)m * 16;
#pragma unroll
        for (int g = 0; g < MMA_G; g++) {
          *reinterpret_cast<uint4*>(base + (size_t)g * T * A_TILE) = rec[g][0];            // K chunk 0: hi halves
          *reinterpret_cast<uint4*>(base + (size_t)g * T * A_TILE + A_LBO) = rec[g][1];    // K chunk 1: lo halves
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
      };

      uint4 ra[MMA_G][2], rb[MMA_G][2];
      if (sub < K_ITERS) load_stage(sub, ra);
#pragma unroll 1
      for (int k = sub; k < K_ITERS; k += 2 * R) {       // this thread's stages: sub, sub + R, ... (two in flight)
        if (k + R < K_ITERS) load_stage(k + R, rb);
        store_stage(it + k, ra);
        if (k + 2 * R < K_ITERS) load_stage(k + 2 * R, ra);
        if (k + R < K_ITERS) store_stage(it + k + R, rb);
      }
      it += K_ITERS;
      if (sub != 0) continue;                            // the first thread of each row owns the epilogue

      // ---- epilogue: this thread's accumulator row
      mbar_wait(bar_accum, local_batch & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * N);
      uint32_t vc[16], vn[16];
      const uint32_t kraw = tmem_ld1(trow + (uint32_t)sp.n_shifts);
      tmem_wait_ld();
      const float known = __uint_as_float(kraw);
      const bool unknown = (double)TDR_FDIV(known, (float)sp.P) < 0.5;                       // :117-120
      float best = 3.402823466e+38f, best_theta = 0.f;                                       // :193-204
#pragma unroll 1
      for (int ch = 0; ch * 16 < sp.n_shifts; ch++) {
        tmem_ld16(trow + (uint32_t)(ch * 16), vc);
        tmem_ld16(trow + (uint32_t)(S_PAD + ch * 16), vn);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int s = ch * 16 + j;
          if (s < sp.n_shifts) {
            float cost = unknown ? __int_as_float(0x7fc00000)
                                 : TDR_FDIV(TDR_FMUL(__uint_as_float(vc[j]), 0.01f), __uint_as_float(vn[j]));   // :137,154
            if (sp.costs && i >= 0) sp.costs[i * sp.n_shifts + s] = cost;
            if (cost < best) { best = cost; best_theta = sp.thetas ? sp.thetas[s] : 0.f; }
          }
        }
      }
      if (i >= 0 && !sp.centers) {
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
    if (lane == 0) {
      for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        for (int k = 0; k < K_ITERS; k++, it++) {
          const uint32_t st = it % NS, ph = (it / NS) & 1u;
          mbar_wait(bar_empty + 8 * st, ph ^ 1u);
          mbar_expect_tx(bar_full + 8 * st, Cfg::kBBytes);
          bulk_g2s(smem_u32(sB + (size_t)st * Cfg::kBBytes),
                   reinterpret_cast<const unsigned char*>(sp.bop) + (size_t)k * Cfg::kBBytes, Cfg::kBBytes,
                   bar_full + 8 * st);
        }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = f16, both K-major, N, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        for (int k = 0; k < K_ITERS; k++, it++) {
          const uint32_t st = it % NS, ph = (it / NS) & 1u;
          mbar_wait(bar_full + 8 * st, ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + (size_t)st * Cfg::kABytes), b0 = smem_u32(sB + (size_t)st * Cfg::kBBytes);
#pragma unroll
          for (int g = 0; g < MMA_G; g++) {
            const uint64_t bdesc = umma_desc(b0 + g * (N * 32), N * 16, 128);
#pragma unroll
            for (int tt = 0; tt < T; tt++) {
              const uint64_t adesc = umma_desc(a0 + (g * T + tt) * A_TILE, A_LBO, 128);
              umma_f16(tmem_base + (uint32_t)(tt * N), adesc, bdesc, idesc, (k > 0 || g > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * st);        // implies tcgen05.fence::before_thread_sync
        }
        umma_commit(bar_accum);
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
static int build_map16(tdr_ctx* ctx) {
  if (ctx->map16_valid) return TDR_OK;
  const size_t L = (size_t)ctx->rows * ctx->cols;
  if (int e = ctx->map16.reserve(L * 32)) return e;
  k_build_map16<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->map_px.as<MapPixel>(), L, ctx->C,
                                                                       ctx->d_cw.as<float>(), ctx->map16.as<uint4>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  ctx->map16_valid = true;
  return TDR_OK;
}

bool mma_usable(tdr_ctx* ctx, int n_shifts) {
  if (ctx->score_impl == 1) return false;
  if (n_shifts < 1 || n_shifts > 111) return false;
  for (int c = 0; c < ctx->C; c++) {
    float w = ctx->fp.class_weights[c];
    if (!(w >= 0.f) || w * 50.f > 60000.f) return false;     // fp16 range of w_c * dist_c (dist <= 50)
  }
  return true;
}

// returns TDR_OK and sets *used = true when the tensor-core path ran; *used = false -> caller falls back
int score_mma(tdr_ctx* ctx, float res, bool grid_mode, long long n_items, float grid_scale, const int32_t* dev_shifts,
              int n_shifts, bool* used) {
  *used = false;
  if (!mma_usable(ctx, n_shifts)) return TDR_OK;
  const int P = ctx->n_theta * ctx->n_r;
  if (P > MMA_TAB_MAX) return TDR_OK;
  const int S_pad = n_shifts + 1 <= 48 ? 48 : 112;
  const int N = 2 * S_pad;
  const int P_pad = (P + 2 * MMA_G - 1) / (2 * MMA_G) * (2 * MMA_G);     // even number of stages keeps the 2x unroll simple
  // ---- scan operand (+ max count check: fp16 integers are exact up to 2048)
  if (int e = ctx->scan_op.reserve((size_t)P_pad * N * 32)) return e;
  int* d_max = reinterpret_cast<int*>(ctx->scal.as<float>() + SC_MMA_MAXCOUNT);
  TDR_CUDA(cudaMemsetAsync(d_max, 0, 4, ctx->stream));
  {
    long long total = (long long)P_pad * N;
    k_build_scan_operand<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->scan_img.as<float>(), ctx->C, ctx->n_theta, ctx->n_r, P, P_pad, dev_shifts, n_shifts, S_pad, N,
        ctx->scan_op.as<uint4>(), d_max);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
  }
  int h_max = 0;
  TDR_CUDA(cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h_max > 2048) return TDR_OK;              // counts not exact in fp16: CUDA-core path
  if (int e = build_map16(ctx)) return e;

  // ---- spatial binning -> perm
  tdr::Particles& pt = ctx->part[ctx->cur];
  BinParams bp; memset(&bp, 0, sizeof(bp));
  if (grid_mode) bp.centers = ctx->grid_centers.as<float>();
  else {
    bp.init_x = pt.init_x.as<float>(); bp.init_y = pt.init_y.as<float>(); bp.dx = pt.dx.as<float>(); bp.dy = pt.dy.as<float>();
    bp.scale = pt.scale.as<float>(); bp.have_init = pt.have_init.as<uint8_t>();
  }
  bp.n = n_items; bp.resolution = ctx->resolution; bp.rows = ctx->rows; bp.cols = ctx->cols;
  // bins = (super-tile, pixel row, 32-px column segment).  The super-tile keeps the hypotheses that are in
  // flight together (sm_count x 128T of them) inside one compact region so that its dilated footprint stays
  // L2-resident; row + segment order puts warp neighbours on the same map row a few pixels apart (shared lines).
  bp.st_shift = ctx->mma_st_shift;
  while ((1 << bp.st_shift) < 32) bp.st_shift++;
  bp.super_x = (ctx->cols >> bp.st_shift) + 1;
  bp.seg_shift = ctx->mma_seg_shift;
  bp.per_super = (1 << bp.st_shift) * ((1 << bp.st_shift) >> bp.seg_shift);
  {
    long long nb = (long long)((ctx->rows >> bp.st_shift) + 1) * bp.super_x * bp.per_super + 1;
    TDR_REQUIRE(nb < (1ll << 28), TDR_EUNSUPPORTED, "map too large for the hypothesis binning (%lld bins)", nb);
    bp.n_bins = (int)nb;
  }
  const int scan_blocks = (bp.n_bins + SCAN_TILE - 1) / SCAN_TILE;
  if (int e = ctx->bin_counts.reserve((size_t)bp.n_bins * 4 + (size_t)scan_blocks * 4 + 64)) return e;
  if (int e = ctx->perm.reserve((size_t)n_items * 4)) return e;
  int* d_counts = ctx->bin_counts.as<int>();
  int* d_sums = d_counts + bp.n_bins;
  TDR_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)bp.n_bins * 4, ctx->stream));
  const int blocks = (int)((n_items + 255) / 256 < ctx->sm_count * 8 ? (n_items + 255) / 256 : ctx->sm_count * 8);
  k_bin_count<<<blocks, 256, 0, ctx->stream>>>(bp, d_counts);
  k_scan_local<<<scan_blocks, SCAN_BLOCK, 0, ctx->stream>>>(d_counts, bp.n_bins, d_sums);
  k_scan_sums<<<1, SCAN_BLOCK, 0, ctx->stream>>>(d_sums, scan_blocks);
  k_scan_add<<<scan_blocks, SCAN_BLOCK, 0, ctx->stream>>>(d_counts, bp.n_bins, d_sums);
  k_bin_scatter<<<blocks, 256, 0, ctx->stream>>>(bp, d_counts, ctx->perm.as<int>());
  count_launch(ctx, 5);
  TDR_CUDA(cudaGetLastError());

  if (g_tab_owner != ctx->tab.p || ctx->tab_dirty) {
    TDR_CUDA(cudaMemcpyToSymbolAsync(c_tab, ctx->tab.p, (size_t)P * 8, 0, cudaMemcpyDeviceToDevice, ctx->stream));
    g_tab_owner = ctx->tab.p; ctx->tab_dirty = false;
  }
  MmaParams sp; memset(&sp, 0, sizeof(sp));
  sp.map16 = ctx->map16.as<uint4>(); sp.rows = ctx->rows; sp.cols = ctx->cols; sp.resolution = ctx->resolution;
  sp.tab = ctx->tab.as<float2>(); sp.P = P; sp.P_pad = P_pad; sp.res = res;
  sp.bop = ctx->scan_op.as<uint4>();
  sp.perm = ctx->perm.as<int>();
  sp.n_shifts = n_shifts;
  if (grid_mode) {
    sp.n_work = n_items; sp.centers = ctx->grid_centers.as<float>(); sp.grid_scale = grid_scale;
    sp.costs = ctx->grid_costs.as<float>();
  } else {
    sp.n_work = ctx->n_uninit;
    sp.init_x = pt.init_x.as<float>(); sp.init_y = pt.init_y.as<float>(); sp.dx = pt.dx.as<float>(); sp.dy = pt.dy.as<float>();
    sp.theta = pt.theta.as<float>(); sp.scale = pt.scale.as<float>(); sp.have_init = pt.have_init.as<uint8_t>();
    sp.weights = ctx->weights.as<float>();
    sp.force_on_map = ctx->fp.force_on_map;
    sp.map_w = (float)ctx->cols * ctx->resolution; sp.map_h = (float)ctx->rows * ctx->resolution;
    sp.scale_gate = ctx->fp.fixed_scale < 0 ? 1 : 0;
    sp.scale_lo = pow(10.0, (double)ctx->fp.scale_log_min); sp.scale_hi = pow(10.0, (double)ctx->fp.scale_log_max);
    sp.regularization = ctx->fp.regularization;
    sp.thetas = ctx->d_search_thetas.as<float>();
  }
#define TDR_LAUNCH_MMA(NN, TT, RR)                                                                                   \
  do {                                                                                                                \
    using Cfg = MmaCfg<NN, TT, RR>;                                                                                   \
    static bool attr = false;                                                                                         \
    if (!attr) { TDR_CUDA(cudaFuncSetAttribute(k_score_mma<NN, TT, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem)); attr = true; } \
    const long long nb = (sp.n_work + 128 * TT - 1) / (128 * TT);                                                     \
    const long long cap = (long long)ctx->sm_count * (ctx->mma_ctas > 0 && ctx->mma_ctas < Cfg::kCtasPerSm ? ctx->mma_ctas : Cfg::kCtasPerSm); \
    const int grid = (int)(nb < cap ? nb : cap);                                                                      \
    k_score_mma<NN, TT, RR><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);                                    \
  } while (0)
  const int cfg = ctx->mma_tiles * 10 + ctx->mma_split;
  if (S_pad == 48) {
    switch (cfg) {
      case 41: TDR_LAUNCH_MMA(96, 4, 1); break;
      case 21: TDR_LAUNCH_MMA(96, 2, 1); break;
      case 22: TDR_LAUNCH_MMA(96, 2, 2); break;
      case 11: TDR_LAUNCH_MMA(96, 1, 1); break;
      case 14: TDR_LAUNCH_MMA(96, 1, 4); break;
      default: TDR_LAUNCH_MMA(96, 1, 2); break;
    }
  } else {
    switch (cfg) {
      case 21: case 41: TDR_LAUNCH_MMA(224, 2, 1); break;
      case 11: TDR_LAUNCH_MMA(224, 1, 1); break;
      default: TDR_LAUNCH_MMA(224, 1, 2); break;
    }
  }
#undef TDR_LAUNCH_MMA
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  *used = true;
  return TDR_OK;
}

}  // namespace tdr
