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
  long long n; float resolution; int rows, cols, tile_shift, tiles_x, n_bins;
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
  return (r >> b.tile_shift) * b.tiles_x + (c >> b.tile_shift);
}
__global__ void k_bin_count(BinParams b, int* __restrict__ counts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (long long)gridDim.x * blockDim.x) {
    int k = bin_of(b, i);
    if (k >= 0) atomicAdd(counts + k, 1);
  }
}
// single CTA exclusive scan of counts -> cursor (in place)
__global__ void __launch_bounds__(1024) k_bin_scan(int* __restrict__ counts, int n_bins) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_bins; base += 1024) {
    int i = base + tid;
    int v = i < n_bins ? counts[i] : 0;
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
    int excl = inc - v + (warp > 0 ? s_w[warp - 1] : 0) + s_carry;
    if (i < n_bins) counts[i] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
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
template <int N, int T> struct MmaCfg {
  static const int kThreads = 128 * T + 64;
  static const int kABytes = MMA_G * T * 4096;           // per stage
  static const int kBBytes = MMA_G * N * 32;             // per stage
  static const int kStageBytes = kABytes + kBBytes;
  static const int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static const int kSmem = kStages * kStageBytes + 256;
};

template <int N, int T>
__global__ void __launch_bounds__(128 * T + 64, 1) k_score_mma(MmaParams sp) {
  using Cfg = MmaCfg<N, T>;
  constexpr int NS = Cfg::kStages;
  constexpr int S_PAD = N / 2;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                                  // [NS][G][T][kc 2][128][16 B]
  unsigned char* sB = smem + (size_t)NS * Cfg::kABytes;       // [NS][G][kc 2][N][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * Cfg::kStageBytes);   // full[NS] empty[NS] accum
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NS + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NS), bar_accum = smem_u32(bars + 2 * NS);

  if (warp == 4 * T + 1) tmem_alloc(smem_u32(s_tmem), 512);
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

  if (warp < 4 * T) {
    // =========================== gather + epilogue ===========================
    const int t = warp >> 2, m = tid & 127;
    const unsigned char* map_bytes = reinterpret_cast<const unsigned char*>(sp.map16);
    for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, local_batch++) {
      const long long slot = batch * per_batch + tid;
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

      auto load_stage = [&](int k, uint4 (&rec)[MMA_G][2]) {
#pragma unroll
        for (int g = 0; g < MMA_G; g++) {
          rec[g][0] = make_uint4(0, 0, 0, 0); rec[g][1] = rec[g][0];
          const int p = k * MMA_G + g;
          if (active && p < sp.P) {
            const float2 tb = __ldg(sp.tab + p);
            const int r = lattice_index(tb.x, sc, sp.res, oy);
            const int c = lattice_index(tb.y, sc, sp.res, ox);
            if (r >= 0 && r < sp.rows && c >= 0 && c < sp.cols) {
              const uint4* px = reinterpret_cast<const uint4*>(map_bytes + ((size_t)r * sp.cols + c) * 32);
              rec[g][0] = __ldg(px); rec[g][1] = __ldg(px + 1);
            }
          }
        }
      };
      auto store_stage = [&](uint32_t iter, const uint4 (&rec)[MMA_G][2]) {
        const uint32_t st = iter % NS, ph = (iter / NS) & 1u;
        mbar_wait(bar_empty + 8 * st, ph ^ 1u);
        unsigned char* base = sA + (size_t)st * Cfg::kABytes + (size_t)t * 4096 + (size_t)m * 16;
#pragma unroll
        for (int g = 0; g < MMA_G; g++) {
          *reinterpret_cast<uint4*>(base + (size_t)g * T * 4096) = rec[g][0];
          *reinterpret_cast<uint4*>(base + (size_t)g * T * 4096 + 2048) = rec[g][1];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * st);
      };

      uint4 ra[MMA_G][2], rb[MMA_G][2];
      load_stage(0, ra);
#pragma unroll 1
      for (int k = 0; k < K_ITERS; k += 2) {
        if (k + 1 < K_ITERS) load_stage(k + 1, rb);
        store_stage(it + k, ra);
        if (k + 2 < K_ITERS) load_stage(k + 2, ra);
        if (k + 1 < K_ITERS) store_stage(it + k + 1, rb);
      }
      it += K_ITERS;

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
  } else if (warp == 4 * T) {
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
              const uint64_t adesc = umma_desc(a0 + (g * T + tt) * 4096, 2048, 128);
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
  if (warp == 4 * T + 1) tmem_dealloc(tmem_base, 512);
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
  bp.tile_shift = 5;
  while ((((long long)(ctx->rows >> bp.tile_shift) + 1) * ((ctx->cols >> bp.tile_shift) + 1)) > 60000) bp.tile_shift++;
  bp.tiles_x = (ctx->cols >> bp.tile_shift) + 1;
  bp.n_bins = ((ctx->rows >> bp.tile_shift) + 1) * bp.tiles_x + 1;
  if (int e = ctx->bin_counts.reserve((size_t)bp.n_bins * 4)) return e;
  if (int e = ctx->perm.reserve((size_t)n_items * 4)) return e;
  TDR_CUDA(cudaMemsetAsync(ctx->bin_counts.p, 0, (size_t)bp.n_bins * 4, ctx->stream));
  const int blocks = (int)((n_items + 255) / 256 < ctx->sm_count * 8 ? (n_items + 255) / 256 : ctx->sm_count * 8);
  k_bin_count<<<blocks, 256, 0, ctx->stream>>>(bp, ctx->bin_counts.as<int>());
  k_bin_scan<<<1, 1024, 0, ctx->stream>>>(ctx->bin_counts.as<int>(), bp.n_bins);
  k_bin_scatter<<<blocks, 256, 0, ctx->stream>>>(bp, ctx->bin_counts.as<int>(), ctx->perm.as<int>());
  count_launch(ctx, 3);
  TDR_CUDA(cudaGetLastError());

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
  const long long n_batches48 = (sp.n_work + 511) / 512, n_batches112 = (sp.n_work + 255) / 256;
  if (S_pad == 48) {
    using Cfg = MmaCfg<96, 4>;
    static bool attr = false;
    if (!attr) { TDR_CUDA(cudaFuncSetAttribute(k_score_mma<96, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem)); attr = true; }
    int grid = (int)(n_batches48 < ctx->sm_count ? n_batches48 : ctx->sm_count);
    k_score_mma<96, 4><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);
  } else {
    using Cfg = MmaCfg<224, 2>;
    static bool attr = false;
    if (!attr) { TDR_CUDA(cudaFuncSetAttribute(k_score_mma<224, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem)); attr = true; }
    int grid = (int)(n_batches112 < ctx->sm_count ? n_batches112 : ctx->sm_count);
    k_score_mma<224, 2><<<grid, Cfg::kThreads, Cfg::kSmem, ctx->stream>>>(sp);
  }
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  *used = true;
  return TDR_OK;
}

}  // namespace tdr
