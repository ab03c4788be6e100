// mma_common.cuh — pieces shared by the two tcgen05 score kernels (score_mma.cu: sliding scan ring, all shifts;
// score_mma_list.cu: streamed per-cell operand, a short candidate list): PTX wrappers, the fp16 hi/lo map copy,
// the spatial binning of the hypotheses and the constant-memory mirror of the polar table.
#pragma once
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
// one lane of a CONVERGED warp (the lowest active one, the same every time).  The single-thread tcgen05 / bulk-copy
// instructions are issued under this predicate from warp-uniform control flow: their operands then live in uniform
// registers.  Issuing them from inside `if (lane == 0) { loop }` makes ptxas wrap every one in an elect / branch
// waterfall, ~150 slow instructions per pipeline stage — it bounded both tensor-core kernels (ncu source view).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global bulk store (async proxy; dst may be a peer GPU's memory mapped over NVLink), tracked by bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// the same with an L2 cache policy (createpolicy): streamed results should not push the map out of L2
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes complete
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
// A operand from tensor memory (128 lanes x K/2 32-bit columns), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one pipeline stage of the tensor-memory path in ONE asm block: G MMAs (M = 128, K = 16 each) whose A operands sit 8
// columns apart in tensor memory and whose B windows start `b_step` descriptor units (16 B) apart.  Keeping the
// address arithmetic inside the block leaves ptxas nothing to re-derive per MMA (the issuing lane is the critical
// path of the kernel: every instruction between two UTCHMMA costs a full uniform-datapath latency).
template <int G>
__device__ __forceinline__ void umma_stage_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc_first,
                                              uint32_t b_step) {
  static_assert(G == 2 || G == 4, "cells per stage");
  if (G == 2) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 a1;\n\t"
        ".reg .b64 b1, bs;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "cvt.u64.u32 bs, %5;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "add.u32 a1, %1, 8;\n\t"
        "add.u64 b1, %2, bs;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc_first), "r"(b_step)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b32 a1, a2, a3;\n\t"
        ".reg .b64 b1, b2, b3, bs;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "cvt.u64.u32 bs, %5;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "add.u32 a1, %1, 8;\n\t"
        "add.u64 b1, %2, bs;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n\t"
        "add.u32 a2, %1, 16;\n\t"
        "add.u64 b2, b1, bs;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n\t"
        "add.u32 a3, %1, 24;\n\t"
        "add.u64 b3, b2, bs;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc_first), "r"(b_step)
        : "memory");
  }
}
// registers -> tensor memory: this thread's lane, 8 consecutive 32-bit columns (warp-collective)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x), "r"(a.y),
               "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t v0, uint32_t v1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(v0), "r"(v1) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// byte offset of pixel (r, c) in the fp16 map copy; ph_log2 > 0: row r is stored as 2^ph_log2 phase rows of ph_cols records
__device__ __forceinline__ size_t map16_offset(int r, int c, int cols, int ph_log2, int ph_cols) {
  if (ph_log2 == 0) return ((size_t)r * cols + c) * 32;
  return ((((size_t)r << ph_log2) + (size_t)(c & ((1 << ph_log2) - 1))) * ph_cols + (size_t)(c >> ph_log2)) * 32;
}
// Everything a gather thread needs to turn a lattice coordinate pair into a record of the fp16 map copy.
struct GatherGeom {
  int cols, ph_log2, ph_cols;      // layout (map16_offset)
  uint32_t zero_rec;               // index of an all-zero record past the map: what cells off the map read
  float row_hi, col_hi;            // rows - 0.5, cols - 0.5 (exact in fp32)
};
// record index of the lattice point (vy, vx) = ((tab * scale) * res + centre / resolution), or zero_rec off the map.
// Same result as f2i_x86(round_half_away(v)) followed by 0 <= index < limit (top_down_map_polar.cpp:28-37), in a third
// of the instructions (the gather threads are issue-bound): round_half_away(v) lies in [0, limit) exactly when
// -0.5 < v < limit - 0.5 (NaN fails both), and inside that interval it is trunc(v) + (v - trunc(v) >= 0.5).
__device__ __forceinline__ uint32_t lattice_record(float vy, float vx, const GatherGeom& g) {
  // lattice_coord (tdr_math.cuh) for both coordinates with ONE select at the end: the rounded values are formed
  // unconditionally (garbage off the map, never used)
  const bool ok = lattice_in(vy, g.row_hi) && lattice_in(vx, g.col_hi);
  const uint32_t ur = (uint32_t)lattice_round(vy), uc = (uint32_t)lattice_round(vx);
  const uint32_t rec = ((ur << g.ph_log2) + (uc & ((1u << g.ph_log2) - 1u))) * (uint32_t)g.ph_cols + (uc >> g.ph_log2);
  return ok ? rec : g.zero_rec;
}
static __global__ void k_scale_tab(const float2* __restrict__ tab, int P, float scale, float res, float2* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) out[p] = make_float2(TDR_FMUL(TDR_FMUL(tab[p].x, scale), res), TDR_FMUL(TDR_FMUL(tab[p].y, scale), res));
}
static __global__ void k_build_map16(const MapPixel* __restrict__ map, size_t n, int C, const float* __restrict__ cw,
                              uint4* __restrict__ out, int cols, int ph_log2, int ph_cols) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t o = map16_offset((int)(i / cols), (int)(i % cols), cols, ph_log2, ph_cols) / 16;
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
  out[o] = h; out[o + 1] = l;
}

// ------------------------------------------------------------------------------------------------
// spatial binning of the hypotheses (L2 locality): counting sort by coarse map tile
// ------------------------------------------------------------------------------------------------
struct BinParams {
  const float *init_x, *init_y, *dx, *dy, *scale; const uint8_t* have_init;   // particle mode
  const float* centers;                                                       // grid mode
  long long n; float resolution; int rows, cols, st_shift, seg_shift, super_x, per_super, n_bins, morton;
  int select_init;          // particle mode: bin the particles WITH a heading (tracking) instead of those without (search)
  const float* theta; int n_theta, shift_lo, shift_hi;   // shift_hi > 0: only particles whose heading maps to a row shift in [lo, hi)
};
__device__ __forceinline__ uint32_t spread_bits16(uint32_t v) {      // abcd -> 0a0b0c0d
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu; v = (v | (v << 4)) & 0x0f0f0f0fu; v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
  return v;
}
__device__ __forceinline__ int bin_of(const BinParams& b, long long i) {
  float x, y;
  if (b.centers) { x = b.centers[2 * i]; y = b.centers[2 * i + 1]; }
  else {
    if ((b.have_init[i] != 0) != (b.select_init != 0)) return -1;     // the other kind of particle: another launch
    if (b.shift_hi > 0) { const int sh = rot_to_shift(b.theta[i], b.n_theta); if (sh < b.shift_lo || sh >= b.shift_hi) return -1; }
    float s = b.scale[i];
    x = TDR_FADD(TDR_FMUL(b.dx[i], s), b.init_x[i]); y = TDR_FADD(TDR_FMUL(b.dy[i], s), b.init_y[i]);
  }
  int c = f2i_x86(TDR_FDIV(x, b.resolution)), r = f2i_x86(TDR_FDIV(y, b.resolution));
  if (c < 0 || r < 0 || c >= b.cols || r >= b.rows) return b.n_bins - 1;      // off-map: last bin
  // super-tile (2^st_shift px square, row-major over the map), then pixel row, then 32-px column segment
  const int S = 1 << b.st_shift, m = S - 1;
  const int sup = (r >> b.st_shift) * b.super_x + (c >> b.st_shift);
  // Morton (Z) order over 2 x 2-px cells: warp neighbours form compact 2-D clusters — with 4 x 2-px lines (the 16-byte
  // map copy) fewer distinct lines per warp load than runs along one pixel row (12.1 against 14.6 simulated on cfg3)
  if (b.morton) return sup * b.per_super + (int)(spread_bits16((uint32_t)(c & m) >> 1) | (spread_bits16((uint32_t)(r & m) >> 1) << 1));
  return sup * b.per_super + (r & m) * (S >> b.seg_shift) + ((c & m) >> b.seg_shift);
}
static __global__ void k_bin_count(BinParams b, int* __restrict__ counts) {
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
static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_local(int* __restrict__ a, int n, int* __restrict__ sums) {
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
static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int* __restrict__ sums, int nb) {
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
static __global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(int* __restrict__ a, int n, const int* __restrict__ sums) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  const int add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) a[base + k] += add;
}
static __global__ void k_bin_scatter(BinParams b, int* __restrict__ cursor, int* __restrict__ perm) {
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
static const int MMA_MAX_EXACT_COUNT = 2048;    // fp16 represents every integer up to here
__constant__ float2 c_tab[MMA_TAB_MAX];
// c_tab exists once per device (and per translation unit that includes this header).  g_tab_on_device[d] = the
// process-unique id (tdr_ctx::tab_id) of the table this unit's copy on device d mirrors; 0 = none / a scaled grid table.
// Contexts on different devices, several contexts on one device and recycled table pointers all key correctly.
// (Two contexts that share a device must not SCORE concurrently from two threads: they share this constant bank.)
static const int MMA_MAX_DEVICES = 64;
static uint64_t g_tab_on_device[MMA_MAX_DEVICES] = {};

static const int MMA_G = 2;        // lattice cells per pipeline stage
// A tile (128 hypotheses x 16 fp16, K-major, no swizzle): K chunk 0 at [0, 2048), K chunk 1 at [A_LBO, A_LBO + 2048).
static const int A_LBO = 2048 + 64;
static const int A_TILE = 4224;

// ph_log2 = 0: the plain row-major copy (particles); > 0: the phase-split copy for a lattice of centres
static int build_map16(tdr_ctx* ctx, int ph_log2, const uint4** out, GatherGeom* geom) {
  const size_t L = (size_t)ctx->rows * ctx->cols;
  const int ph_cols = (ctx->cols + (1 << ph_log2) - 1) >> ph_log2;
  const size_t n_rec = ((size_t)ctx->rows << ph_log2) * (size_t)ph_cols;
  TDR_REQUIRE(n_rec < (1ull << 31), TDR_EUNSUPPORTED, "map too large for the fp16 copy (%zu records)", n_rec);
  geom->cols = ctx->cols; geom->ph_log2 = ph_log2; geom->ph_cols = ph_cols; geom->zero_rec = (uint32_t)n_rec;
  geom->row_hi = (float)ctx->rows - 0.5f; geom->col_hi = (float)ctx->cols - 0.5f;
  tdr::DevBuf& buf = ph_log2 ? ctx->map16g : ctx->map16;
  *out = nullptr;
  const bool valid = ph_log2 ? ctx->map16g_log2 == ph_log2 : ctx->map16_valid;
  if (!valid) {
    const size_t bytes = (n_rec + 1) * 32;                                    // + the all-zero record
    if (int e = buf.reserve(bytes)) return e;
    if (ph_log2) TDR_CUDA(cudaMemsetAsync(buf.p, 0, bytes, ctx->stream));     // padding records past the last column
    else TDR_CUDA(cudaMemsetAsync(buf.as<unsigned char>() + n_rec * 32, 0, 32, ctx->stream));
    k_build_map16<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->map_px.as<MapPixel>(), L, ctx->C, ctx->d_cw.as<float>(),
                                                                         buf.as<uint4>(), ctx->cols, ph_log2, ph_cols);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    if (ph_log2) ctx->map16g_log2 = ph_log2; else ctx->map16_valid = true;
  }
  *out = buf.as<uint4>();
  return TDR_OK;
}

// refresh this translation unit's constant-memory mirror of the polar table when it is stale
// grid launches (one scale for every centre) mirror (tab * scale) * res instead: two FMULs less per coordinate
static int sync_const_tab_scaled(tdr_ctx* ctx, int P, float scale, float res) {
  if (int e = ctx->tab_scaled.reserve((size_t)P * 8)) return e;
  k_scale_tab<<<(P + 255) / 256, 256, 0, ctx->stream>>>(ctx->tab.as<float2>(), P, scale, res, ctx->tab_scaled.as<float2>());
  count_launch(ctx);
  TDR_CUDA(cudaMemcpyToSymbolAsync(c_tab, ctx->tab_scaled.p, (size_t)P * 8, 0, cudaMemcpyDeviceToDevice, ctx->stream));
  g_tab_on_device[ctx->device % MMA_MAX_DEVICES] = 0;     // the next particle launch mirrors the plain table again
  return TDR_OK;
}
static int sync_const_tab(tdr_ctx* ctx, int P) {
  uint64_t& seen = g_tab_on_device[ctx->device % MMA_MAX_DEVICES];
  if (seen != ctx->tab_id || ctx->tab_id == 0) {
    TDR_CUDA(cudaMemcpyToSymbolAsync(c_tab, ctx->tab.p, (size_t)P * 8, 0, cudaMemcpyDeviceToDevice, ctx->stream));
    seen = ctx->tab_id;
  }
  return TDR_OK;
}

// spatial binning of the hypotheses -> ctx->perm (counting sort by super-tile, pixel row, column segment)
// shift_hi > 0 (tracking in passes): only the particles whose heading lies in that window of row shifts are listed; their
// number is not known on the host — *count_dev (if given) receives the device word that holds it once the scatter ran
static int build_perm(tdr_ctx* ctx, bool grid_mode, long long n_items, bool morton = false, bool select_init = false,
                      int shift_lo = 0, int shift_hi = 0, const int** count_dev = nullptr) {
  if (int e = score_i8_prepare_join(ctx)) return e;                  // never two sorts into the same buffers at once
  if (grid_mode && ctx->perm_grid_n == n_items) return TDR_OK;       // resident centres, same map: the order still holds
  ctx->perm_grid_n = -1;
  tdr::Particles& pt = ctx->part[ctx->cur];
  BinParams bp; memset(&bp, 0, sizeof(bp));
  if (grid_mode) bp.centers = ctx->grid_centers.as<float>();
  else {
    bp.init_x = pt.init_x.as<float>(); bp.init_y = pt.init_y.as<float>(); bp.dx = pt.dx.as<float>(); bp.dy = pt.dy.as<float>();
    bp.scale = pt.scale.as<float>(); bp.have_init = pt.have_init.as<uint8_t>();
  }
  bp.n = n_items; bp.resolution = ctx->resolution; bp.rows = ctx->rows; bp.cols = ctx->cols;
  // bins = (super-tile, pixel row, column segment).  The super-tile keeps the hypotheses that are in flight
  // together inside one compact region so that its dilated footprint stays L2-resident; row + segment order puts
  // warp neighbours on the same map row a few pixels apart.
  bp.st_shift = ctx->mma_st_shift;
  while ((1 << bp.st_shift) < 32) bp.st_shift++;
  bp.super_x = (ctx->cols >> bp.st_shift) + 1;
  bp.seg_shift = ctx->mma_seg_shift;
  bp.morton = morton ? 1 : 0;
  bp.select_init = select_init ? 1 : 0;
  bp.theta = pt.theta.as<float>(); bp.n_theta = ctx->n_theta; bp.shift_lo = shift_lo; bp.shift_hi = shift_hi;
  if (morton) bp.seg_shift = 2;                   // (S/2)^2 Morton cells = S * (S >> 2) bins: the same count
  if (morton && bp.st_shift > 16) bp.st_shift = 16;
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
  if (grid_mode) ctx->perm_grid_n = n_items;
  if (count_dev) *count_dev = d_counts + (bp.n_bins - 1);       // the last bin's cursor ends at the number of listed items
  return TDR_OK;
}

}  // namespace tdr
