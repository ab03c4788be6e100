// gather_bench2.cu — round-2 micro-benchmarks behind the cfg3 score-kernel redesign:
//   A. how global-load wavefronts depend on WHICH lanes share a 128-byte line (adjacent lanes / aligned groups /
//      lanes 16 apart), for 32-, 16- and 8-byte records (LDG.256 / LDG.128 / LDG.64) out of an L2-resident window;
//   B. random shared-memory gathers of 16- / 32-byte records (LDS.128, one or two per record) out of a staged region,
//      alone and while cp.async.bulk keeps refilling the region from L2 (the "TMA-staged map tile" design);
//   C. plain cp.async.bulk global->shared throughput per SM (the staging path's ceiling).
// Prints records / clk / SM (A, B) or bytes / clk / SM (C).   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------------------
// A. byte offset (multiple of REC) of the record lane `lane` of warp `warp` reads in iteration `it`.
//    window = 32 MB (L2-resident).  share = lanes per 128-byte line; how = which lanes share:
//      0: aligned groups of `share` consecutive lanes      1: lanes l, l + 32/share, ... (strided)      2: random pairing
//    within the line every lane takes its own REC-byte slot (distinct sectors where REC >= 32).
template <int REC>
__global__ void __launch_bounds__(512) k_ldg(const unsigned char* __restrict__ base, int share, int how, int iters,
                                             uint32_t* __restrict__ sink) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = gid >> 5;
  const uint32_t n_lines = (32u << 20) / 128u, per_line = 128 / REC;
  // lane -> (group, slot) once (the loop must stay load-bound, not ALU-bound)
  uint32_t grp, slot;
  if (how == 0) { grp = lane / share; slot = lane % share; }
  else if (how == 1) { const uint32_t ng = 32 / share; grp = lane % ng; slot = lane / ng; }
  else { const uint32_t perm = (lane * 13u) & 31u; grp = perm / share; slot = perm % share; }
  const uint32_t slot_off = (slot % per_line) * REC;
  uint32_t q = hash32(warp * 131u + grp * 977u + 1u);          // same stream for every lane of a group
  uint32_t acc = 0;
#pragma unroll 4
  for (int it = 0; it < iters; it++) {
    q = q * 1664525u + 1013904223u;
    const unsigned char* p = base + (size_t)(__umulhi(q, n_lines) * 128u + (REC >= 32 ? ((q >> 3) & 3u) * 32u * (share == 1) : 0u) + slot_off);
    if (REC == 32) {
      uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
      asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "l"(p));
      acc ^= v0 ^ v3 ^ v5 ^ v6;
    } else if (REC == 16) {
      uint32_t v0, v1, v2, v3;
      asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(p));
      acc ^= v0 ^ v3;
    } else {
      uint32_t v0, v1;
      asm volatile("ld.global.nc.v2.b32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "l"(p));
      acc ^= v0 ^ v1;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// B. shared-memory gathers.  Region = REGION bytes of dynamic smem; each thread reads a pseudo-random record per
//    iteration (spread = how far apart the 8 lanes of a quarter warp may be, in records: small = spatially sorted
//    particles).  refill: warp 0's elected lane keeps issuing cp.async.bulk of CHUNK bytes into a second buffer.
template <int REC>
__global__ void __launch_bounds__(544) k_lds(const unsigned char* __restrict__ gsrc, int region_bytes, int spread, int iters, int refill,
                                             int swz, uint32_t* __restrict__ sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ int done_flag;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_rec = region_bytes / REC;
  for (uint32_t i = tid; i < (uint32_t)region_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  if (tid == 0) {
    done_flag = 0;
    for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  uint32_t acc = 0;
  if (refill && warp == 16) {
    // a 17th warp keeps `refill` CHUNK-byte bulk copies in flight into the second buffer until the gather warps finish
    const uint32_t CHUNK = 8192;
    uint32_t issued = 0;
    const uint32_t dst0 = smem_u32(smem + region_bytes);
    if (lane == 0) {
      while (*reinterpret_cast<volatile int*>(&done_flag) < 16) {
        const uint32_t s = issued % (uint32_t)refill, b = smem_u32(&bars[s]);
        if (issued >= (uint32_t)refill) {
          const uint32_t par = ((issued / refill) - 1) & 1;
          asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(par));
        }
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(CHUNK));
        const unsigned char* src = gsrc + (size_t)(hash32(blockIdx.x * 7u + issued) % 4000u) * CHUNK;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0 + s * CHUNK),
                     "l"(src), "r"(CHUNK), "r"(b));
        issued++;
      }
      // drain
      for (uint32_t k = issued > (uint32_t)refill ? issued - refill : 0; k < issued; k++) {
        const uint32_t s = k % (uint32_t)refill, b = smem_u32(&bars[s]), par = (k / refill) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(par));
      }
      atomicAdd(reinterpret_cast<unsigned long long*>(sink + 2), (unsigned long long)issued);
    }
  } else if (warp < 16) {
    const uint32_t s_base = smem_u32(smem);
    // cheap address streams (the loop must stay LDS-bound, not ALU-bound): one LCG per quarter warp for the anchor,
    // one per lane for the offset inside the neighbourhood
    uint32_t qa = hash32((blockIdx.x * 64 + warp * 4 + (lane >> 3)) * 131u + 7u), ql = hash32(tid * 9781u + blockIdx.x);
    const uint32_t span = n_rec - spread;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
      qa = qa * 1664525u + 1013904223u; ql = ql * 22695477u + 1u;
      const uint32_t r = __umulhi(qa, span) + __umulhi(ql, (uint32_t)spread);
      uint32_t off = r * REC;
      uint32_t v0, v1, v2, v3;
      if (REC == 32) {
        const uint32_t x = swz ? ((off >> 7) & 1u) << 4 : 0u;          // SWIZZLE_32B: 16-byte halves swap on address bit 7
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(s_base + (off ^ x)));
        acc ^= v0 ^ v3;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(s_base + ((off + 16) ^ x)));
        acc ^= v1 ^ v2;
      } else if (REC == 16) {
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(s_base + off));
        acc ^= v0 ^ v3;
      } else {
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(s_base + off));
        acc ^= v0 ^ v1;
      }
    }
    __syncwarp();
    if (lane == 0) atomicAdd(&done_flag, 1);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// C. bulk-copy throughput: every CTA streams CHUNK-byte copies from an L2-resident (or DRAM) source into smem
__global__ void __launch_bounds__(128) k_bulk(const unsigned char* __restrict__ gsrc, uint32_t chunk, uint32_t window_chunks, int n_copies, int depth) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    for (int c = 0; c < n_copies + depth; c++) {
      const int s = c % depth;
      const uint32_t b = smem_u32(&bars[s]);
      if (c >= depth) {
        const uint32_t par = ((c / depth) - 1) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(par));
      }
      if (c < n_copies) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(chunk));
        const unsigned char* src = gsrc + (size_t)(hash32(blockIdx.x * 977u + c) % window_chunks) * chunk;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + s * chunk),
                     "l"(src), "r"(chunk), "r"(b));
      }
    }
  }
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) { float ms; cudaEventElapsedTime(&ms, e0, e1); return ms; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, %d kHz\n", prop.name, sms, clock_khz);
  unsigned char* d; uint32_t* sink;
  const size_t bytes = 512u << 20;
  CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes)); CK(cudaMalloc(&sink, 64));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048, threads = 512;

  printf("== A. global loads from a 32 MB window: lanes per 128-byte line (share), which lanes (how: 0 aligned groups, 1 strided, 2 scrambled)\n");
#define RUN_LDG(REC)                                                                                                     \
  for (int share = 1; share <= 128 / REC && share <= 32; share *= 2)                                                   \
    for (int how = 0; how < (share == 1 ? 1 : 3); how++)                                                                \
      for (int cps = 2; cps <= 4; cps += 2) {                                                                           \
        k_ldg<REC><<<sms * cps, threads>>>(d, share, how, 64, sink);                                                    \
        CK(cudaDeviceSynchronize());                                                                                    \
        cudaEventRecord(e0);                                                                                            \
        k_ldg<REC><<<sms * cps, threads>>>(d, share, how, iters, sink);                                                 \
        cudaEventRecord(e1);                                                                                            \
        CK(cudaDeviceSynchronize());                                                                                    \
        const float ms = time_ms(e0, e1);                                                                               \
        const double recs = (double)sms * cps * threads * iters, clk = ms * 1e-3 * clock_khz * 1e3;                     \
        printf("LDG rec %2d B  share %2d how %d ctas/sm %d : %7.3f ms  %6.3f rec/clk/SM  %6.1f B/clk/SM\n", REC, share, how, cps, ms, \
               recs / clk / sms, recs * REC / clk / sms);                                                               \
      }
  RUN_LDG(32) RUN_LDG(16) RUN_LDG(8)

  printf("== B. shared-memory gathers (16 gather warps per CTA, 1 CTA/SM; refill = 8 KB bulk copies per 64 iterations)\n");
  const int region = 64 * 1024;
#define RUN_LDS(REC, SWZ)                                                                                               \
  {                                                                                                                     \
    CK(cudaFuncSetAttribute(k_lds<REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, region + 4 * 8192));                \
    for (int spread : {16, 64, 2048})                                                                                   \
      for (int refill : {0, 2, 4}) {                                                                                   \
        k_lds<REC><<<sms, 544, region + 4 * 8192>>>(d, region, spread, 64, refill, SWZ, sink);                           \
        CK(cudaDeviceSynchronize());                                                                                    \
        CK(cudaMemset(sink, 0, 64));                                                                                    \
        cudaEventRecord(e0);                                                                                            \
        k_lds<REC><<<sms, 544, region + 4 * 8192>>>(d, region, spread, iters * 4, refill, SWZ, sink);                    \
        CK(cudaGetLastError());                                                                                         \
        cudaEventRecord(e1);                                                                                            \
        CK(cudaDeviceSynchronize());                                                                                    \
        const float ms = time_ms(e0, e1);                                                                               \
        const double recs = (double)sms * 512 * iters * 4, clk = ms * 1e-3 * clock_khz * 1e3;                           \
        unsigned long long issued = 0;                                                                                  \
        CK(cudaMemcpy(&issued, sink + 2, 8, cudaMemcpyDeviceToHost));                                                    \
        printf("LDS rec %2d B swz %d spread %4d bulk depth %d : %7.3f ms  %6.3f rec/clk/SM  (+ %5.1f bulk B/clk/SM)\n", REC, SWZ, spread, \
               refill, ms, recs / clk / sms, (double)issued * 8192 / clk / sms);                                        \
      }                                                                                                                 \
  }
  RUN_LDS(16, 0) RUN_LDS(32, 0) RUN_LDS(32, 1) RUN_LDS(8, 0)

  printf("== C. cp.async.bulk global->shared, one issuing thread per CTA\n");
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  for (uint32_t chunk : {2048u, 8192u, 16384u})
    for (int depth : {2, 4})
      for (int cps = 1; cps <= 2; cps++)
        for (int dram = 0; dram < 2; dram++) {
          const uint32_t window = dram ? (uint32_t)(bytes / chunk) : (32u << 20) / chunk;
          const int n_copies = (int)((64u << 20) / chunk / 8);
          k_bulk<<<sms * cps, 128, depth * chunk>>>(d, chunk, window, 64, depth);
          CK(cudaDeviceSynchronize());
          cudaEventRecord(e0);
          k_bulk<<<sms * cps, 128, depth * chunk>>>(d, chunk, window, n_copies, depth);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          const float ms = time_ms(e0, e1);
          const double clk = ms * 1e-3 * clock_khz * 1e3;
          printf("bulk chunk %5u depth %d ctas/sm %d %s : %7.3f ms  %6.1f B/clk/SM  %7.1f GB/s\n", chunk, depth, cps, dram ? "dram" : "l2  ", ms,
                 (double)n_copies * chunk * cps / clk, (double)n_copies * chunk * cps * sms / (ms * 1e-3) / 1e9);
        }
  return 0;
}
