// gather_bench.cu — micro-benchmark: how fast can one SM pull scattered 32-byte records (one L2 sector each)?
// Variants: per-thread 2xLDG.128, lane pairs (l,l+16) / (2j,2j+1), 256-bit loads, cp.async (LDGSTS) 16 B x2,
// cp.async.bulk 32 B.  Patterns: random over 512 MB (DRAM), random over 32 MB (L2), neighbours (4 per line),
// near (distinct lines, same 4 KB); 4-7: L2-resident window with 32 / 32 / 16 / 8 lines per warp load.  Prints records / clk / SM.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// record index for (thread gid, iteration it)
__device__ __forceinline__ uint32_t rec_index(int pattern, uint32_t gid, uint32_t it, uint32_t n_rec) {
  uint32_t warp = gid >> 5, lane = gid & 31;
  switch (pattern) {
    case 0: return hash32(gid * 9781u + it * 6271u) % n_rec;                          // random, whole array
    case 1: return hash32(gid * 9781u + it * 6271u) % (1u << 20);                     // random, 32 MB window
    case 2: return (hash32(warp * 131u + it * 7919u) % (n_rec - 64)) + lane;          // 32 consecutive records (8 lines)
    case 3: return (hash32(warp * 131u + it * 7919u) % (n_rec - 4096)) + lane * 5;    // distinct lines, close by
    case 4: return (hash32(warp * 131u + it * 7919u) % ((1u << 20) - 256)) + lane * 4;   // L2 window, 32 consecutive lines, 1 sector each (grid rows)
    case 5: return (hash32(warp * 131u + it * 7919u) % ((1u << 20) - 256)) + lane * 4 + (hash32(gid + it) & 3u);   // same, random sector in the line
    case 6: return (hash32(warp * 131u + it * 7919u) % ((1u << 20) - 256)) + lane * 2;   // L2 window, 16 lines x 2 sectors
    default: return (hash32(warp * 131u + it * 7919u) % ((1u << 20) - 256)) + lane;      // L2 window, 8 full lines
  }
}

template <int VARIANT>
__global__ void __launch_bounds__(512) k_gather(const uint4* __restrict__ recs, uint32_t n_rec, int pattern, int iters,
                                                uint32_t* __restrict__ sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  uint32_t acc = 0;
  if (VARIANT == 0) {          // each thread: both halves of its own record
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
      uint32_t r = rec_index(pattern, gid, it, n_rec);
      uint4 a = __ldg(recs + 2 * (size_t)r), b = __ldg(recs + 2 * (size_t)r + 1);
      acc ^= a.x ^ a.w ^ b.y ^ b.z;
    }
  } else if (VARIANT == 1 || VARIANT == 2) {   // lane pairs: 1 = (l, l+16), 2 = (2j, 2j+1); two loads cover 32 records
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
      uint32_t r = rec_index(pattern, gid, it, n_rec);
      uint32_t half = VARIANT == 1 ? lane >> 4 : lane & 1, sub = VARIANT == 1 ? lane & 15 : lane >> 1;
      uint32_t r0 = __shfl_sync(0xffffffffu, r, sub), r1 = __shfl_sync(0xffffffffu, r, sub + 16);
      uint4 a = __ldg(recs + 2 * (size_t)r0 + half), b = __ldg(recs + 2 * (size_t)r1 + half);
      acc ^= a.x ^ a.w ^ b.y ^ b.z;
    }
  } else if (VARIANT == 3) {   // 256-bit load
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
      uint32_t r = rec_index(pattern, gid, it, n_rec);
      uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
      asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7)
                   : "l"(recs + 2 * (size_t)r));
      acc ^= v0 ^ v3 ^ v5 ^ v6;
    }
  } else if (VARIANT == 4) {   // cp.async 16 B x2 into smem (LDGSTS), groups of 4 iterations
    unsigned char* mine = smem + threadIdx.x * 32;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(mine);
    for (int it = 0; it < iters; it += 4) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        uint32_t r = rec_index(pattern, gid, it + u, n_rec);
        const uint4* src = recs + 2 * (size_t)r;
        uint32_t d = dst + u * 512 * 32;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(src + 1));
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 1;");
    }
    asm volatile("cp.async.wait_group 0;");
    acc ^= *reinterpret_cast<uint32_t*>(mine);
  } else if (VARIANT == 5) {   // cp.async.bulk 32 B per thread, one mbarrier per (warp, slot)
    __shared__ __align__(8) uint64_t bars[16 * 4];
    const uint32_t warp = threadIdx.x >> 5;
    unsigned char* mine = smem + threadIdx.x * 32;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(mine);
    if (lane == 0)
      for (int u = 0; u < 4; u++) {
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bars[warp * 4 + u]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
      }
    asm volatile("fence.mbarrier_init.release.cluster;");
    __syncthreads();
    for (int it = 0; it < iters; it += 4) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        uint32_t b = (uint32_t)__cvta_generic_to_shared(&bars[warp * 4 + u]);
        if (it > 0) {   // wait for the previous use of this slot
          uint32_t par = ((it / 4) - 1) & 1;
          asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(par));
        }
        if (lane == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(32 * 32));
        __syncwarp();
        uint32_t r = rec_index(pattern, gid, it + u, n_rec);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + u * 512 * 32),
                     "l"(recs + 2 * (size_t)r), "r"(32), "r"(b));
      }
    }
    for (int u = 0; u < 4; u++) {
      uint32_t b = (uint32_t)__cvta_generic_to_shared(&bars[warp * 4 + u]);
      uint32_t par = ((iters / 4) - 1) & 1;
      asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(par));
    }
    acc ^= *reinterpret_cast<uint32_t*>(mine);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int V>
static int run(const char* name, const uint4* d, uint32_t n_rec, uint32_t* sink, int sms, int clock_khz) {
  const int iters = 2048, threads = 512;
  size_t smem = (V >= 4) ? 4 * 512 * 32 : 0;
  CK(cudaFuncSetAttribute(k_gather<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int pattern = (V == 3 ? 0 : 8); pattern < 8; pattern++) {
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ctas_per_sm++) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      k_gather<V><<<sms * ctas_per_sm, threads, smem>>>(d, n_rec, pattern, 64, sink);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      k_gather<V><<<sms * ctas_per_sm, threads, smem>>>(d, n_rec, pattern, iters, sink);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double recs = (double)sms * ctas_per_sm * threads * iters;
      double clk = ms * 1e-3 * clock_khz * 1e3;
      printf("%-22s pattern %d ctas/sm %d : %8.3f ms  %6.3f rec/clk/SM  %7.1f GB/s\n", name, pattern, ctas_per_sm, ms,
             recs / clk / sms, recs * 32 / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, %d kHz\n", prop.name, prop.multiProcessorCount, clock_khz);
  const uint32_t n_rec = 16u << 20;          // 16M records x 32 B = 512 MB
  uint4* d; uint32_t* sink;
  CK(cudaMalloc(&d, (size_t)n_rec * 32)); CK(cudaMemset(d, 1, (size_t)n_rec * 32)); CK(cudaMalloc(&sink, 64));
  int sms = prop.multiProcessorCount;
  run<0>("2xLDG.128 own record", d, n_rec, sink, sms, clock_khz);
  run<1>("pair (l, l+16)", d, n_rec, sink, sms, clock_khz);
  run<2>("pair (2j, 2j+1)", d, n_rec, sink, sms, clock_khz);
  run<3>("ld.v8.b32 (256 bit)", d, n_rec, sink, sms, clock_khz);
  run<4>("cp.async 16Bx2", d, n_rec, sink, sms, clock_khz);
  run<5>("cp.async.bulk 32B", d, n_rec, sink, sms, clock_khz);
  return 0;
}
