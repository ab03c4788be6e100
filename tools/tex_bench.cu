// tex_bench.cu — can the texture path deliver scattered 16-byte records faster than the LSU's one data wavefront per
// sector?  uint4 texels from a 4096 x 4096 cudaArray (block-linear) by tex2D point fetches vs ld.global from a
// row-major copy; clustered (lanes within a few px of a per-warp anchor) and random coordinates; LSU and TEX mixed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// mode 0: tex only, 1: lsu (16 B, LDG.128) only, 2: alternate per iteration, 3: lsu 32 B pair loads (LDG.256)
__global__ void __launch_bounds__(512) k(cudaTextureObject_t tex, const uint4* __restrict__ lin, int mode, int spread, int window, int iters, uint32_t* sink) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, warp = gid >> 5;
  uint32_t qa = hash32(warp * 131u + 7u), ql = hash32(gid * 9781u + 3u);
  uint32_t acc = 0;
#pragma unroll 4
  for (int it = 0; it < iters; it++) {
    qa = qa * 1664525u + 1013904223u; ql = ql * 22695477u + 1u;
    const uint32_t ax = __umulhi(qa, (uint32_t)window), ay = __umulhi(qa * 2654435761u, (uint32_t)window);
    const uint32_t x = ax + __umulhi(ql, (uint32_t)spread), y = ay + __umulhi(ql * 40503u, (uint32_t)spread);
    if (mode == 0 || (mode == 2 && (it & 1))) {
      const uint4 v = tex2D<uint4>(tex, (float)x, (float)y);
      acc ^= v.x ^ v.w;
    } else if (mode == 3) {
      uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
      asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7)
                   : "l"(lin + (((size_t)y * 4096 + x) & ~(size_t)1)));
      acc ^= v0 ^ v7;
    } else {
      const uint4 v = __ldg(lin + (size_t)y * 4096 + x);
      acc ^= v.x ^ v.w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount;
  const int W = 4096, H = 4096;
  cudaChannelFormatDesc fd = cudaCreateChannelDesc<uint4>();
  cudaArray_t arr; CK(cudaMallocArray(&arr, &fd, W, H));
  uint4* lin; CK(cudaMalloc(&lin, (size_t)W * H * 16)); CK(cudaMemset(lin, 1, (size_t)W * H * 16));
  CK(cudaMemcpy2DToArray(arr, 0, 0, lin, (size_t)W * 16, (size_t)W * 16, H, cudaMemcpyDeviceToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  uint32_t* sink; CK(cudaMalloc(&sink, 64));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048, threads = 512;
  printf("%s, %d SMs\n", prop.name, sms);
  const char* names[] = {"tex2D uint4", "LDG.128 row-major", "alternating tex / LDG.128", "LDG.256 pair row-major"};
  for (int mode = 0; mode < 4; mode++)
    for (int window : {1024, 4000})           // anchors within a 1024^2 (16 MB: L2) or the whole 4000^2 array (256 MB)
      for (int spread : {4, 8, 32})           // lanes of a warp within spread x spread px of the anchor
        for (int cps = 2; cps <= 4; cps += 2) {
          k<<<sms * cps, threads>>>(tex, lin, mode, spread, window - spread, 64, sink);
          CK(cudaDeviceSynchronize());
          cudaEventRecord(e0);
          k<<<sms * cps, threads>>>(tex, lin, mode, spread, window - spread, iters, sink);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          const double recs = (double)sms * cps * threads * iters, clk = ms * 1e-3 * clock_khz * 1e3;
          printf("%-26s window %4d spread %2d ctas/sm %d : %7.3f ms  %6.3f rec/clk/SM\n", names[mode], window, spread, cps, ms, recs / clk / sms);
        }
  return 0;
}
