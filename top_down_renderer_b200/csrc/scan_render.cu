// scan_render.cu — a1/a2: semantic scan rasterisers.
//
// Reference: ScanRendererPolar::renderSemanticTopDown (src/scan_renderer_polar.cpp:83-109) and
// ScanRenderer::renderSemanticTopDown (src/scan_renderer.cpp:55-78).
//
// The images are integer histograms (the reference's float "+= 1" is exact below 2^24), so the
// result is order-independent and bit-exact as soon as the bin index is (tdr_math.cuh).
// Small images (the live 100 x 25 polar image): ATOMIC-FREE.  Every warp owns a private uint16
// histogram in shared memory; inside a warp, lanes that hit the same bin are aggregated with
// match.any so exactly one lane per distinct bin updates the warp's histogram; the per-warp
// histograms are summed per CTA into a partial in global memory and a second kernel folds the
// partials.  No atomics anywhere, bit-reproducible.
// Large images (refine_map-style batch rasterisation, BASELINE cfg5): warp-aggregated integer
// red.global (still exact, order-independent).
#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

struct ScanParams {
  const uint8_t* pts; int stride; int ioff; long long n;
  const int32_t* lut; int n_lut; int C;
  float res; float ang_res;   // polar
  int d0, d1;                 // polar: n_theta, n_r ; cart: rows, cols
  int polar;
};

__device__ __forceinline__ int scan_key(const ScanParams& sp, long long i) {
  const uint8_t* p = sp.pts + i * sp.stride;
  float x = *reinterpret_cast<const float*>(p);
  float y = *reinterpret_cast<const float*>(p + 4);
  int a, b;
  bool ok = sp.polar ? polar_bin(x, y, sp.res, sp.ang_res, sp.d0, sp.d1, &a, &b)
                     : cart_bin(x, y, sp.res, sp.d0, sp.d1, &a, &b);
  if (!ok) return -1;
  float inten = *reinterpret_cast<const float*>(p + sp.ioff);
  int cls = f2i_x86(inten);                       // int pt_class = intensity (:103)
  if (cls < 0 || cls >= sp.n_lut) return -1;      // unchecked in the reference (UB); dropped here
  int f = sp.lut[cls];
  if (f < 0 || f >= sp.C) return -1;
  // polar: img(theta_ind, r_ind) -> r*n_theta + theta ; cart: img(y_ind, x_ind) -> x*rows + y  (a = x_ind, b = y_ind)
  int cell = sp.polar ? (b * sp.d0 + a) : (a * sp.d0 + b);
  return f * (sp.d0 * sp.d1) + cell;
}

// grid.x CTAs x W warps; warp g handles points [g*per_warp, (g+1)*per_warp)
__global__ void k_scan_bin_private(ScanParams sp, int bins, long long per_warp, int32_t* __restrict__ partial) {
  extern __shared__ uint16_t sh_hist[];
  const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < W * bins; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  uint16_t* mine = sh_hist + (size_t)warp * bins;
  long long g = (long long)blockIdx.x * W + warp;
  long long lo = g * per_warp, hi = lo + per_warp;
  if (hi > sp.n) hi = sp.n;
  for (long long base = lo; base < hi; base += 32) {
    long long i = base + lane;
    int key = (i < hi) ? scan_key(sp, i) : -1;
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && lane == (__ffs(peers) - 1)) mine[key] = (uint16_t)(mine[key] + __popc(peers));
    __syncwarp();
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x) {
    int s = 0;
    for (int w = 0; w < W; w++) s += sh_hist[(size_t)w * bins + b];
    partial[(size_t)blockIdx.x * bins + b] = s;
  }
}

__global__ void k_hist_fold(const int32_t* __restrict__ partial, int nparts, int bins, int accumulate,
                            int32_t* __restrict__ hist) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  int s = accumulate ? hist[b] : 0;
  for (int p = 0; p < nparts; p++) s += partial[(size_t)p * bins + b];
  hist[b] = s;
}

__global__ void k_scan_bin_global(ScanParams sp, int32_t* __restrict__ hist) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31;
  int key = (i < sp.n) ? scan_key(sp, i) : -1;
  unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(hist + key, __popc(peers));
}

// counts -> the reference's float images ("+= 1" on a float saturates at 2^24)
__global__ void k_hist_to_float(const int32_t* __restrict__ hist, int bins, float* __restrict__ img) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bins) return;
  int v = hist[b];
  img[b] = (float)(v > 16777216 ? 16777216 : v);
}

// scan_pack[p][c] = count_c[p] * 0.01 * w_c (c < C), slot 7 = sum_c count_c[p]
// (state_particle.cpp:136-142: cost += S*0.01*w_c ; normalization += sum scan_c*known)
__global__ void k_scan_pack(const float* __restrict__ img, int P, int C, const float* __restrict__ cw,
                            float* __restrict__ pack) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float out[8];
#pragma unroll
  for (int c = 0; c < 8; c++) out[c] = 0.f;
  float tot = 0.f;
  for (int c = 0; c < C; c++) {
    float v = img[(size_t)c * P + p];
    out[c] = (float)((double)v * 0.01 * (double)cw[c]);
    tot += v;
  }
  out[7] = tot;
  float4* dst = reinterpret_cast<float4*>(pack + (size_t)p * 8);
  dst[0] = make_float4(out[0], out[1], out[2], out[3]);
  dst[1] = make_float4(out[4], out[5], out[6], out[7]);
}

// cfg5 / refine_map-style batch binning (src/refine_map.cpp:76-94): ind = floor(pt/res) + (int)(centre/res), one uint8
// counter per (class, y, x) that wraps mod 256 like the reference's uint8 "+= 1".  Integer adds commute, so a
// warp-aggregated int32 histogram followed by "& 255" is bit-exact for any point order.
__global__ void k_refine_bin(const float2* __restrict__ xy, const int32_t* __restrict__ cls, long long n, float res,
                             int off_x, int off_y, int width, int height, int C, int32_t* __restrict__ hist) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  long long key = -1;
  if (i < n) {
    const float2 p = xy[i];
    const int c = cls[i];
    const double fx = floor((double)p.x / (double)res), fy = floor((double)p.y / (double)res);
    if (fx > -2e9 && fx < 2e9 && fy > -2e9 && fy < 2e9 && c >= 0 && c < C) {
      const int ix = (int)fx + off_x, iy = (int)fy + off_y;
      if (ix >= 0 && ix < width && iy >= 0 && iy < height) key = ((long long)c * height + iy) * width + ix;
    }
  }
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(hist + key, __popc(peers));
}
__global__ void k_hist_to_u8(const int32_t* __restrict__ hist, size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t)(hist[i] & 255);
}

static const int kSmemBudget = 200 * 1024;

int scan_render(tdr_ctx* ctx, bool polar, float res, float ang_res, int d0, int d1, float* dev_img_out) {
  TDR_REQUIRE(ctx->n_lut > 0, TDR_ESTATE, "tdr_scan_set_lut has not been called");
  TDR_REQUIRE(res > 0.f && d0 > 0 && d1 > 0, TDR_EINVAL, "bad raster arguments");
  const int C = ctx->lut_classes;
  const long long cells = (long long)d0 * d1;
  TDR_REQUIRE(cells * C < (1ll << 30), TDR_EINVAL, "image too large");
  const int bins = (int)(cells * C);
  ScanParams sp;
  sp.pts = ctx->pts.as<uint8_t>(); sp.stride = ctx->pts_stride; sp.ioff = ctx->pts_ioff; sp.n = ctx->n_pts;
  sp.lut = ctx->lut.as<int32_t>(); sp.n_lut = ctx->n_lut; sp.C = C;
  sp.res = res; sp.ang_res = ang_res; sp.d0 = d0; sp.d1 = d1; sp.polar = polar ? 1 : 0;
  if (int e = ctx->hist.reserve((size_t)bins * 4)) return e;
  int W = kSmemBudget / (2 * bins);
  if (W > 8) W = 8;
  if (W >= 2 && ctx->n_pts > 0) {
    // atomic-free path
    TDR_SMEM_OPTIN(ctx, OPTIN_SCAN_BIN, k_scan_bin_private, kSmemBudget + 8192);
    long long done = 0;
    int round = 0;
    while (done < ctx->n_pts) {
      long long n_here = ctx->n_pts - done;
      // every warp handles at most 65504 points (uint16 counters), aim for >= 512 per warp
      long long warps_needed = (n_here + 511) / 512;
      int ctas = (int)((warps_needed + W - 1) / W);
      if (ctas > ctx->sm_count) ctas = ctx->sm_count;
      if (ctas < 1) ctas = 1;
      long long per_warp = (n_here + (long long)ctas * W - 1) / ((long long)ctas * W);
      per_warp = (per_warp + 31) / 32 * 32;
      if (per_warp > 65504) { per_warp = 65504; n_here = per_warp * ctas * W; }
      ScanParams s2 = sp; s2.pts = sp.pts + done * sp.stride; s2.n = n_here;
      if (int e = ctx->scratch.reserve((size_t)ctas * bins * 4)) return e;
      k_scan_bin_private<<<ctas, W * 32, (size_t)W * bins * 2, ctx->stream>>>(s2, bins, per_warp, ctx->scratch.as<int32_t>());
      k_hist_fold<<<(bins + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch.as<int32_t>(), ctas, bins, round > 0,
                                                               ctx->hist.as<int32_t>());
      count_launch(ctx, 2);
      done += n_here; round++;
    }
  } else {
    TDR_CUDA(cudaMemsetAsync(ctx->hist.p, 0, (size_t)bins * 4, ctx->stream));
    if (ctx->n_pts > 0) {
      long long blocks = (ctx->n_pts + 255) / 256;
      k_scan_bin_global<<<(unsigned)blocks, 256, 0, ctx->stream>>>(sp, ctx->hist.as<int32_t>());
      count_launch(ctx);
    }
  }
  k_hist_to_float<<<(bins + 255) / 256, 256, 0, ctx->stream>>>(ctx->hist.as<int32_t>(), bins, dev_img_out);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

// counters -> class-presence seeds for the distance-field rebuild: class c is present where its (wrapped uint8)
// counter is non-zero, i.e. the binary layer (counter == 0 ? 1 : 0) of the composition in SURVEY section 8 (cfg5);
// a pixel with no class at all is unknown (computeDists' mask, top_down_map.cpp:294-299).  Same bit layout as
// k_layers_to_seeds.
__global__ void k_hist_to_seeds(const int32_t* __restrict__ hist, size_t L, int C, uint8_t* __restrict__ seed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  uint32_t bits = 0;
  for (int c = 0; c < C; c++) if (hist[(size_t)c * L + i] & 255) bits |= 1u << c;
  if (bits == 0) bits = 0x80u;
  seed[i] = (uint8_t)bits;
}

int refine_begin(tdr_ctx* ctx, float res, float cx, float cy, int width, int height, int C) {
  TDR_REQUIRE(res > 0.f && width > 0 && height > 0 && C >= 1 && C <= 7, TDR_EINVAL, "bad refine arguments");
  const size_t cells = (size_t)C * width * height;
  if (int e = ctx->hist.reserve(cells * 4)) return e;
  TDR_CUDA(cudaMemsetAsync(ctx->hist.p, 0, cells * 4, ctx->stream));
  ctx->refine_w = width; ctx->refine_h = height; ctx->refine_C = C; ctx->refine_res = res;
  ctx->refine_off_x = (int)(cx / res); ctx->refine_off_y = (int)(cy / res);        // src/refine_map.cpp:80-81
  return TDR_OK;
}

int refine_add(tdr_ctx* ctx, const float* xy, const int32_t* cls, long long n, bool on_device) {
  TDR_REQUIRE(ctx->refine_w > 0, TDR_ESTATE, "tdr_refine_begin has not been called");
  TDR_REQUIRE(n >= 0 && (n == 0 || (xy && cls)), TDR_EINVAL, "bad refine points");
  if (n == 0) return TDR_OK;
  const float2* d_xy = reinterpret_cast<const float2*>(xy);
  const int32_t* d_cls = cls;
  int slot = -1;
  if (!on_device) {
    // H2D on the copy stream into one of two staging slots, the kernel on the context stream behind an event: the
    // copy of chunk k + 1 overlaps the binning of chunk k.  Returns once the copy is done (the caller may reuse its
    // buffers); the kernel is still running.
    if (!ctx->copy_stream) {
      TDR_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 2; k++) {
        TDR_CUDA(cudaEventCreateWithFlags(&ctx->refine_copied[k], cudaEventDisableTiming));
        TDR_CUDA(cudaEventCreateWithFlags(&ctx->refine_binned[k], cudaEventDisableTiming));
      }
    }
    slot = ctx->refine_slot; ctx->refine_slot ^= 1;
    TDR_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->refine_binned[slot], 0));     // the kernel that read this slot is done
    if (ctx->refine_stage[slot].cap < (size_t)n * 12) {
      TDR_CUDA(cudaStreamSynchronize(ctx->copy_stream));
      if (int e = ctx->refine_stage[slot].reserve((size_t)n * 12)) return e;
    }
    float2* sxy = ctx->refine_stage[slot].as<float2>();
    int32_t* scl = reinterpret_cast<int32_t*>(ctx->refine_stage[slot].as<unsigned char>() + (size_t)n * 8);
    TDR_CUDA(cudaMemcpyAsync(sxy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    TDR_CUDA(cudaMemcpyAsync(scl, cls, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    TDR_CUDA(cudaEventRecord(ctx->refine_copied[slot], ctx->copy_stream));
    TDR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->refine_copied[slot], 0));
    d_xy = sxy; d_cls = scl;
  }
  k_refine_bin<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_xy, d_cls, n, ctx->refine_res, ctx->refine_off_x, ctx->refine_off_y,
                                                                    ctx->refine_w, ctx->refine_h, ctx->refine_C, ctx->hist.as<int32_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (slot >= 0) {
    TDR_CUDA(cudaEventRecord(ctx->refine_binned[slot], ctx->stream));
    TDR_CUDA(cudaEventSynchronize(ctx->refine_copied[slot]));
  }
  return TDR_OK;
}

int refine_counts(tdr_ctx* ctx, uint8_t* maps_out) {
  TDR_REQUIRE(ctx->refine_w > 0 && maps_out, TDR_ESTATE, "no batch binning in progress / null output");
  const size_t cells = (size_t)ctx->refine_C * ctx->refine_w * ctx->refine_h;
  if (int e = ctx->scratch2.reserve(cells)) return e;
  k_hist_to_u8<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(ctx->hist.as<int32_t>(), cells, ctx->scratch2.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(maps_out, ctx->scratch2.p, cells, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int refine_rebuild_map(tdr_ctx* ctx, float resolution) {
  TDR_REQUIRE(ctx->refine_w > 0, TDR_ESTATE, "no batch binning in progress");
  const int rows = ctx->refine_h, cols = ctx->refine_w, C = ctx->refine_C;
  const size_t L = (size_t)rows * cols;
  if (int e = ctx->scratch2.reserve(L)) return e;
  k_hist_to_seeds<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->hist.as<int32_t>(), L, C, ctx->scratch2.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return map_from_seeds(ctx, rows, cols, C, resolution);     // copies the seeds into the map's own buffer, then the EDT
}

int refine_bin(tdr_ctx* ctx, const float* xy, const int32_t* cls, long long n, float res, float cx, float cy, int width,
               int height, int C, uint8_t* maps_out) {
  TDR_REQUIRE(n >= 0 && maps_out, TDR_EINVAL, "bad refine_bin arguments");
  if (int e = refine_begin(ctx, res, cx, cy, width, height, C)) return e;
  if (int e = refine_add(ctx, xy, cls, n, false)) return e;
  return refine_counts(ctx, maps_out);
}

int scan_pack(tdr_ctx* ctx) {
  TDR_REQUIRE(ctx->have_scan && ctx->have_params, TDR_ESTATE, "scan images / filter params missing");
  int P = ctx->scan_theta * ctx->scan_r;
  if (int e = ctx->scan_pack.reserve((size_t)P * 8 * 4)) return e;
  k_scan_pack<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->scan_img.as<float>(), P, ctx->scan_C,
                                                        ctx->d_cw.as<float>(), ctx->scan_pack.as<float>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

}  // namespace tdr
