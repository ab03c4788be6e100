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

// ------------------------------------------------------------------------------------------------
// f2: the geometric renderers (scan_renderer_polar.cpp:6-81, scan_renderer.cpp:7-53) — the last reference function of
// SURVEY section 8 without a kernel.  The cloud is organised: `width` columns x `height` rows, point (col, row) at
// row * width + col, visited column by column.
// ------------------------------------------------------------------------------------------------
static const int GEO_CAP = 8192;             // points one angular bin can hold (shared memory: keys + xyz)

// Polar: one CTA per angular bin.  (1) the bin's points in visiting order (block-wide compaction), (2) bitonic sort
// by descending range — equal ranges keep the visiting order, where std::sort (:50-52) leaves their order open —
// (3) one thread walks the sorted points with the slope rule (:58-78): the walk is a dependent chain by construction.
__global__ void __launch_bounds__(1024) k_geo_polar(const uint8_t* __restrict__ pts, int stride, int width, int height, float res,
                                                    float ang_res, int n_theta, int n_r, float* __restrict__ out,
                                                    int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);          // [GEO_CAP]
  float* xyz = reinterpret_cast<float*>(keys + GEO_CAP);                           // [GEO_CAP][3]
  float* row = xyz + 3 * GEO_CAP;                                                   // [2][n_r]
  __shared__ int s_warp[32];
  __shared__ int s_count;
  const int bin = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_count = 0;
  for (int q = tid; q < 2 * n_r; q += blockDim.x) row[q] = 0.f;
  __syncthreads();
  const long long n = (long long)width * height;
  for (long long base = 0; base < n; base += blockDim.x) {
    const long long o = base + tid;                     // visiting order: o = col * height + row
    bool mine = false;
    float x = 0.f, y = 0.f, z = 0.f, r = 0.f;
    if (o < n) {
      const int col = (int)(o / height), rw = (int)(o - (long long)col * height);
      const uint8_t* p = pts + ((size_t)rw * width + col) * stride;
      x = *reinterpret_cast<const float*>(p); y = *reinterpret_cast<const float*>(p + 4); z = *reinterpret_cast<const float*>(p + 8);
      if (!(x == 0.f && y == 0.f)) {                                                       // :30
        const float theta = fdlibm_atan2f(x, y);                                           // :32 atan2(pt.x, pt.y)
        r = TDR_FSQRT(TDR_FADD(TDR_FMUL(x, x), TDR_FMUL(y, y)));                           // :33
        float t = TDR_FADD(round_half_away(TDR_FDIV(theta, ang_res)), (float)(n_theta / 2));   // :36-37
        if (t == t) {
          t = (t < 0.f) ? 0.f : (((float)(n_theta - 1) < t) ? (float)(n_theta - 1) : t);   // std::clamp<float>
          mine = (int)t == bin;
        }
      }
    }
    // ordered append: positions by a block-wide exclusive count
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_count;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (mine) {
      if (pos < GEO_CAP) {
        keys[pos] = ((unsigned long long)(~__float_as_uint(r)) << 32) | (unsigned)pos;      // r >= 0: bits order like values
        xyz[3 * pos] = x; xyz[3 * pos + 1] = y; xyz[3 * pos + 2] = z;
      } else atomicExch(overflow, 1);
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_warp[w]; s_count += t; }
    __syncthreads();
  }
  const int count = s_count < GEO_CAP ? s_count : GEO_CAP;
  int m = 1; while (m < count) m <<= 1;
  for (int q = count + tid; q < m; q += blockDim.x) keys[q] = ~0ull;                        // padding sorts last
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = tid; q < m; q += blockDim.x) {
        const int l = q ^ j;
        if (l > q) {
          const unsigned long long a = keys[q], b = keys[l];
          const bool up = (q & k) == 0;
          if ((a > b) == up) { keys[q] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  if (tid == 0) {
    float lx = 0.f, ly = 0.f, lz = 0.f;                                                     // :55
    bool last_high = false;
    int last_r = 0;
    for (int q = 0; q < count; q++) {
      const unsigned long long key = keys[q];
      const int src = (int)(key & 0xffffffffu);
      const float r = __uint_as_float(~(unsigned)(key >> 32));
      const float x = xyz[3 * src], y = xyz[3 * src + 1], z = xyz[3 * src + 2];
      const float dx = TDR_FSUB(x, lx), dy = TDR_FSUB(y, ly);
      const float dist = TDR_FSQRT(TDR_FADD(TDR_FADD(0.f, TDR_FMUL(dx, dx)), TDR_FMUL(dy, dy)));   // :59
      const float slope = TDR_FDIV(fabsf(TDR_FSUB(z, lz)), dist);                           // :60
      const int r_ind = f2i_x86(round_half_away(TDR_FDIV(r, res)));                         // :61
      if (slope > 1.f) {                                                                    // :63
        if (r_ind >= 0 && r_ind < n_r) row[n_r + r_ind] += 1.f;
        last_high = true;
      } else if ((double)slope < 0.3 && !last_high) {                                       // :68
        for (int i = last_r; i <= r_ind; i++) if (i >= 0 && i < n_r) row[i] += 1.f;         // :69-73
      } else last_high = false;
      lx = x; ly = y; lz = z; last_r = r_ind;
    }
  }
  __syncthreads();
  for (int q = tid; q < 2 * n_r; q += blockDim.x) {
    const int img = q / n_r, r = q - img * n_r;
    out[((size_t)img * n_r + r) * n_theta + bin] = row[q];                                  // img(theta, r) at r * n_theta + theta
  }
}

// Cartesian: one thread per vertical scan line (column), integer counters (exact, order-independent)
__global__ void k_geo_cart(const uint8_t* __restrict__ pts, int stride, int width, int height, float res, int rows, int cols,
                           int32_t* __restrict__ cnt) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width) return;
  const int sx = cols, sy = rows;
  float lx = 0.f, ly = 0.f, lz = 0.f;
  int last_x = sx / 2, last_y = sy / 2;                                                     // :19
  bool last_high = false;
  for (int idy = 0; idy < height; idy++) {
    const uint8_t* p = pts + ((size_t)idy * width + idx) * stride;
    const float x = *reinterpret_cast<const float*>(p), y = *reinterpret_cast<const float*>(p + 4), z = *reinterpret_cast<const float*>(p + 8);
    if (x == 0.f && y == 0.f) continue;                                                     // :26
    const int x_ind = f2i_x86(TDR_FADD(round_half_away(TDR_FDIV(x, res)), (float)(sx / 2)));    // :27
    const int y_ind = f2i_x86(TDR_FADD(round_half_away(TDR_FDIV(y, res)), (float)(sy / 2)));    // :28
    const float dx = TDR_FSUB(x, lx), dy = TDR_FSUB(y, ly);
    const float dist = TDR_FSQRT(TDR_FADD(TDR_FADD(0.f, TDR_FMUL(dx, dx)), TDR_FMUL(dy, dy)));     // :30
    const float slope = TDR_FDIV(fabsf(TDR_FSUB(z, lz)), dist);                             // :31
    if (slope > 1.f) {
      if (x_ind >= 0 && x_ind < sx && y_ind >= 0 && y_ind < sy) atomicAdd(cnt + (size_t)rows * cols + (size_t)x_ind * rows + y_ind, 1);
      last_high = true;
    } else if ((double)slope < 0.3 && !last_high) {
      const int ddx = x_ind - last_x, ddy = y_ind - last_y;                                 // :38
      const int nrm = (int)sqrt((double)((long long)ddx * ddx + (long long)ddy * ddy));     // Vector2i::norm(): integer
      const double step = 1.0 / (double)nrm;                                                // inf when the cell did not change
      for (float i = 0.f; i < 1.f; i = (float)((double)i + step)) {                         // :39
        const int ix = f2i_x86(round_half_away(TDR_FADD((float)last_x, TDR_FMUL(i, (float)ddx))));   // :40
        const int iy = f2i_x86(round_half_away(TDR_FADD((float)last_y, TDR_FMUL(i, (float)ddy))));
        if (ix >= 0 && ix < sx && iy >= 0 && iy < sy) atomicAdd(cnt + (size_t)ix * rows + iy, 1);
      }
    } else last_high = false;
    lx = x; ly = y; lz = z; last_x = x_ind; last_y = y_ind;
  }
}

// polar: d0 = n_theta, d1 = n_r; cart: d0 = rows, d1 = cols.  dev_out: 2 images, column-major.
int scan_render_geometric(tdr_ctx* ctx, bool polar, float res, float ang_res, int d0, int d1, int width, int height, float* dev_out) {
  TDR_REQUIRE(ctx->n_pts > 0 && (long long)width * height == ctx->n_pts, TDR_EINVAL,
              "organised cloud of %d x %d points expected, %lld resident", width, height, (long long)ctx->n_pts);
  const uint8_t* pts = ctx->pts.as<uint8_t>();
  if (polar) {
    const size_t smem = (size_t)GEO_CAP * (8 + 12) + (size_t)2 * d1 * 4;
    TDR_SMEM_OPTIN(ctx, OPTIN_GEO_POLAR, k_geo_polar, smem);
    int* d_over = reinterpret_cast<int*>(ctx->scal.as<float>() + SC_MMA_MAXCOUNT);
    TDR_CUDA(cudaMemsetAsync(d_over, 0, 4, ctx->stream));
    k_geo_polar<<<d0, 1024, smem, ctx->stream>>>(pts, ctx->pts_stride, width, height, res, ang_res, d0, d1, dev_out, d_over);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    int over = 0;
    TDR_CUDA(cudaMemcpyAsync(&over, d_over, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
    TDR_REQUIRE(!over, TDR_EUNSUPPORTED, "an angular bin holds more than %d points", GEO_CAP);
    return TDR_OK;
  }
  const size_t cells = (size_t)2 * d0 * d1;
  if (int e = ctx->hist.reserve(cells * 4)) return e;
  TDR_CUDA(cudaMemsetAsync(ctx->hist.p, 0, cells * 4, ctx->stream));
  k_geo_cart<<<(width + 127) / 128, 128, 0, ctx->stream>>>(pts, ctx->pts_stride, width, height, res, d0, d1, ctx->hist.as<int32_t>());
  k_hist_to_float<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(ctx->hist.as<int32_t>(), (int)cells, dev_out);
  count_launch(ctx, 2);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
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
