// map_build.cu — a3/a4/a5: class image -> seeds -> exact truncated EDT -> device map pixels.
//
// Reference: TopDownMap::loadCompressedRasterMap (src/top_down_map.cpp:116-144),
// TopDownMap::computeDists (:289-326), getGeoRasterMap (:410-427).
//
// The distance fields are min(sqrtf(d2) * resolution, 50) with d2 the exact integer squared
// distance to the nearest class pixel (what cv::distanceTransform(DIST_L2, DIST_MASK_PRECISE)
// returns).  Because of the truncation at 50 only seeds within rcap = ceil(50/resolution) px
// can matter, so the separable search is windowed and trivially parallel:
//   pass V: g_c(y, x)  = vertical distance to the nearest class-c seed in column x (<= rcap, else 255)
//   pass H: d2_c(y, x) = min_{|dx| <= rcap} dx^2 + g_c(y, x+dx)^2        (early exit once dx^2 >= best)
// Everything is integer until the final sqrtf/mul, hence bit-exact.
#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

// seed byte: bit c (c < 7) = class c present at the pixel (binary layer == 0); bit 7 = unknown (class_mask_)
__global__ void k_class_image_to_seeds(const uint8_t* __restrict__ img, int h_img, int w_img, int stride,
                                       const int32_t* __restrict__ lut, int n_lut, int C, float res, int rows,
                                       int cols, uint8_t* __restrict__ seed) {
  int xi = blockIdx.x * blockDim.x + threadIdx.x;
  int yi = blockIdx.y * blockDim.y + threadIdx.y;
  if (xi >= cols || yi >= rows) return;
  // top_down_map.cpp:137-138   max<int>(height - yi*res - 1, 0), min<int>(xi*res, width-1)
  int src_row = f2i_x86(TDR_FSUB(TDR_FSUB((float)h_img, TDR_FMUL((float)yi, res)), 1.0f));
  src_row = src_row > 0 ? src_row : 0;
  int src_col = f2i_x86(TDR_FMUL((float)xi, res));
  src_col = src_col < w_img - 1 ? src_col : w_img - 1;
  int v = img[(size_t)src_row * stride + src_col];
  int cls = (v < n_lut) ? lut[v] : -1;
  uint8_t s = (cls >= 0 && cls < C) ? (uint8_t)(1u << cls) : (uint8_t)0x80;  // no class -> every layer stays 1 -> unknown
  seed[(size_t)yi * cols + xi] = s;
}

// binary float layers (col-major) -> seed bytes.  EDT seed: convertTo(CV_8U) == 0 i.e. cvRound(v) <= 0
// (top_down_map.cpp:306); mask: sum_c (uint8)trunc(v_c) > C-1 (:294-299).
__global__ void k_layers_to_seeds(const float* __restrict__ layers, int rows, int cols, int C,
                                  uint8_t* __restrict__ seed) {
  int yi = blockIdx.x * blockDim.x + threadIdx.x;   // y fastest in the col-major input
  int xi = blockIdx.y * blockDim.y + threadIdx.y;
  if (xi >= cols || yi >= rows) return;
  size_t L = (size_t)rows * cols;
  uint32_t bits = 0, msum = 0;
  for (int c = 0; c < C; c++) {
    float v = layers[(size_t)c * L + (size_t)xi * rows + yi];
    float r = rintf(v);                       // cvRound: nearest-even
    if (!(r >= 1.0f)) bits |= (1u << c);      // saturate_cast<uchar> of <= 0 (or NaN -> 0) is 0 -> seed
    int t = (int)v;                           // Eigen cast<uint8_t>: truncation
    msum = (msum + (uint32_t)(t & 0xff)) & 0xffu;
  }
  if ((int)msum > C - 1) bits |= 0x80u;
  seed[(size_t)yi * cols + xi] = (uint8_t)bits;
}

// geo seeds from class seeds: geo[1] is 0 where any class >= 3 is present, geo[0] = 1 - geo[1]
// (top_down_map.cpp:417-426).  bit0 = geo0 seed, bit1 = geo1 seed; mask never set (geo0 + geo1 == 1).
__global__ void k_geo_seeds(const uint8_t* __restrict__ seed, size_t n, int C, uint8_t* __restrict__ gseed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t hi = (C > 3) ? (((1u << C) - 1u) & ~7u) : 0u;
  bool obstacle = (seed[i] & hi) != 0;
  gseed[i] = obstacle ? 2 : 1;
}

// pass V as two column sweeps per band of `band` rows (thread = column x band): a counter per class holds the
// distance to the last seed seen, so the cost does not depend on how rare a class is (a window scan per pixel only
// stops when EVERY class has been found).  Down sweep writes the partial result, up sweep takes the minimum.
// The band starts rcap rows early so that its counters are exact where they are used.  segmin[c][y][x/32] = min of
// the final g over a 32-pixel segment lets pass H skip classes that have no seed anywhere near.
// (segmin == nullptr: not wanted — the packed row pass does not use it.)
__global__ void __launch_bounds__(128) k_edt_vsweep(const uint8_t* __restrict__ seed, int rows, int cols, int C, int rcap,
                                                    uint8_t* __restrict__ g, uint8_t* __restrict__ segmin, int segs, int band) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y_lo = blockIdx.y * band, y_hi = min(rows, y_lo + band) - 1;
  const bool live = x < cols;
  const int xc = live ? x : cols - 1;
  const size_t L = (size_t)rows * cols;
  uint32_t cnt[7];
#pragma unroll
  for (int c = 0; c < 7; c++) cnt[c] = 255;
  for (int y = max(0, y_lo - rcap); y <= y_hi; y++) {
    const uint32_t bits = seed[(size_t)y * cols + xc];
#pragma unroll
    for (int c = 0; c < 7; c++) {
      cnt[c] = ((bits >> c) & 1u) ? 0u : (cnt[c] >= 254u ? 255u : cnt[c] + 1u);
      if (c < C && y >= y_lo && live) g[(size_t)c * L + (size_t)y * cols + x] = (uint8_t)cnt[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 7; c++) cnt[c] = 255;
  for (int y = min(rows - 1, y_hi + rcap); y >= y_lo; y--) {
    const uint32_t bits = seed[(size_t)y * cols + xc];
#pragma unroll
    for (int c = 0; c < 7; c++) {
      cnt[c] = ((bits >> c) & 1u) ? 0u : (cnt[c] >= 254u ? 255u : cnt[c] + 1u);
      if (c < C && y <= y_hi) {                          // warp-uniform
        uint32_t v = 255u;
        if (live) {
          const size_t at = (size_t)c * L + (size_t)y * cols + x;
          v = min((uint32_t)g[at], cnt[c]);
          if (v > (uint32_t)rcap) v = 255u;              // beyond the truncation window: "no seed"
          g[at] = (uint8_t)v;
        }
        if (segmin) {                                    // kernel-uniform
          const uint32_t m = __reduce_min_sync(0xffffffffu, v);
          if ((threadIdx.x & 31) == 0 && x < cols) segmin[((size_t)c * rows + y) * segs + (x >> 5)] = (uint8_t)m;
        }
      }
    }
  }
}

// TO_MAP: write MapPixel records (class layers + known).  else: planar col-major float layers.
template <bool TO_MAP>
__global__ void k_edt_horizontal(const uint8_t* __restrict__ seed, const uint8_t* __restrict__ g,
                                 const uint8_t* __restrict__ segmin, int segs, int rows, int cols,
                                 int C, int rcap, uint32_t capcode, float resolution, MapPixel* __restrict__ map_px,
                                 float* __restrict__ planar) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  const size_t L = (size_t)rows * cols;
  const bool unknown = (seed[(size_t)y * cols + x] & 0x80u) != 0;
  const int s_lo = max(0, x - rcap) >> 5, s_hi = min(cols - 1, x + rcap) >> 5;
  float out[8];
#pragma unroll
  for (int c = 0; c < 8; c++) out[c] = 0.f;
#pragma unroll
  for (int c = 0; c < 7; c++) {
    if (c < C) {
      uint32_t best = 0x3fffffffu;
      if (!unknown) {                                     // unknown pixels are zeroed anyway (top_down_map.cpp:317)
        const uint8_t* sm = segmin + ((size_t)c * rows + y) * segs;
        uint32_t near = 255u;
        for (int sgi = s_lo; sgi <= s_hi; sgi++) near = min(near, (uint32_t)sm[sgi]);
        if (near != 255u) {                               // some seed of this class within reach of the window
          const uint8_t* gr = g + (size_t)c * L + (size_t)y * cols;
          uint32_t g0 = gr[x];
          best = (g0 == 255u) ? 0x3fffffffu : g0 * g0;
          for (int dx = 1; dx <= rcap && (uint32_t)(dx * dx) < best; dx++) {
            uint32_t dx2 = (uint32_t)(dx * dx);
            if (x - dx >= 0) { uint32_t gv = gr[x - dx]; if (gv != 255u) { uint32_t cand = dx2 + gv * gv; best = cand < best ? cand : best; } }
            if (x + dx < cols) { uint32_t gv = gr[x + dx]; if (gv != 255u) { uint32_t cand = dx2 + gv * gv; best = cand < best ? cand : best; } }
          }
        }
      }
      uint32_t d2 = best < capcode ? best : capcode;       // every d2 >= capcode is worth exactly 50
      out[c] = unknown ? 0.f : dist_value(d2, resolution);  // top_down_map.cpp:312-317
    }
  }
  if (TO_MAP) {
    out[7] = unknown ? 0.f : 1.f;
    float4* dst = reinterpret_cast<float4*>(map_px + (size_t)y * cols + x);
    dst[0] = make_float4(out[0], out[1], out[2], out[3]);
    dst[1] = make_float4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int c = 0; c < 7; c++)
      if (c < C) planar[(size_t)c * L + (size_t)x * rows + y] = out[c];
  }
}

// ------------------------------------------------------------------------------------------------
// Pass H, packed: the same minimum  d2_c(y, x) = min_dx dx^2 + g_c(y, x + dx)^2  as k_edt_horizontal, evaluated with the
// DPX instruction VIADDMNMX.U16x2 (__viaddmin_u16x2: min(a + b, c) on two 16-bit lanes): ONE instruction covers two
// taps, against ~6 for the scalar tap (load, test, multiply-add, compare, select) — that kernel was instruction-bound at
// 3.6 % of its HBM roofline.  A CTA owns EDT_XT pixels of one row: for every class the squared column distances g^2 of
// the tile and its +-R halo go to shared memory as 16-bit words (0x7fff = no seed: every sum with it stays above any
// capped distance and below 2^16); a thread owns 8 adjacent pixels and walks the 2R + 8 taps they share, eight taps
// per 128-bit shared load.  The squared offsets come from one table E[j] = ((j - R - 7)^2, (j - R - 6)^2), read with
// warp-uniform addresses (broadcast).  No early exit, no branches: every pixel costs the same ~65 instructions per class.
// Taps beyond the truncation window only add candidates >= capcode, which the final min(d2, capcode) removes, so the
// result equals the windowed scan bit for bit (tests/test_gpu_parity.py EDT cases, the cv2 fixture, the reference build).
// Needs (R + 7)^2 <= 0x7fff, i.e. windows up to 174 px (resolution >= 0.29 m/px); finer maps keep k_edt_horizontal.
static const int EDT_XT_THREADS = 128, EDT_XT = EDT_XT_THREADS * 8;
template <bool TO_MAP>
__global__ void __launch_bounds__(EDT_XT_THREADS) k_edt_rows_dpx(const uint8_t* __restrict__ seed, const uint8_t* __restrict__ g,
                                                                 int rows, int cols, int C, int R, uint32_t capcode,
                                                                 float resolution, MapPixel* __restrict__ map_px,
                                                                 float* __restrict__ planar) {
  extern __shared__ __align__(16) unsigned char edt_smem[];
  const int span = EDT_XT + 2 * R + 8;                          // 16-bit words per class (multiple of 8)
  const int n_e = 2 * R + 24;                                   // table entries (multiple of 8)
  uint32_t* s_e = reinterpret_cast<uint32_t*>(edt_smem);
  uint16_t* s_g2 = reinterpret_cast<uint16_t*>(edt_smem + (size_t)n_e * 4);
  const int y = blockIdx.y, x0 = blockIdx.x * EDT_XT, tid = threadIdx.x;
  const size_t L = (size_t)rows * cols;
  for (int j = tid; j < n_e; j += EDT_XT_THREADS) {
    const int d = j - R - 7;
    s_e[j] = (uint32_t)(d * d) | ((uint32_t)((d + 1) * (d + 1)) << 16);
  }
  for (int c = 0; c < C; c++) {
    const uint8_t* gr = g + (size_t)c * L + (size_t)y * cols;
    uint16_t* dst = s_g2 + (size_t)c * span;
    for (int j = tid; j < span; j += EDT_XT_THREADS) {
      const int x = x0 - R + j;
      uint32_t v = 255u;
      if (x >= 0 && x < cols) v = gr[x];
      dst[j] = v == 255u ? (uint16_t)0x7fff : (uint16_t)(v * v);
    }
  }
  __syncthreads();
  const int xb = x0 + 8 * tid;                                   // this thread's pixels xb .. xb + 7
  if (xb >= cols) return;
  uint32_t unk = 0;
#pragma unroll
  for (int p = 0; p < 8; p++) if (xb + p < cols && (seed[(size_t)y * cols + xb + p] & 0x80u)) unk |= 1u << p;
  uint32_t d2p[8][4];                                            // [pixel][class pair] two capped d2 per word
#pragma unroll
  for (int p = 0; p < 8; p++) { d2p[p][0] = d2p[p][1] = d2p[p][2] = d2p[p][3] = 0; }
  const int n_blocks = (2 * R + 8) / 8;
#pragma unroll
  for (int c = 0; c < 7; c++) {
    if (c < C) {
      uint32_t best[8];
#pragma unroll
      for (int p = 0; p < 8; p++) best[p] = 0x7fff7fffu;
      const uint4* src = reinterpret_cast<const uint4*>(s_g2 + (size_t)c * span + 8 * tid);
      const uint4* et = reinterpret_cast<const uint4*>(s_e);
#pragma unroll 1
      for (int b = 0; b < n_blocks; b++) {
        const uint4 gv = src[b];                                 // taps 8b .. 8b + 7 of this thread's window
        const uint4 e0 = et[2 * b], e1 = et[2 * b + 1], e2 = et[2 * b + 2], e3 = et[2 * b + 3];
        const uint32_t e[16] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
        const uint32_t gq[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int p = 0; p < 8; p++) best[p] = __viaddmin_u16x2(gq[q], e[2 * q + 7 - p], best[p]);   // taps 8b+2q, +1 against pixel p
      }
      // (visiting the blocks from the centre outwards and stopping once no unvisited column can improve any of the eight
      // minima was measured SLOWER, 0.71 vs 0.51 ms at 4000^2 x 6: the test costs more than the taps it saves)
#pragma unroll
      for (int p = 0; p < 8; p++) {
        uint32_t d2 = min(best[p] & 0xffffu, best[p] >> 16);
        d2 = d2 < capcode ? d2 : capcode;
        d2p[p][c >> 1] |= d2 << (16 * (c & 1));
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int x = xb + p;
    if (x >= cols) break;
    const bool unknown = (unk >> p) & 1u;
    float out[8];
#pragma unroll
    for (int c = 0; c < 8; c++) out[c] = 0.f;
#pragma unroll
    for (int c = 0; c < 7; c++)
      if (c < C) out[c] = unknown ? 0.f : dist_value((d2p[p][c >> 1] >> (16 * (c & 1))) & 0xffffu, resolution);   // top_down_map.cpp:312-317
    if (TO_MAP) {
      out[7] = unknown ? 0.f : 1.f;
      float4* dst = reinterpret_cast<float4*>(map_px + (size_t)y * cols + x);
      dst[0] = make_float4(out[0], out[1], out[2], out[3]);
      dst[1] = make_float4(out[4], out[5], out[6], out[7]);
    } else {
#pragma unroll
      for (int c = 0; c < 7; c++)
        if (c < C) planar[(size_t)c * L + (size_t)x * rows + y] = out[c];
    }
  }
}

// cached distance layers (col-major) + mask -> MapPixel
__global__ void k_dist_layers_to_map(const float* __restrict__ layers, const uint8_t* __restrict__ mask, int rows,
                                     int cols, int C, MapPixel* __restrict__ map_px) {
  int y = blockIdx.x * blockDim.x + threadIdx.x;
  int x = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  size_t L = (size_t)rows * cols;
  float out[8];
#pragma unroll
  for (int c = 0; c < 8; c++) out[c] = 0.f;
  for (int c = 0; c < C; c++) out[c] = layers[(size_t)c * L + (size_t)x * rows + y];
  out[7] = mask[(size_t)x * rows + y] ? 0.f : 1.f;
  float4* dst = reinterpret_cast<float4*>(map_px + (size_t)y * cols + x);
  dst[0] = make_float4(out[0], out[1], out[2], out[3]);
  dst[1] = make_float4(out[4], out[5], out[6], out[7]);
}

// MapPixel -> the reference's col-major layers + mask
__global__ void k_map_to_layers(const MapPixel* __restrict__ map_px, int rows, int cols, int C,
                                float* __restrict__ layers, uint8_t* __restrict__ mask) {
  int y = blockIdx.x * blockDim.x + threadIdx.x;
  int x = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= cols || y >= rows) return;
  size_t L = (size_t)rows * cols;
  const float4* src = reinterpret_cast<const float4*>(map_px + (size_t)y * cols + x);
  float4 a = src[0], b = src[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  for (int c = 0; c < C; c++) layers[(size_t)c * L + (size_t)x * rows + y] = v[c];
  if (mask) mask[(size_t)x * rows + y] = v[7] != 0.f ? 0 : 1;
}

// smallest d2 whose value is already 50 (host, same IEEE ops as dist_value)
static uint32_t compute_capcode(float resolution) {
  for (uint32_t d2 = 0; d2 < (1u << 24); d2++) {
    float d = sqrtf((float)d2) * resolution;
    if (d >= 50.0f) return d2;
  }
  return 1u << 24;
}

static int run_edt(tdr_ctx* ctx, const uint8_t* seed, int rows, int cols, int C, float resolution, bool to_map,
                   float* planar_out) {
  uint32_t capcode = compute_capcode(resolution);
  int rcap = (int)ceil(sqrt((double)capcode)) + 1;
  TDR_REQUIRE(rcap <= 254, TDR_EUNSUPPORTED, "resolution %g needs an EDT window of %d px (> 254)", resolution, rcap);
  size_t L = (size_t)rows * cols;
  const int segs = (cols + 31) / 32;
  const size_t seg_bytes = (size_t)C * rows * segs;
  if (int e = ctx->edt_g.reserve(L * (size_t)C + seg_bytes + 256)) return e;
  uint8_t* d_g = ctx->edt_g.as<uint8_t>();
  uint8_t* d_seg = d_g + ((L * (size_t)C + 255) / 256) * 256;
  // the packed row pass (16-bit lanes) wherever its window fits; TDR_EDT_IMPL=1 forces the scalar tap scan
  const int R = (rcap + 3) / 4 * 4;
  const bool packed = (R + 7) * (R + 7) <= 0x7fff && capcode <= 0x7fffu && ctx->edt_impl != 1;
  const int band = ctx->edt_band > 0 ? ctx->edt_band : 64;      // 128: 0.39 ms, 64: 0.36, 32: 0.35, 256: 0.67 at 4000^2 x 6
  dim3 vgrd((cols + 127) / 128, (rows + band - 1) / band);
  k_edt_vsweep<<<vgrd, 128, 0, ctx->stream>>>(seed, rows, cols, C, rcap, d_g, packed ? nullptr : d_seg, segs, band);
  if (packed) {
    const size_t smem = (size_t)(2 * R + 24) * 4 + (size_t)C * (EDT_XT + 2 * R + 8) * 2;      // <= 21 KB (C <= 7, R <= 176)
    dim3 rgrd((cols + EDT_XT - 1) / EDT_XT, rows);
    if (to_map) {
      k_edt_rows_dpx<true><<<rgrd, EDT_XT_THREADS, smem, ctx->stream>>>(seed, d_g, rows, cols, C, R, capcode, resolution,
                                                                       ctx->map_px.as<MapPixel>(), nullptr);
    } else {
      k_edt_rows_dpx<false><<<rgrd, EDT_XT_THREADS, smem, ctx->stream>>>(seed, d_g, rows, cols, C, R, capcode, resolution, nullptr,
                                                                              planar_out);
    }
    count_launch(ctx, 2);
    TDR_CUDA(cudaGetLastError());
    return TDR_OK;
  }
  dim3 blk(32, 8), grd((cols + 31) / 32, (rows + 7) / 8);
  if (to_map)
    k_edt_horizontal<true><<<grd, blk, 0, ctx->stream>>>(seed, d_g, d_seg, segs, rows, cols, C, rcap, capcode,
                                                         resolution, ctx->map_px.as<MapPixel>(), nullptr);
  else
    k_edt_horizontal<false><<<grd, blk, 0, ctx->stream>>>(seed, d_g, d_seg, segs, rows, cols, C, rcap, capcode,
                                                          resolution, nullptr, planar_out);
  count_launch(ctx, 2);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

static int map_alloc(tdr_ctx* ctx, int rows, int cols, int C, float resolution) {
  TDR_REQUIRE(rows > 0 && cols > 0, TDR_EINVAL, "empty map %d x %d", rows, cols);
  TDR_REQUIRE(C >= 1 && C <= TDR_MAX_CLASSES, TDR_EUNSUPPORTED, "num_classes %d not in [1, %d]", C, TDR_MAX_CLASSES);
  TDR_REQUIRE(resolution > 0.f, TDR_EINVAL, "resolution must be positive");
  size_t L = (size_t)rows * cols;
  if (int e = ctx->map_px.reserve(L * sizeof(MapPixel))) return e;
  if (int e = ctx->seedbits.reserve(L)) return e;
  ctx->rows = rows; ctx->cols = cols; ctx->C = C; ctx->resolution = resolution;
  ctx->map16_valid = false; ctx->map8_valid = false; ctx->map16g_log2 = -1; ctx->perm_grid_n = -1; ctx->geo_valid = false;
  return TDR_OK;
}

int map_set_class_image(tdr_ctx* ctx, const uint8_t* img, int h_img, int w_img, int stride, const int32_t* lut,
                        int n_lut, int C, float resolution) {
  TDR_REQUIRE(img && lut && h_img > 0 && w_img > 0 && stride >= w_img && n_lut > 0, TDR_EINVAL, "bad class image arguments");
  TDR_REQUIRE(resolution > 0.f, TDR_EINVAL, "resolution must be positive");
  int rows = (int)(h_img / resolution), cols = (int)(w_img / resolution);   // top_down_map.cpp:121-122
  if (int e = map_alloc(ctx, rows, cols, C, resolution)) return e;
  size_t img_bytes = (size_t)h_img * stride;
  if (int e = ctx->scratch.reserve(img_bytes)) return e;
  if (int e = ctx->scratch2.reserve((size_t)n_lut * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch.p, img, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch2.p, lut, (size_t)n_lut * 4, cudaMemcpyHostToDevice, ctx->stream));
  dim3 blk(32, 8), grd((cols + 31) / 32, (rows + 7) / 8);
  k_class_image_to_seeds<<<grd, blk, 0, ctx->stream>>>(ctx->scratch.as<uint8_t>(), h_img, w_img, stride,
                                                       ctx->scratch2.as<int32_t>(), n_lut, C, resolution, rows, cols,
                                                       ctx->seedbits.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (int e = run_edt(ctx, ctx->seedbits.as<uint8_t>(), rows, cols, C, resolution, true, nullptr)) return e;
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));  // img / lut are caller-owned pageable buffers
  ctx->have_map = true; ctx->have_seeds = true;
  return TDR_OK;
}

int map_set_binary_layers(tdr_ctx* ctx, const float* layers, int rows, int cols, int C, float resolution) {
  TDR_REQUIRE(layers, TDR_EINVAL, "null layers");
  if (int e = map_alloc(ctx, rows, cols, C, resolution)) return e;
  size_t L = (size_t)rows * cols;
  if (int e = ctx->scratch.reserve(L * C * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch.p, layers, L * C * 4, cudaMemcpyHostToDevice, ctx->stream));
  dim3 blk(32, 8), grd((rows + 31) / 32, (cols + 7) / 8);
  k_layers_to_seeds<<<grd, blk, 0, ctx->stream>>>(ctx->scratch.as<float>(), rows, cols, C, ctx->seedbits.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (int e = run_edt(ctx, ctx->seedbits.as<uint8_t>(), rows, cols, C, resolution, true, nullptr)) return e;
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_map = true; ctx->have_seeds = true;
  return TDR_OK;
}

// ---- vector map -> binary class layers (getRasterMap :391-408, samplePts :367-389, getClasses :328-365)
// Eigen LinSpaced(size, -res*(size-1)/2., res*(size-1)/2.)[i] (same as k_local_cart's)
__device__ __forceinline__ float raster_linspaced(int size, float sres, int i) {
  float low = (float)((double)TDR_FMUL(-sres, (float)(size - 1)) / 2.);
  float high = (float)((double)TDR_FMUL(sres, (float)(size - 1)) / 2.);
  if (size == 1) return low;
  float step = TDR_FDIV(TDR_FSUB(high, low), (float)(size - 1));
  bool flip = fabsf(high) < fabsf(low);
  int size1 = size - 1;
  if (flip) return (i == 0) ? low : TDR_FSUB(high, TDR_FMUL((float)(size1 - i), step));
  return (i == size1) ? high : TDR_FADD(low, TDR_FMUL((float)i, step));
}
struct PolyParams {
  const float2* verts; const int32_t* start; const int32_t* cls; const float4* bbox; int n_poly;   // bbox: xmin, xmax, ymin, ymax
  int rows, cols, C; float resolution, cr, sr, c0, c1;
  int excl[2 * TDR_MAX_CLASSES]; int n_excl;
};
// One thread per sample point (= map pixel).  The reference runs the even-odd rule edge by edge over ALL points
// (O(pixels x edges), :339-349); a polygon can only flip points whose y lies inside its y-range and whose x is not
// clearly to its right, and left of it the flips of a closed polygon cancel — those polygons are skipped (exactly),
// the edges of the rest are evaluated with the reference's fp32 expression.  A block (256 consecutive points of a map
// column) first collects, cooperatively, the polygons whose padded bounding box meets the block's extent.
static const int RASTER_THREADS = 256, RASTER_CHUNK = 2048;
__global__ void __launch_bounds__(RASTER_THREADS) k_raster_polygons(PolyParams q, float* __restrict__ layers) {
  __shared__ int s_list[RASTER_CHUNK];
  __shared__ int s_n;
  __shared__ float s_ext[4][RASTER_THREADS / 32];                            // per-warp min/max of py, px
  const size_t L = (size_t)q.rows * q.cols;
  const size_t p_raw = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p_raw < L;
  const size_t p = live ? p_raw : L - 1;
  const float a = raster_linspaced(q.rows, q.resolution, (int)(p % q.rows));   // pts(0, p)
  const float b = raster_linspaced(q.cols, q.resolution, (int)(p / q.rows));   // pts(1, p)
  const float py = TDR_FADD(TDR_FADD(TDR_FMUL(q.cr, a), TDR_FMUL(-q.sr, b)), q.c1);   // rotm * pts, += center[1]
  const float px = TDR_FADD(TDR_FADD(TDR_FMUL(q.sr, a), TDR_FMUL(q.cr, b)), q.c0);    // += center[0]
  // extent of this block's sample points: only polygons whose (padded) bounding box meets it can flip any of them
  float ylo = py, yhi = py, xlo = px, xhi = px;
  for (int o = 16; o > 0; o >>= 1) {
    ylo = fminf(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = fmaxf(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
    xlo = fminf(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = fmaxf(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    s_ext[0][w] = ylo; s_ext[1][w] = yhi; s_ext[2][w] = xlo; s_ext[3][w] = xhi;
  }
  __syncthreads();
  for (int w = 0; w < RASTER_THREADS / 32; w++) {
    ylo = fminf(ylo, s_ext[0][w]); yhi = fmaxf(yhi, s_ext[1][w]); xlo = fminf(xlo, s_ext[2][w]); xhi = fmaxf(xhi, s_ext[3][w]);
  }
  float fills[TDR_MAX_CLASSES];
#pragma unroll
  for (int c = 0; c < TDR_MAX_CLASSES; c++) fills[c] = -1.f;                 // class_fills = -1
  for (int k0 = 0; k0 < q.n_poly; k0 += RASTER_CHUNK) {
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int k = k0 + threadIdx.x; k < q.n_poly && k < k0 + RASTER_CHUNK; k += RASTER_THREADS) {
      const float4 bb = __ldg(q.bbox + k);
      if (yhi >= bb.z && ylo < bb.w && xhi >= bb.x - 1.0f && xlo < bb.y + 1.0f) s_list[atomicAdd(&s_n, 1)] = k;
    }
    __syncthreads();
    const int n_list = s_n;
    for (int e = 0; e < n_list; e++) {
      const int k = s_list[e];
      const float4 bb = __ldg(q.bbox + k);
      if (py < bb.z || !(py < bb.w)) continue;                               // no edge straddles this y
      if (!(px < bb.y + 1.0f) || px < bb.x - 1.0f) continue;                 // right of it: no flip; left of it: an even number
      const int s0 = q.start[k], n = q.start[k + 1] - s0;
      float buf = -1.f;                                                      // class_fills_buf = -1
      float2 vj = __ldg(q.verts + s0 + n - 1);
      for (int i = 0; i < n; i++) {
        const float2 vi = __ldg(q.verts + s0 + i);
        const bool ca = (py < vi.y) != (py < vj.y);
        const bool cb = px < TDR_FADD(vi.x, TDR_FDIV(TDR_FMUL(TDR_FSUB(vj.x, vi.x), TDR_FSUB(py, vi.y)), TDR_FSUB(vj.y, vi.y)));
        if (ca && cb) buf = -buf;                                            // *= -2 * cond + 1
        vj = vi;
      }
      const int c = q.cls[k];
#pragma unroll
      for (int cc = 0; cc < TDR_MAX_CLASSES; cc++) if (cc == c) fills[cc] = fmaxf(fills[cc], buf);
    }
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < TDR_MAX_CLASSES; c++) fills[c] = TDR_FDIV(TDR_FADD(TDR_FMUL(fills[c], -1.f), 1.f), 2.f);   // :351-353
  for (int ia = 0; ia < q.n_excl; ia++) {                                   // "only one ground type per cell" :357-364
    const int under = q.excl[ia];
    float u = 0.f;
#pragma unroll
    for (int cc = 0; cc < TDR_MAX_CLASSES; cc++) if (cc == under) u = fills[cc];
    for (int ib = 0; ib < q.n_excl; ib++) {
      const int cls = q.excl[ib];
      if (under < cls) {
        float v = 0.f;
#pragma unroll
        for (int cc = 0; cc < TDR_MAX_CLASSES; cc++) if (cc == cls) v = fills[cc];
        u = TDR_FADD(u, TDR_FSUB(1.f, v));
      }
    }
    u = fminf(u, 1.f);
#pragma unroll
    for (int cc = 0; cc < TDR_MAX_CLASSES; cc++) if (cc == under) fills[cc] = u;
  }
#pragma unroll
  for (int c = 0; c < TDR_MAX_CLASSES; c++) if (c < q.C) layers[(size_t)c * L + p] = fills[c];
}

int map_set_polygons(tdr_ctx* ctx, const float* verts, const int32_t* poly_start, const int32_t* poly_class, int n_poly, int map_w,
                     int map_h, float rot, int C, float resolution, const int32_t* exclusive, int n_excl, float* layers_out) {
  TDR_REQUIRE(n_poly >= 0 && map_w > 0 && map_h > 0 && resolution > 0.f, TDR_EINVAL, "bad polygon map arguments");
  TDR_REQUIRE(n_poly == 0 || (verts && poly_start && poly_class), TDR_EINVAL, "null polygon arrays");
  TDR_REQUIRE(n_excl >= 0 && n_excl <= 2 * TDR_MAX_CLASSES && (n_excl == 0 || exclusive), TDR_EINVAL, "bad exclusive class list");
  const int rows = (int)(map_h / resolution), cols = (int)(map_w / resolution);   // :399-400
  if (int e = map_alloc(ctx, rows, cols, C, resolution)) return e;
  const size_t L = (size_t)rows * cols;
  const int n_verts = n_poly ? poly_start[n_poly] : 0;
  std::vector<float4> bbox((size_t)(n_poly > 0 ? n_poly : 1));
  for (int k = 0; k < n_poly; k++) {
    TDR_REQUIRE(poly_class[k] >= 0 && poly_class[k] < C && poly_start[k + 1] > poly_start[k], TDR_EINVAL, "bad polygon %d", k);
    float4 bb = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
    for (int i = poly_start[k]; i < poly_start[k + 1]; i++) {
      bb.x = fminf(bb.x, verts[2 * i]); bb.y = fmaxf(bb.y, verts[2 * i]);
      bb.z = fminf(bb.z, verts[2 * i + 1]); bb.w = fmaxf(bb.w, verts[2 * i + 1]);
    }
    bbox[k] = bb;
  }
  for (int k = 0; k < n_excl; k++) TDR_REQUIRE(exclusive[k] >= 0 && exclusive[k] < C, TDR_EINVAL, "exclusive class %d out of range", exclusive[k]);
  // device staging: layers (C*L floats) in scratch; polygon arrays in scratch2
  const size_t off_start = (size_t)n_verts * 8, off_cls = off_start + (size_t)(n_poly + 1) * 4, off_bbox = ((off_cls + (size_t)n_poly * 4 + 15) / 16) * 16;
  if (int e = ctx->scratch.reserve(L * C * 4)) return e;
  if (int e = ctx->scratch2.reserve(off_bbox + (size_t)(n_poly + 1) * 16 + L)) return e;   // + L: map_get_layers' mask staging
  unsigned char* d2 = ctx->scratch2.as<unsigned char>();
  if (n_poly > 0) {
    TDR_CUDA(cudaMemcpyAsync(d2, verts, (size_t)n_verts * 8, cudaMemcpyHostToDevice, ctx->stream));
    TDR_CUDA(cudaMemcpyAsync(d2 + off_start, poly_start, (size_t)(n_poly + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    TDR_CUDA(cudaMemcpyAsync(d2 + off_cls, poly_class, (size_t)n_poly * 4, cudaMemcpyHostToDevice, ctx->stream));
    TDR_CUDA(cudaMemcpyAsync(d2 + off_bbox, bbox.data(), (size_t)n_poly * 16, cudaMemcpyHostToDevice, ctx->stream));
  }
  PolyParams q; memset(&q, 0, sizeof(q));
  q.verts = reinterpret_cast<const float2*>(d2); q.start = reinterpret_cast<const int32_t*>(d2 + off_start);
  q.cls = reinterpret_cast<const int32_t*>(d2 + off_cls); q.bbox = reinterpret_cast<const float4*>(d2 + off_bbox); q.n_poly = n_poly;
  q.rows = rows; q.cols = cols; q.C = C; q.resolution = resolution; q.cr = cosf(rot); q.sr = sinf(rot);
  q.c0 = (float)map_w / 2; q.c1 = (float)map_h / 2;
  q.n_excl = n_excl;
  for (int k = 0; k < n_excl; k++) q.excl[k] = exclusive[k];
  k_raster_polygons<<<(unsigned)((L + RASTER_THREADS - 1) / RASTER_THREADS), RASTER_THREADS, 0, ctx->stream>>>(q, ctx->scratch.as<float>());
  dim3 blk(32, 8), grd((rows + 31) / 32, (cols + 7) / 8);
  k_layers_to_seeds<<<grd, blk, 0, ctx->stream>>>(ctx->scratch.as<float>(), rows, cols, C, ctx->seedbits.as<uint8_t>());
  count_launch(ctx, 2);
  TDR_CUDA(cudaGetLastError());
  if (layers_out) TDR_CUDA(cudaMemcpyAsync(layers_out, ctx->scratch.p, L * C * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (int e = run_edt(ctx, ctx->seedbits.as<uint8_t>(), rows, cols, C, resolution, true, nullptr)) return e;
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));      // bbox (host vector) and the caller's arrays may go away
  ctx->have_map = true; ctx->have_seeds = true;
  return TDR_OK;
}

// seeds produced on the device (ctx->scratch2, rows*cols bytes, k_layers_to_seeds' bit layout) -> map + distance fields
int map_from_seeds(tdr_ctx* ctx, int rows, int cols, int C, float resolution) {
  if (int e = map_alloc(ctx, rows, cols, C, resolution)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->seedbits.p, ctx->scratch2.p, (size_t)rows * cols, cudaMemcpyDeviceToDevice, ctx->stream));
  if (int e = run_edt(ctx, ctx->seedbits.as<uint8_t>(), rows, cols, C, resolution, true, nullptr)) return e;
  ctx->have_map = true; ctx->have_seeds = true;
  return TDR_OK;
}

int map_set_dist_layers(tdr_ctx* ctx, const float* layers, const uint8_t* mask, int rows, int cols, int C,
                        float resolution) {
  TDR_REQUIRE(layers && mask, TDR_EINVAL, "null layers / mask");
  if (int e = map_alloc(ctx, rows, cols, C, resolution)) return e;
  size_t L = (size_t)rows * cols;
  if (int e = ctx->scratch.reserve(L * C * 4)) return e;
  if (int e = ctx->scratch2.reserve(L)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch.p, layers, L * C * 4, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch2.p, mask, L, cudaMemcpyHostToDevice, ctx->stream));
  dim3 blk(32, 8), grd((rows + 31) / 32, (cols + 7) / 8);
  k_dist_layers_to_map<<<grd, blk, 0, ctx->stream>>>(ctx->scratch.as<float>(), ctx->scratch2.as<uint8_t>(), rows, cols,
                                                     C, ctx->map_px.as<MapPixel>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_map = true; ctx->have_seeds = false;   // class presence is not recoverable from distances under the mask
  return TDR_OK;
}

int map_get_layers(tdr_ctx* ctx, float* layers, uint8_t* mask) {
  TDR_REQUIRE(ctx->have_map, TDR_ESTATE, "no map");
  TDR_REQUIRE(layers, TDR_EINVAL, "null layers");
  size_t L = (size_t)ctx->rows * ctx->cols;
  if (int e = ctx->scratch.reserve(L * ctx->C * 4)) return e;
  if (int e = ctx->scratch2.reserve(L)) return e;
  dim3 blk(32, 8), grd((ctx->rows + 31) / 32, (ctx->cols + 7) / 8);
  k_map_to_layers<<<grd, blk, 0, ctx->stream>>>(ctx->map_px.as<MapPixel>(), ctx->rows, ctx->cols, ctx->C,
                                                ctx->scratch.as<float>(), ctx->scratch2.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(layers, ctx->scratch.p, L * ctx->C * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (mask) TDR_CUDA(cudaMemcpyAsync(mask, ctx->scratch2.p, L, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

// the geo distance layers resident on the device (for getLocalGeoMap)
int map_geo_resident(tdr_ctx* ctx) {
  if (ctx->have_map && ctx->geo_valid) return TDR_OK;            // built earlier, or uploaded from the map cache
  TDR_REQUIRE(ctx->have_map && ctx->have_seeds, TDR_ESTATE, "geo layers need a map built from class seeds (or tdr_map_set_geo_dist_layers)");
  size_t L = (size_t)ctx->rows * ctx->cols;
  if (int e = ctx->scratch2.reserve(L)) return e;
  if (int e = ctx->geo_planar.reserve(L * 2 * 4)) return e;
  k_geo_seeds<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->seedbits.as<uint8_t>(), L, ctx->C, ctx->scratch2.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (int e = run_edt(ctx, ctx->scratch2.as<uint8_t>(), ctx->rows, ctx->cols, 2, ctx->resolution, false, ctx->geo_planar.as<float>())) return e;
  ctx->geo_valid = true;
  return TDR_OK;
}

int map_get_geo_layers(tdr_ctx* ctx, float* geo_layers) {
  TDR_REQUIRE(ctx->have_map && ctx->have_seeds, TDR_ESTATE, "geo layers need a map built from class seeds");
  TDR_REQUIRE(geo_layers, TDR_EINVAL, "null output");
  size_t L = (size_t)ctx->rows * ctx->cols;
  if (int e = ctx->scratch2.reserve(L)) return e;
  if (int e = ctx->scratch.reserve(L * 2 * 4)) return e;
  k_geo_seeds<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->seedbits.as<uint8_t>(), L, ctx->C,
                                                                     ctx->scratch2.as<uint8_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (int e = run_edt(ctx, ctx->scratch2.as<uint8_t>(), ctx->rows, ctx->cols, 2, ctx->resolution, false,
                      ctx->scratch.as<float>()))
    return e;
  TDR_CUDA(cudaMemcpyAsync(geo_layers, ctx->scratch.p, L * 2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

}  // namespace tdr
