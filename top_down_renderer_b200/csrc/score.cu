// score.cu — a7/a9/a10: polar local-map gather + circular shift-correlation + weight.
//
// Reference: TopDownMapPolar::getLocalMap (src/top_down_map_polar.cpp:21-53),
// StateParticle::getCostForRot (src/state_particle.cpp:112-155),
// StateParticle::computeWeight (:157-219).
//
// Data layout: the map is one 32-byte record per pixel (MapPixel: <=7 class distances + known),
// so ONE sector serves every class layer and the mask of a lattice point (the reference's planar
// col-major layers need C+1 sectors).  A lane pair fetches the two 16-byte halves of a record.
// The scan is pre-packed per lattice cell as 8 floats (class counts scaled by 0.01*w_c, and the
// class-summed count for the normalisation) and staged in shared memory once per CTA.
//
//   k_score_track : one warp per particle, one row shift (tracking; have_init == true)
//   k_score_search: one CTA per hypothesis centre, n_shifts row shifts evaluated against the SAME
//                   gathered local map held in registers (theta search; also the exhaustive grid)
#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

struct ScoreParams {
  const MapPixel* map; int rows, cols; float resolution;
  const float* tab; int n_theta, n_r, P;
  const float* scan_pack;
  float res;
  // particles
  const float *init_x, *init_y, *dx, *dy; float* theta; const float* scale; uint8_t* have_init;
  long long n;
  // gates / weight
  int force_on_map; float map_w, map_h; int scale_gate; double scale_lo, scale_hi; float regularization;
  // search list
  const float* thetas; const int32_t* shifts; int n_shifts;
  // outputs
  float* weights;
  // grid mode
  const float* centers; float grid_scale; float* costs;
  // launched BEHIND a tensor-core kernel: run only if that one bailed out (scan counts above 2048, decided on the device)
  const int* only_if_bailed;
};

__device__ __forceinline__ float4 ldg4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float weight_from_cost(float cost, float reg) {
  return (float)(1.0 / (double)TDR_FADD(cost, reg));   // state_particle.cpp:212
}

// centre + gates (state_particle.cpp:161-176).  returns false when the particle is gated (weight 0).
__device__ __forceinline__ bool particle_centre(const ScoreParams& sp, long long i, float* cx, float* cy, float* sc) {
  float s = sp.scale[i];
  float x = TDR_FADD(TDR_FMUL(sp.dx[i], s), sp.init_x[i]);
  float y = TDR_FADD(TDR_FMUL(sp.dy[i], s), sp.init_y[i]);
  *cx = x; *cy = y; *sc = s;
  if (sp.force_on_map && (x < 0.f || y < 0.f || x > sp.map_w || y > sp.map_h)) return false;
  if (sp.scale_gate && ((double)s < sp.scale_lo || (double)s > sp.scale_hi)) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// tracking: warp per particle, lane per lattice cell.
// shared: scan_pack (P*8 floats) | tab (P float2) | cell (P x ushort2: theta, n_theta*r).  Every record is ONE
// 256-bit load (one sector, one L1 wavefront: tools/gather_bench.cu) and four cells per lane are in flight.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void ldg256f(const void* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}

__global__ void __launch_bounds__(512, 2) k_score_track(ScoreParams sp) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (sp.only_if_bailed && *sp.only_if_bailed == 0) return;      // launched behind the tensor-core kernel, which did the work
  float* s_scan = reinterpret_cast<float*>(smem);
  float2* s_tab = reinterpret_cast<float2*>(s_scan + (size_t)sp.P * 8);
  ushort2* s_cell = reinterpret_cast<ushort2*>(s_tab + sp.P);
  for (int i = threadIdx.x; i < sp.P * 2; i += blockDim.x)       // float4 copies of the packed scan
    reinterpret_cast<float4*>(s_scan)[i] = ldg4(sp.scan_pack + (size_t)i * 4);
  for (int i = threadIdx.x; i < sp.P; i += blockDim.x) {
    s_tab[i] = reinterpret_cast<const float2*>(sp.tab)[i];
    const int r = i / sp.n_theta;
    s_cell[i] = make_ushort2((unsigned short)(i - r * sp.n_theta), (unsigned short)(r * sp.n_theta));
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const char* map_bytes = reinterpret_cast<const char*>(sp.map);
  const int n_theta = sp.n_theta;
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < sp.n; i += warps_total) {
    if (!sp.have_init[i]) continue;                       // searched by the theta-search kernels
    float cx, cy, sc;
    if (!particle_centre(sp, i, &cx, &cy, &sc)) { if (lane == 0) sp.weights[i] = 0.f; continue; }
    const float oy = TDR_FDIV(cy, sp.resolution), ox = TDR_FDIV(cx, sp.resolution);   // top_down_map_polar.cpp:29-30
    const int shift = rot_to_shift(sp.theta[i], n_theta);
    float acc_c = 0.f, acc_n = 0.f, acc_k = 0.f;
#pragma unroll 1
    for (int p0 = 0; p0 < sp.P; p0 += 128) {              // 4 cells per lane in flight
      float4 va[4], vb[4];
      int cellidx[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int p = p0 + q * 32 + lane;
        va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q];
        cellidx[q] = -1;
        if (p < sp.P) {
          const float2 tb = s_tab[p];
          const int r = lattice_index(tb.x, sc, sp.res, oy);
          const int c = lattice_index(tb.y, sc, sp.res, ox);
          const ushort2 cell = s_cell[p];
          int th = cell.x + shift;                         // scan row (m + s) mod n_theta pairs with map row m
          if (th >= n_theta) th -= n_theta;
          cellidx[q] = cell.y + th;
          if (r >= 0 && r < sp.rows && c >= 0 && c < sp.cols)      // out of bounds: value 0, mask 1
            ldg256f(map_bytes + ((long long)r * sp.cols + c) * 32, va[q], vb[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (cellidx[q] >= 0) {
          const float4* sv = reinterpret_cast<const float4*>(s_scan + (size_t)cellidx[q] * 8);
          const float4 s0 = sv[0], s1 = sv[1];
          acc_c = fmaf(va[q].x, s0.x, acc_c); acc_c = fmaf(va[q].y, s0.y, acc_c);
          acc_c = fmaf(va[q].z, s0.z, acc_c); acc_c = fmaf(va[q].w, s0.w, acc_c);
          acc_c = fmaf(vb[q].x, s1.x, acc_c); acc_c = fmaf(vb[q].y, s1.y, acc_c); acc_c = fmaf(vb[q].z, s1.z, acc_c);
          acc_n = fmaf(vb[q].w, s1.w, acc_n);
          acc_k += vb[q].w;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
      acc_n += __shfl_xor_sync(0xffffffffu, acc_n, o);
      acc_k += __shfl_xor_sync(0xffffffffu, acc_k, o);
    }
    if (lane == 0) {
      float cost;
      if ((double)TDR_FDIV(acc_k, (float)sp.P) < 0.5) cost = __int_as_float(0x7fc00000);   // :117-120
      else cost = TDR_FDIV(acc_c, acc_n);                                                 // :154
      sp.weights[i] = weight_from_cost(cost, sp.regularization);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// theta search / exhaustive grid: CTA per centre, the gathered local map lives in registers
// shared: scan_pack (P*8 floats) | partial sums [n_shifts][warps][2] | misc
// ------------------------------------------------------------------------------------------------
template <int THREADS, int JMAX>
__global__ void __launch_bounds__(THREADS) k_score_search(ScoreParams sp, int grid_mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int WARPS = THREADS / 32;
  if (sp.only_if_bailed && *sp.only_if_bailed == 0) return;
  float* s_scan = reinterpret_cast<float*>(smem);
  float* s_part = s_scan + (size_t)sp.P * 8;               // [n_shifts][WARPS][2]
  float* s_kc = s_part + (size_t)sp.n_shifts * WARPS * 2;   // [WARPS]
  for (int i = threadIdx.x; i < sp.P * 2; i += THREADS)
    reinterpret_cast<float4*>(s_scan)[i] = ldg4(sp.scan_pack + (size_t)i * 4);

  // this thread's lattice points never change: p = tid + THREADS*j
  float ty[JMAX], tx[JMAX];
  int cth[JMAX], crb[JMAX];
#pragma unroll
  for (int j = 0; j < JMAX; j++) {
    int p = threadIdx.x + THREADS * j;
    if (p < sp.P) {
      ty[j] = sp.tab[2 * p]; tx[j] = sp.tab[2 * p + 1];
      int r = p / sp.n_theta;
      cth[j] = p - r * sp.n_theta; crb[j] = r * sp.n_theta;
    } else { ty[j] = 0.f; tx[j] = 0.f; cth[j] = -1; crb[j] = 0; }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const char* map_bytes = reinterpret_cast<const char*>(sp.map);

  for (long long i = blockIdx.x; i < sp.n; i += gridDim.x) {
    float cx, cy, sc;
    if (grid_mode) { cx = sp.centers[2 * i]; cy = sp.centers[2 * i + 1]; sc = sp.grid_scale; }
    else {
      if (sp.have_init[i]) continue;                       // block-uniform
      if (!particle_centre(sp, i, &cx, &cy, &sc)) { if (threadIdx.x == 0) sp.weights[i] = 0.f; continue; }
    }
    const float oy = TDR_FDIV(cy, sp.resolution), ox = TDR_FDIV(cx, sp.resolution);
    float4 ma[JMAX], mb[JMAX];
    float kc = 0.f;
#pragma unroll
    for (int j = 0; j < JMAX; j++) {
      ma[j] = make_float4(0.f, 0.f, 0.f, 0.f); mb[j] = ma[j];
      if (cth[j] >= 0) {
        int r = lattice_index(ty[j], sc, sp.res, oy);
        int c = lattice_index(tx[j], sc, sp.res, ox);
        if (r >= 0 && r < sp.rows && c >= 0 && c < sp.cols) {
          const char* px = map_bytes + ((long long)r * sp.cols + c) * 32;
          ma[j] = ldg4(px); mb[j] = ldg4(px + 16);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < JMAX; j++) kc += mb[j].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kc += __shfl_xor_sync(0xffffffffu, kc, o);
    if (lane == 0) s_kc[warp] = kc;

    for (int k = 0; k < sp.n_shifts; k++) {
      const int shift = sp.shifts[k];
      float acc_c = 0.f, acc_n = 0.f;
#pragma unroll
      for (int j = 0; j < JMAX; j++) {
        if (cth[j] >= 0) {
          int th = cth[j] + shift;
          if (th >= sp.n_theta) th -= sp.n_theta;
          const float4* sv = reinterpret_cast<const float4*>(s_scan + ((size_t)(crb[j] + th)) * 8);
          float4 s0 = sv[0], s1 = sv[1];
          acc_c = fmaf(ma[j].x, s0.x, acc_c); acc_c = fmaf(ma[j].y, s0.y, acc_c);
          acc_c = fmaf(ma[j].z, s0.z, acc_c); acc_c = fmaf(ma[j].w, s0.w, acc_c);
          acc_c = fmaf(mb[j].x, s1.x, acc_c); acc_c = fmaf(mb[j].y, s1.y, acc_c);
          acc_c = fmaf(mb[j].z, s1.z, acc_c);
          acc_n = fmaf(mb[j].w, s1.w, acc_n);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
        acc_n += __shfl_xor_sync(0xffffffffu, acc_n, o);
      }
      if (lane == 0) { s_part[((size_t)k * WARPS + warp) * 2] = acc_c; s_part[((size_t)k * WARPS + warp) * 2 + 1] = acc_n; }
    }
    __syncthreads();
    // finalise: thread k owns shift k (fixed summation order -> deterministic)
    float known = 0.f;
    for (int w = 0; w < WARPS; w++) known += s_kc[w];
    const bool unknown = (double)TDR_FDIV(known, (float)sp.P) < 0.5;
    for (int k = threadIdx.x; k < sp.n_shifts; k += THREADS) {
      float c = 0.f, nrm = 0.f;
      for (int w = 0; w < WARPS; w++) { c += s_part[((size_t)k * WARPS + w) * 2]; nrm += s_part[((size_t)k * WARPS + w) * 2 + 1]; }
      float cost = unknown ? __int_as_float(0x7fc00000) : TDR_FDIV(c, nrm);
      if (grid_mode) sp.costs[i * sp.n_shifts + k] = cost;
      s_part[(size_t)k * WARPS * 2] = cost;               // slot (k, warp 0, 0) is only read by thread k above
    }
    __syncthreads();
    if (!grid_mode && threadIdx.x == 0) {
      float best = 3.402823466e+38f, best_theta = 0.f;      // state_particle.cpp:193-204
      for (int k = 0; k < sp.n_shifts; k++) {
        float c = s_part[(size_t)k * WARPS * 2];
        if (c < best) { best = c; best_theta = sp.thetas[k]; }
      }
      sp.theta[i] = best_theta;
      sp.have_init[i] = 1;
      sp.weights[i] = weight_from_cost(best, sp.regularization);
    }
    __syncthreads();
  }
}

// a7 as a stand-alone operator (ActiveLocalizer / debug / parity): n centres -> dists n x C x P, mask n x P
__global__ void k_local_polar(const MapPixel* __restrict__ map, int rows, int cols, int C, float resolution,
                              const float* __restrict__ tab, int P, const float* __restrict__ centers, int n,
                              float scale, float res, float* __restrict__ dists, uint8_t* __restrict__ mask) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (p >= P || i >= n) return;
  float oy = TDR_FDIV(centers[2 * i + 1], resolution), ox = TDR_FDIV(centers[2 * i], resolution);
  int r = lattice_index(tab[2 * p], scale, res, oy);
  int c = lattice_index(tab[2 * p + 1], scale, res, ox);
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r >= 0 && r < rows && c >= 0 && c < cols) {
    const float4* px = reinterpret_cast<const float4*>(map + (size_t)r * cols + c);
    float4 a = px[0], b = px[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  for (int k = 0; k < C; k++) dists[((size_t)i * C + k) * P + p] = v[k];
  mask[(size_t)i * P + p] = v[7] != 0.f ? 0 : 1;
}

// TopDownMapPolar::getLocalGeoMap (top_down_map_polar.cpp:55-76): the same gather on the two geometric
// distance layers (col-major planar), no mask; n centres -> n x 2 x P
__global__ void k_local_geo_polar(const float* __restrict__ geo, int rows, int cols, float resolution, const float* __restrict__ tab,
                                  int P, const float* __restrict__ centers, int n, float scale, float res, float* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (p >= P || i >= n) return;
  float oy = TDR_FDIV(centers[2 * i + 1], resolution), ox = TDR_FDIV(centers[2 * i], resolution);
  int r = lattice_index(tab[2 * p], scale, res, oy);
  int c = lattice_index(tab[2 * p + 1], scale, res, ox);
  const bool in = r >= 0 && r < rows && c >= 0 && c < cols;
  const size_t L = (size_t)rows * cols;
#pragma unroll
  for (int k = 0; k < 2; k++) out[((size_t)i * 2 + k) * P + p] = in ? geo[(size_t)k * L + (size_t)c * rows + r] : 0.f;
}

// ActiveLocalizer::computeTotalDifference (active_localizer.cpp:7-20) for every candidate relative position at once:
// maps[cfg][i][c][P] are the gathered local maps of prediction i (UN-rotated); the rows of prediction i are rotated
// down by shifts[i] (getLocalMap :38-41) on the fly.  totals[cfg] = sum over pairs i > j, classes, cells of
// |L_i - L_j|, accumulated in double (the reference: fp32 Eigen sums; compared at 1e-5 relative).
__global__ void k_active_pairwise(const float* __restrict__ maps, const int* __restrict__ shifts, int n, int C, int n_theta, int n_r,
                                  float* __restrict__ totals) {
  const int cfg = blockIdx.x, P = n_theta * n_r;
  const float* base = maps + (size_t)cfg * n * C * P;
  double acc = 0.0;
  for (int i = 1; i < n; i++) {
    const int si = shifts[i];
    for (int j = 0; j < i; j++) {
      const int sj = shifts[j];
      for (int q = threadIdx.x; q < C * P; q += blockDim.x) {
        const int c = q / P, p = q - c * P, col = p / n_theta, r = p - col * n_theta;
        const int ri = r < si ? r + n_theta - si : r - si, rj = r < sj ? r + n_theta - sj : r - sj;
        const float a = base[((size_t)i * C + c) * P + col * n_theta + ri], b = base[((size_t)j * C + c) * P + col * n_theta + rj];
        acc += (double)fabsf(TDR_FSUB(a, b));
      }
    }
  }
  __shared__ double s_acc[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_acc[w];
    totals[cfg] = (float)t;
  }
}

// a8: Cartesian local map (top_down_map.cpp:429-459 with samplePts :367-389)
__device__ __forceinline__ float linspaced(int size, float sres, int i) {
  float low = (float)((double)TDR_FMUL(-sres, (float)(size - 1)) / 2.);
  float high = (float)((double)TDR_FMUL(sres, (float)(size - 1)) / 2.);
  if (size == 1) return low;
  float step = TDR_FDIV(TDR_FSUB(high, low), (float)(size - 1));
  bool flip = fabsf(high) < fabsf(low);
  int size1 = size - 1;
  if (flip) return (i == 0) ? low : TDR_FSUB(high, TDR_FMUL((float)(size1 - i), step));
  return (i == size1) ? high : TDR_FADD(low, TDR_FMUL((float)i, step));
}

__global__ void k_local_cart(const MapPixel* __restrict__ map, int rows, int cols, int C, float resolution, float cx,
                             float cy, float cr, float sr, float res, int out_rows, int out_cols,
                             float* __restrict__ dists, uint8_t* __restrict__ mask) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int P = out_rows * out_cols;
  if (p >= P) return;
  float sres = TDR_FDIV(res, resolution);
  float c0 = TDR_FDIV(cx, resolution), c1 = TDR_FDIV(cy, resolution);
  float x = linspaced(out_rows, sres, p % out_rows);
  float y = linspaced(out_cols, sres, p / out_rows);
  float xr = TDR_FADD(TDR_FMUL(cr, x), TDR_FMUL(-sr, y));
  float yr = TDR_FADD(TDR_FMUL(sr, x), TDR_FMUL(cr, y));
  xr = TDR_FADD(xr, c1); yr = TDR_FADD(yr, c0);
  int r = f2i_x86(round_half_away(xr)), c = f2i_x86(round_half_away(yr));
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r >= 0 && r < rows && c >= 0 && c < cols) {
    const float4* px = reinterpret_cast<const float4*>(map + (size_t)r * cols + c);
    float4 a = px[0], b = px[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  for (int k = 0; k < C; k++) dists[(size_t)k * P + p] = v[k];
  mask[p] = v[7] != 0.f ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
static const int SEARCH_THREADS = 512;
static const int SEARCH_JMAX = 5;     // 512 * 5 = 2560 >= P = 2500

static int fill_params(tdr_ctx* ctx, float res, ScoreParams* sp) {
  TDR_REQUIRE(ctx->have_map, TDR_ESTATE, "no map");
  TDR_REQUIRE(ctx->have_tab, TDR_ESTATE, "no polar table");
  TDR_REQUIRE(ctx->have_scan, TDR_ESTATE, "no scan images");
  TDR_REQUIRE(ctx->have_params, TDR_ESTATE, "no filter params");
  TDR_REQUIRE(ctx->scan_theta == ctx->n_theta && ctx->scan_r == ctx->n_r, TDR_EINVAL,
              "scan image %dx%d does not match the polar table %dx%d", ctx->scan_theta, ctx->scan_r, ctx->n_theta, ctx->n_r);
  TDR_REQUIRE(ctx->scan_C == ctx->C && ctx->fp.num_classes == ctx->C, TDR_EINVAL,
              "class counts differ: scan %d, map %d, params %d", ctx->scan_C, ctx->C, ctx->fp.num_classes);
  sp->map = ctx->map_px.as<MapPixel>(); sp->rows = ctx->rows; sp->cols = ctx->cols; sp->resolution = ctx->resolution;
  sp->tab = ctx->tab.as<float>(); sp->n_theta = ctx->n_theta; sp->n_r = ctx->n_r; sp->P = ctx->n_theta * ctx->n_r;
  sp->scan_pack = ctx->scan_pack.as<float>();
  sp->res = res;
  sp->force_on_map = ctx->fp.force_on_map;
  // StateParticle::width_/height_ = map->size().cast<float>() * map->resolution()  (state_particle.cpp:11,46-47)
  sp->map_w = (float)ctx->cols * ctx->resolution; sp->map_h = (float)ctx->rows * ctx->resolution;
  sp->scale_gate = ctx->fp.fixed_scale < 0 ? 1 : 0;                                   // :169
  sp->scale_lo = pow(10.0, (double)ctx->fp.scale_log_min); sp->scale_hi = pow(10.0, (double)ctx->fp.scale_log_max);
  sp->regularization = ctx->fp.regularization;
  sp->thetas = ctx->d_search_thetas.as<float>(); sp->shifts = ctx->d_search_shifts.as<int32_t>();
  sp->n_shifts = (int)ctx->search_shifts.size();
  sp->centers = nullptr; sp->grid_scale = 1.f; sp->costs = nullptr; sp->only_if_bailed = nullptr;
  return TDR_OK;
}

static int score_particles_impl(tdr_ctx* ctx, float res) {
  ScoreParams sp;
  if (int e = fill_params(ctx, res, &sp)) return e;
  tdr::Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(pt.n > 0, TDR_ESTATE, "no particles");
  sp.init_x = pt.init_x.as<float>(); sp.init_y = pt.init_y.as<float>(); sp.dx = pt.dx.as<float>(); sp.dy = pt.dy.as<float>();
  sp.theta = pt.theta.as<float>(); sp.scale = pt.scale.as<float>(); sp.have_init = pt.have_init.as<uint8_t>();
  sp.n = pt.n;
  if (int e = ctx->weights.reserve((size_t)pt.n * 4)) return e;
  sp.weights = ctx->weights.as<float>();
  ctx->n_weights = pt.n;
  const int P = sp.P;
  TDR_REQUIRE(P <= 65535 && P <= SEARCH_THREADS * SEARCH_JMAX, TDR_EUNSUPPORTED, "polar image of %d cells is too large (max %d)", P, SEARCH_THREADS * SEARCH_JMAX);
  TDR_SMEM_OPTIN(ctx, OPTIN_SCORE_TRACK, k_score_track, 200 * 1024);
  TDR_SMEM_OPTIN(ctx, OPTIN_SCORE_SEARCH, (k_score_search<SEARCH_THREADS, SEARCH_JMAX>), 200 * 1024);
  if (int e = sync_uninit(ctx)) return e;
  // a gated particle (force_on_map / scale range) returns before the search and keeps have_init = 0
  // (state_particle.cpp:163-176): with a gate switched on the set is recounted after the search instead of assumed clean
  const bool gates_on = ctx->fp.force_on_map != 0 || ctx->fp.fixed_scale < 0;
  // particles that already have a heading are tracked (one shift); the rest run the theta search,
  // which sets theta / have_init (state_particle.cpp:195-206).  Track first: it only READS have_init.
  const size_t track_smem = (size_t)P * 32 + (size_t)P * 8 + (size_t)P * 4;
  if (ctx->n_uninit < pt.n) {
    // LARGE tracked sets go through the tensor-core ring kernel as well: it computes all n_theta shifts per particle in
    // spatially sorted tiles and each particle keeps the column of its own heading — 62.7 ms -> ~9 ms for 1e6 tracked
    // particles (k_score_track walks the particles in resampling order, one warp each, and is DRAM-latency-bound there:
    // profiles/r02_tracking_1m.txt).  Small sets (cfg2: 1e4) stay on k_score_track.
    bool tracked_by_mma = false;
    const long long n_init = pt.n - ctx->n_uninit;
    if (sp.n_shifts > 0 && (ctx->score_impl == 2 || (ctx->score_impl == 0 && n_init * ctx->count_scale >= 65536))) {
      // the integer kernel in passes of 40 row shifts first (every particle gathered once); else the fp16 ring kernel
      if (ctx->mma_kernel == 0 || ctx->mma_kernel == 3) { if (int e = score_mma_i8_track(ctx, res, &tracked_by_mma)) return e; }
      if (!tracked_by_mma) { if (int e = score_mma(ctx, res, false, pt.n, 1.f, sp.shifts, ctx->search_shifts.data(), sp.n_shifts, &tracked_by_mma, true)) return e; }
    }
    ScoreParams st = sp;
    if (tracked_by_mma) st.only_if_bailed = reinterpret_cast<const int*>(ctx->scal.as<float>() + SC_MMA_BAILED);   // guarded fallback
    const long long warps = pt.n;
    long long ctas = (warps + 15) / 16;
    if (ctas > (long long)ctx->sm_count * 2) ctas = (long long)ctx->sm_count * 2;
    k_score_track<<<(unsigned)ctas, 512, track_smem, ctx->stream>>>(st);
    count_launch(ctx);
  }
  if (ctx->n_uninit > 0) {
    TDR_REQUIRE(sp.n_shifts > 0, TDR_ESTATE, "theta-search list not set (tdr_pf_set_search)");
    // large searches go to the tensor cores (score_mma.cu); small ones stay on the CUDA cores
    bool used = false;
    if (ctx->score_impl == 2 || (ctx->score_impl == 0 && (long long)ctx->n_uninit * ctx->count_scale >= 4096)) {
      // short candidate lists: streamed-operand kernel (two pipelines per SM); long ones: all-shifts ring kernel
      // integer 16-byte records where their error bound and the count range allow (score_mma_i8.cu)
      if (ctx->mma_kernel == 0 || ctx->mma_kernel == 3) { if (int e = score_mma_i8(ctx, res, sp.shifts, sp.n_shifts, &used)) return e; }
      if (!used && ctx->mma_kernel != 2) { if (int e = score_mma_list(ctx, res, false, pt.n, 1.f, sp.shifts, sp.n_shifts, &used)) return e; }
      if (!used) { if (int e = score_mma(ctx, res, false, pt.n, 1.f, sp.shifts, ctx->search_shifts.data(), sp.n_shifts, &used)) return e; }
    }
    const size_t smem = (size_t)P * 32 + (size_t)sp.n_shifts * (SEARCH_THREADS / 32) * 8 + 256;
    const long long ctas = sp.n < (long long)ctx->sm_count * 8 ? sp.n : (long long)ctx->sm_count * 8;
    // behind a tensor-core launch the CUDA-core kernel runs guarded: it leaves at once unless that kernel found scan
    // counts above 2048 (not exact in fp16) ON THE DEVICE and left the search to this one — no host round trip
    if (used) sp.only_if_bailed = reinterpret_cast<const int*>(ctx->scal.as<float>() + SC_MMA_BAILED);
    k_score_search<SEARCH_THREADS, SEARCH_JMAX><<<(unsigned)ctas, SEARCH_THREADS, smem, ctx->stream>>>(sp, 0);
    count_launch(ctx);
    TDR_CUDA(cudaGetLastError());
    if (gates_on) return recount_uninit(ctx, pt);
    ctx->n_uninit = 0;
  }
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int score_grid(tdr_ctx* ctx, long long n, float scale, float res) {
  ScoreParams sp;
  if (int e = fill_params(ctx, res, &sp)) return e;
  sp.n = n; sp.centers = ctx->grid_centers.as<float>(); sp.grid_scale = scale; sp.costs = grid_costs_ptr(ctx);
  sp.shifts = ctx->grid_shifts.as<int32_t>(); sp.n_shifts = ctx->grid_shifts_n; sp.thetas = nullptr;
  sp.init_x = sp.init_y = sp.dx = sp.dy = nullptr; sp.theta = nullptr; sp.scale = nullptr; sp.have_init = nullptr; sp.weights = nullptr;
  const int P = sp.P;
  bool used = false;
  ctx->grid_key_valid = false;
  if (ctx->score_impl == 2 || ctx->grid_n_peers || (ctx->score_impl == 0 && n >= 4096)) {
    if (ctx->mma_kernel == 1 && !ctx->grid_n_peers) { if (int e = score_mma_list(ctx, res, true, n, scale, sp.shifts, sp.n_shifts, &used)) return e; }
    if (!used) { if (int e = score_mma(ctx, res, true, n, scale, sp.shifts, ctx->grid_shifts_host.data(), sp.n_shifts, &used)) return e; }
    if (!used && ctx->mma_kernel != 1 && !ctx->grid_n_peers) { if (int e = score_mma_list(ctx, res, true, n, scale, sp.shifts, sp.n_shifts, &used)) return e; }
  }
  // the tensor-core kernels check the fp16 exactness of the scan counts on the device; behind one of them the CUDA-core
  // kernel is launched guarded (runs only if that one bailed out).  The fused peer all-gather has no such fallback:
  // tdr_grid_best_key reports a bail-out there.
  if (used && (ctx->grid_n_peers || P > SEARCH_THREADS * SEARCH_JMAX)) return TDR_OK;
  if (used) sp.only_if_bailed = reinterpret_cast<const int*>(ctx->scal.as<float>() + SC_MMA_BAILED);
  TDR_REQUIRE(ctx->grid_n_peers == 0, TDR_EUNSUPPORTED, "the fused peer all-gather needs the tensor-core ring kernel (n_theta <= 112 and even, distinct shifts)");
  TDR_REQUIRE(P <= SEARCH_THREADS * SEARCH_JMAX, TDR_EUNSUPPORTED, "polar image too large");
  TDR_SMEM_OPTIN(ctx, OPTIN_SCORE_SEARCH, (k_score_search<SEARCH_THREADS, SEARCH_JMAX>), 200 * 1024);
  size_t smem = (size_t)P * 32 + (size_t)sp.n_shifts * (SEARCH_THREADS / 32) * 8 + 256;
  long long ctas = n < (long long)ctx->sm_count * 8 ? n : (long long)ctx->sm_count * 8;
  k_score_search<SEARCH_THREADS, SEARCH_JMAX><<<(unsigned)ctas, SEARCH_THREADS, smem, ctx->stream>>>(sp, 1);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int local_polar(tdr_ctx* ctx, const float* dev_centers, int n, float scale, float res, float* dev_dists, uint8_t* dev_mask) {
  int P = ctx->n_theta * ctx->n_r;
  dim3 grd((P + 127) / 128, n);
  k_local_polar<<<grd, 128, 0, ctx->stream>>>(ctx->map_px.as<MapPixel>(), ctx->rows, ctx->cols, ctx->C, ctx->resolution,
                                              ctx->tab.as<float>(), P, dev_centers, n, scale, res, dev_dists, dev_mask);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int local_geo_polar(tdr_ctx* ctx, const float* dev_centers, int n, float scale, float res, float* dev_geo) {
  if (int e = map_geo_resident(ctx)) return e;
  int P = ctx->n_theta * ctx->n_r;
  dim3 grd((P + 127) / 128, n);
  k_local_geo_polar<<<grd, 128, 0, ctx->stream>>>(ctx->geo_planar.as<float>(), ctx->rows, ctx->cols, ctx->resolution, ctx->tab.as<float>(), P,
                                                  dev_centers, n, scale, res, dev_geo);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int active_pairwise(tdr_ctx* ctx, const float* dev_maps, const int* dev_shifts, int n_cfg, int n_preds, float* dev_totals) {
  k_active_pairwise<<<n_cfg, 256, 0, ctx->stream>>>(dev_maps, dev_shifts, n_preds, ctx->C, ctx->n_theta, ctx->n_r, dev_totals);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int local_cart(tdr_ctx* ctx, float cx, float cy, float rot, float res, int out_rows, int out_cols, float* dev_dists,
               uint8_t* dev_mask) {
  int P = out_rows * out_cols;
  k_local_cart<<<(P + 127) / 128, 128, 0, ctx->stream>>>(ctx->map_px.as<MapPixel>(), ctx->rows, ctx->cols, ctx->C,
                                                         ctx->resolution, cx, cy, cosf(rot), sinf(rot), res, out_rows,
                                                         out_cols, dev_dists, dev_mask);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

int score_particles(tdr_ctx* ctx, float res) {
  const int e = score_particles_impl(ctx, res);
  const int j = score_i8_prepare_join(ctx);        // a head start the chosen kernels did not consume ends here, used or not
  return e ? e : j;
}

}  // namespace tdr
