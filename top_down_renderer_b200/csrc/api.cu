// api.cu — the extern "C" surface declared in include/tdr.h.
#include <atomic>
#include <cmath>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "tdr_ctx.cuh"
#include "tdr_math.cuh"

namespace tdr {

static thread_local char g_err[512] = "";
uint64_t next_tab_id() { static std::atomic<uint64_t> next{0}; return ++next; }
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int Particles::reserve(int64_t cap) {
  size_t b = (size_t)cap * 4;
  if (int e = init_x.reserve(b)) return e;
  if (int e = init_y.reserve(b)) return e;
  if (int e = dx.reserve(b)) return e;
  if (int e = dy.reserve(b)) return e;
  if (int e = theta.reserve(b)) return e;
  if (int e = scale.reserve(b)) return e;
  if (int e = last_dist.reserve(b)) return e;
  if (int e = have_init.reserve((size_t)cap)) return e;
  return TDR_OK;
}
void Particles::release() {
  init_x.release(); init_y.release(); dx.release(); dy.release(); theta.release(); scale.release();
  last_dist.release(); have_init.release(); n = 0;
}

struct SoA { float *ix, *iy, *dx, *dy, *th, *sc; uint8_t* hi; };
static SoA soa_of(Particles& p) {
  SoA s; s.ix = p.init_x.as<float>(); s.iy = p.init_y.as<float>(); s.dx = p.dx.as<float>(); s.dy = p.dy.as<float>();
  s.th = p.theta.as<float>(); s.sc = p.scale.as<float>(); s.hi = p.have_init.as<uint8_t>();
  return s;
}

// 28-byte State records <-> SoA (7 words per record; have_init is byte 24)
__global__ void k_aos_to_soa(const uint32_t* __restrict__ aos, long long n, SoA s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t* r = aos + i * 7;
    s.ix[i] = __uint_as_float(r[0]); s.iy[i] = __uint_as_float(r[1]); s.dx[i] = __uint_as_float(r[2]);
    s.dy[i] = __uint_as_float(r[3]); s.th[i] = __uint_as_float(r[4]); s.sc[i] = __uint_as_float(r[5]);
    s.hi[i] = (uint8_t)((r[6] & 0xffu) ? 1 : 0);
  }
}
__global__ void k_soa_to_aos(SoA s, long long n, uint32_t* __restrict__ aos) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t* r = aos + i * 7;
    r[0] = __float_as_uint(s.ix[i]); r[1] = __float_as_uint(s.iy[i]); r[2] = __float_as_uint(s.dx[i]);
    r[3] = __float_as_uint(s.dy[i]); r[4] = __float_as_uint(s.th[i]); r[5] = __float_as_uint(s.sc[i]);
    r[6] = s.hi[i] ? 1u : 0u;
  }
}

// multi-GPU shard block: TDR_SHARD_ROWS rows of n floats — weight, init_x, init_y, dx, dy, theta, scale,
// have_init (0/1 as float), last_dist.  One block per rank is what the all-gather moves.
struct ShardSrc { const float *w, *ix, *iy, *dx, *dy, *th, *sc, *ld; const uint8_t* hi; };
__global__ void k_pack_shard(ShardSrc s, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[i] = s.w ? s.w[i] : 0.f;
    out[n + i] = s.ix[i]; out[2 * n + i] = s.iy[i]; out[3 * n + i] = s.dx[i]; out[4 * n + i] = s.dy[i];
    out[5 * n + i] = s.th[i]; out[6 * n + i] = s.sc[i]; out[7 * n + i] = s.hi[i] ? 1.f : 0.f; out[8 * n + i] = s.ld[i];
  }
}
struct ShardDst { float *w, *ix, *iy, *dx, *dy, *th, *sc, *ld; uint8_t* hi; };
__global__ void k_unpack_all(const float* __restrict__ in, int ranks, long long n_local, ShardDst d) {
  const long long N = (long long)ranks * n_local;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    const long long g = j / n_local, i = j - g * n_local;
    const float* b = in + g * TDR_SHARD_ROWS * n_local;
    if (d.w) d.w[j] = b[i];
    d.ix[j] = b[n_local + i]; d.iy[j] = b[2 * n_local + i]; d.dx[j] = b[3 * n_local + i]; d.dy[j] = b[4 * n_local + i];
    d.th[j] = b[5 * n_local + i]; d.sc[j] = b[6 * n_local + i]; d.hi[j] = b[7 * n_local + i] != 0.f ? 1 : 0;
    d.ld[j] = b[8 * n_local + i];
  }
}

// split wire format (two all-gathers: the 8 B/particle the normalisation needs first, the states behind it)
__global__ void k_pack_split(ShardSrc s, long long n, float* __restrict__ wl, float* __restrict__ st) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    wl[i] = s.w[i]; wl[n + i] = s.ld[i];
    st[i] = s.ix[i]; st[n + i] = s.iy[i]; st[2 * n + i] = s.dx[i]; st[3 * n + i] = s.dy[i];
    st[4 * n + i] = s.th[i]; st[5 * n + i] = s.sc[i]; st[6 * n + i] = s.hi[i] ? 1.f : 0.f;
  }
}
__global__ void k_unpack_wl(const float* __restrict__ in, int ranks, long long n_local, float* __restrict__ w, float* __restrict__ ld) {
  const long long N = (long long)ranks * n_local;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    const long long g = j / n_local, i = j - g * n_local;
    const float* b = in + g * 2 * n_local;
    w[j] = b[i]; ld[j] = b[n_local + i];
  }
}
__global__ void k_unpack_states(const float* __restrict__ in, int ranks, long long n_local, ShardDst d) {
  const long long N = (long long)ranks * n_local;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
    const long long g = j / n_local, i = j - g * n_local;
    const float* b = in + g * 7 * n_local;
    d.ix[j] = b[i]; d.iy[j] = b[n_local + i]; d.dx[j] = b[2 * n_local + i]; d.dy[j] = b[3 * n_local + i];
    d.th[j] = b[4 * n_local + i]; d.sc[j] = b[5 * n_local + i]; d.hi[j] = b[6 * n_local + i] != 0.f ? 1 : 0;
  }
}

static int grid_for(tdr_ctx* ctx, long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)ctx->sm_count * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace tdr

using namespace tdr;

#define CTX_CHECK(ctx)                                                           \
  do {                                                                           \
    if (!(ctx)) { tdr::set_error("null context"); return TDR_EINVAL; }           \
    cudaError_t e__ = cudaSetDevice((ctx)->device);                              \
    if (e__ != cudaSuccess) { tdr::set_error("cudaSetDevice: %s", cudaGetErrorString(e__)); return TDR_ECUDA; } \
  } while (0)

extern "C" {

int tdr_abi_version(void) { return TDR_ABI_VERSION; }
const char* tdr_last_error(void) { return tdr::g_err; }

int tdr_create(tdr_ctx** out, int device) {
  TDR_REQUIRE(out, TDR_EINVAL, "null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    tdr::set_error("no usable CUDA device (%s); libtdr_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return TDR_ENOGPU;
  }
  TDR_REQUIRE(device >= 0 && device < count, TDR_EINVAL, "device %d out of range (%d devices)", device, count);
  TDR_CUDA(cudaSetDevice(device));
  tdr_ctx* c = new tdr_ctx();
  c->device = device;
  cudaDeviceProp prop;
  TDR_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    tdr::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    delete c;
    return TDR_ENOGPU;
  }
  if (const char* e = getenv("TDR_SEQ_IMPL")) c->seq_impl = atoi(e) == 1 ? 1 : 0;
  if (const char* e = getenv("TDR_MMA_TILES")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) c->mma_tiles = v; }
  if (const char* e = getenv("TDR_MMA_SPLIT")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) c->mma_split = v; }
  if (const char* e = getenv("TDR_MMA_SEG_SHIFT")) { int v = atoi(e); if (v >= 0 && v <= 5) c->mma_seg_shift = v; }
  if (const char* e = getenv("TDR_MMA_A_TMEM")) c->mma_a_tmem = atoi(e) ? 1 : 0;
  if (const char* e = getenv("TDR_MMA_RING_CFG")) c->mma_ring_cfg = atoi(e);
  if (const char* e = getenv("TDR_MMA_KERNEL")) c->mma_kernel = atoi(e);
  if (const char* e = getenv("TDR_MMA_CTAS")) c->mma_ctas = atoi(e);
  if (const char* e = getenv("TDR_MMA_I8")) { int v = atoi(e); if (v >= 0 && v <= 2) c->mma_i8 = v; }
  if (const char* e = getenv("TDR_MMA_I8_CFG")) c->mma_i8_cfg = atoi(e);
  if (const char* e = getenv("TDR_MMA_TEX")) c->mma_tex = atoi(e) ? 1 : 0;
  if (const char* e = getenv("TDR_MMA_SKIP_RINGS")) c->mma_skip_rings = atoi(e) ? 1 : 0;
  if (const char* e = getenv("TDR_MMA_SORT")) { int v = atoi(e); if (v >= 0 && v <= 1) c->mma_sort = v; }
  if (const char* e = getenv("TDR_MMA_GRID_CAP")) { int v = atoi(e); if (v > 0) c->mma_grid_cap = v; }
  if (const char* e = getenv("TDR_MMA_ST_SHIFT")) { int v = atoi(e); if (v >= 5 && v <= 16) c->mma_st_shift = v; }
  TDR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  if (int r = c->scal.reserve(SC_TOTAL * 4)) { delete c; return r; }
  TDR_CUDA(cudaMemsetAsync(c->scal.p, 0, SC_TOTAL * 4, c->stream));
  if (int r = c->d_cw.reserve(64)) { delete c; return r; }
  if (int r = c->uninit_dev.reserve(16)) { delete c; return r; }
  TDR_CUDA(cudaEventCreateWithFlags(&c->uninit_ev, cudaEventDisableTiming));
  TDR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->uninit_pin), 16));
  TDR_CUDA(cudaEventCreateWithFlags(&c->scan_max_ev, cudaEventDisableTiming));
  TDR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->scan_max_pin), 16));
  *c->scan_max_pin = 0;
  TDR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->grid_key_pin), 32));
  memset(c->grid_key_pin, 0, 32);
  if (const char* e = getenv("TDR_EDT_IMPL")) c->edt_impl = atoi(e);
  if (const char* e = getenv("TDR_EDT_BAND")) { int v = atoi(e); if (v >= 16 && v <= 1024) c->edt_band = v; }
  if (const char* e = getenv("TDR_GRID_SELF_ONLY")) c->grid_self_only = atoi(e) != 0;
  if (const char* e = getenv("TDR_GRID_STORE_HINT")) c->grid_store_hint = atoi(e) ? 1 : 0;
  *out = c;
  return TDR_OK;
}

void tdr_destroy(tdr_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  tdr_shard_finalize(c);
  if (c->stream) { cudaStreamSynchronize(c->stream); }
  tdr::DevBuf* bufs[] = {&c->map_px, &c->seedbits, &c->edt_g, &c->scratch, &c->scratch2, &c->tab, &c->pts, &c->lut,
                         &c->scan_img, &c->scan_pack, &c->hist, &c->d_search_thetas, &c->d_search_shifts, &c->weights,
                         &c->prefix, &c->idx, &c->scal, &c->pose_tmp, &c->grid_centers, &c->grid_costs,
                         &c->grid_shifts, &c->d_cw, &c->map16, &c->map16g, &c->geo_planar, &c->tab_scaled, &c->grid_key, &c->seq_ws, &c->scan_op, &c->bin_counts, &c->perm};
  for (auto* b : bufs) b->release();
  for (void* p : c->grid_opened) cudaIpcCloseMemHandle(p);
  c->grid_full.release();
  c->part[0].release(); c->part[1].release(); c->ckpt.release(); c->all.release();
  c->pin.release();
  c->uninit_dev.release();
  c->raw_weights.release();
  c->ident_shifts.release();
  if (c->uninit_ev) cudaEventDestroy(c->uninit_ev);
  if (c->uninit_pin) cudaFreeHost(c->uninit_pin);
  if (c->map8_tex) cudaDestroyTextureObject((cudaTextureObject_t)c->map8_tex);
  c->map8.release();
  if (c->scan_max_ev) cudaEventDestroy(c->scan_max_ev);
  if (c->scan_max_pin) cudaFreeHost(c->scan_max_pin);
  if (c->prep_stream) { cudaStreamSynchronize(c->prep_stream); cudaStreamDestroy(c->prep_stream); }
  if (c->prep_fork) cudaEventDestroy(c->prep_fork);
  if (c->prep_done) cudaEventDestroy(c->prep_done);
  if (c->grid_key_pin) cudaFreeHost(c->grid_key_pin);
  for (int k = 0; k <= TDR_N_STAGES; k++) if (c->stage_ev[k]) cudaEventDestroy(c->stage_ev[k]);
  for (int k = 0; k < 2; k++) { if (c->refine_copied[k]) cudaEventDestroy(c->refine_copied[k]); if (c->refine_binned[k]) cudaEventDestroy(c->refine_binned[k]); c->refine_stage[k].release(); }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int tdr_sync(tdr_ctx* ctx) { CTX_CHECK(ctx); TDR_CUDA(cudaStreamSynchronize(ctx->stream)); return TDR_OK; }
void* tdr_stream(tdr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int tdr_launch_count(tdr_ctx* ctx, int64_t* n) { TDR_REQUIRE(ctx && n, TDR_EINVAL, "null argument"); *n = ctx->launches; return TDR_OK; }

int tdr_set_score_impl(tdr_ctx* ctx, int impl) {
  TDR_REQUIRE(ctx && impl >= 0 && impl <= 2, TDR_EINVAL, "score impl must be 0 (auto), 1 (CUDA cores) or 2 (tensor cores)");
  ctx->score_impl = impl;
  return TDR_OK;
}

int tdr_profile_enable(tdr_ctx* ctx, int on) {
  CTX_CHECK(ctx);
  if (on && !ctx->stage_ev[0])
    for (int k = 0; k <= TDR_N_STAGES; k++) TDR_CUDA(cudaEventCreate(&ctx->stage_ev[k]));
  ctx->profiling = on != 0; ctx->stage_valid = false;
  return TDR_OK;
}
int tdr_profile_stage_ms(tdr_ctx* ctx, float ms[TDR_N_STAGES]) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ms, TDR_EINVAL, "null output");
  TDR_REQUIRE(ctx->profiling && ctx->stage_valid, TDR_ESTATE, "no profiled step (tdr_profile_enable, then tdr_step / tdr_pf_update)");
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < TDR_N_STAGES; k++) TDR_CUDA(cudaEventElapsedTime(&ms[k], ctx->stage_ev[k], ctx->stage_ev[k + 1]));
  return TDR_OK;
}

// ---- map ------------------------------------------------------------------------------------------
int tdr_map_set_class_image(tdr_ctx* ctx, const uint8_t* img, int h_img, int w_img, int stride,
                            const int32_t* flatten_lut, int n_lut, int num_classes, float resolution) {
  CTX_CHECK(ctx);
  return map_set_class_image(ctx, img, h_img, w_img, stride, flatten_lut, n_lut, num_classes, resolution);
}
int tdr_map_set_binary_layers(tdr_ctx* ctx, const float* layers, int rows, int cols, int num_classes, float resolution) {
  CTX_CHECK(ctx);
  return map_set_binary_layers(ctx, layers, rows, cols, num_classes, resolution);
}
int tdr_map_set_dist_layers(tdr_ctx* ctx, const float* layers, const uint8_t* mask, int rows, int cols, int num_classes,
                            float resolution) {
  CTX_CHECK(ctx);
  return map_set_dist_layers(ctx, layers, mask, rows, cols, num_classes, resolution);
}
int tdr_map_set_polygons(tdr_ctx* ctx, const float* verts_xy, const int32_t* poly_start, const int32_t* poly_class, int n_poly,
                         int map_w, int map_h, float rot, int num_classes, float resolution, const int32_t* exclusive,
                         int n_exclusive, float* layers_out) {
  CTX_CHECK(ctx);
  return map_set_polygons(ctx, verts_xy, poly_start, poly_class, n_poly, map_w, map_h, rot, num_classes, resolution, exclusive,
                          n_exclusive, layers_out);
}
int tdr_map_get_layers(tdr_ctx* ctx, float* layers, uint8_t* mask) { CTX_CHECK(ctx); return map_get_layers(ctx, layers, mask); }
int tdr_map_get_geo_layers(tdr_ctx* ctx, float* geo) { CTX_CHECK(ctx); return map_get_geo_layers(ctx, geo); }
int tdr_map_info(tdr_ctx* ctx, int* rows, int* cols, int* num_classes, float* resolution) {
  TDR_REQUIRE(ctx, TDR_EINVAL, "null context");
  TDR_REQUIRE(ctx->have_map, TDR_ESTATE, "no map");
  if (rows) *rows = ctx->rows;
  if (cols) *cols = ctx->cols;
  if (num_classes) *num_classes = ctx->C;
  if (resolution) *resolution = ctx->resolution;
  return TDR_OK;
}

int tdr_map_set_polar_table(tdr_ctx* ctx, const float* tab, int n_theta, int n_r) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(tab && n_theta > 0 && n_r > 0, TDR_EINVAL, "bad polar table");
  TDR_REQUIRE((long long)n_theta * n_r <= 65535 && n_theta <= 65535, TDR_EUNSUPPORTED, "polar table too large");
  size_t bytes = (size_t)n_theta * n_r * 2 * 4;
  if (int e = ctx->tab.reserve(bytes)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->tab.p, tab, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->n_theta = n_theta; ctx->n_r = n_r; ctx->have_tab = true; ctx->tab_version++; ctx->tab_id = tdr::next_tab_id();
  return TDR_OK;
}

int tdr_map_local_polar(tdr_ctx* ctx, const float* centers_xy, int n, float scale, float res, float* dists, uint8_t* mask) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->have_map && ctx->have_tab, TDR_ESTATE, "map / polar table missing");
  TDR_REQUIRE(centers_xy && n > 0 && n <= 65535 && dists && mask, TDR_EINVAL, "bad arguments");
  const int P = ctx->n_theta * ctx->n_r;
  size_t db = (size_t)n * ctx->C * P * 4, mb = (size_t)n * P;
  if (int e = ctx->scratch.reserve(db + mb + 256)) return e;
  if (int e = ctx->scratch2.reserve((size_t)n * 8)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch2.p, centers_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  float* dd = ctx->scratch.as<float>();
  uint8_t* dm = ctx->scratch.as<uint8_t>() + db;
  if (int e = local_polar(ctx, ctx->scratch2.as<float>(), n, scale, res, dd, dm)) return e;
  TDR_CUDA(cudaMemcpyAsync(dists, dd, db, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(mask, dm, mb, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int tdr_map_set_geo_dist_layers(tdr_ctx* ctx, const float* geo_layers) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->have_map && geo_layers, TDR_ESTATE, "no map / null layers");
  const size_t bytes = (size_t)ctx->rows * ctx->cols * 2 * 4;
  if (int e = ctx->geo_planar.reserve(bytes)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->geo_planar.p, geo_layers, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->geo_valid = true;
  return TDR_OK;
}
int tdr_map_local_geo_polar(tdr_ctx* ctx, const float* centers_xy, int n, float scale, float res, float* geo) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->have_map && ctx->have_tab, TDR_ESTATE, "map / polar table missing");
  TDR_REQUIRE(centers_xy && n > 0 && n <= 65535 && geo, TDR_EINVAL, "bad arguments");
  const int P = ctx->n_theta * ctx->n_r;
  const size_t gb = (size_t)n * 2 * P * 4;
  if (int e = map_geo_resident(ctx)) return e;            // uses scratch2 for the seeds: before the staging below
  const size_t off_c = ((gb + 255) / 256) * 256;
  if (int e = ctx->scratch.reserve(off_c + (size_t)n * 8)) return e;
  float* dg = ctx->scratch.as<float>();
  float* dc = reinterpret_cast<float*>(ctx->scratch.as<unsigned char>() + off_c);
  TDR_CUDA(cudaMemcpyAsync(dc, centers_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (int e = local_geo_polar(ctx, dc, n, scale, res, dg)) return e;
  TDR_CUDA(cudaMemcpyAsync(geo, dg, gb, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

// ActiveLocalizer::getBestRelPos (active_localizer.cpp:45-82): the candidate loop runs on the host exactly as written
// (float theta += M_PI / 8, dist 50, 75, ... while best_diff < 6000), the gathers of every candidate x prediction and
// the pairwise differences run on the device in two launches
int tdr_active_best_rel_pos(tdr_ctx* ctx, const float* preds_xyt, int n_preds, float rel_pos[2], float* best_diff_out) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->have_map && ctx->have_tab, TDR_ESTATE, "map / polar table missing");
  TDR_REQUIRE(preds_xyt && n_preds >= 1 && n_preds <= 64 && rel_pos, TDR_EINVAL, "bad predictions");
  std::vector<float> thetas, dists;
  for (float theta = 0; theta < 2 * M_PI; theta += M_PI / 8) thetas.push_back(theta);     // :63
  for (float dist = 50; dist < 150; dist += 25) dists.push_back(dist);                     // :59,62,80
  const int T = (int)thetas.size(), D = (int)dists.size(), n_cfg = D * T, P = ctx->n_theta * ctx->n_r, C = ctx->C;
  const int n_total = n_cfg * n_preds;
  std::vector<float> centers((size_t)n_total * 2);
  std::vector<int> shifts((size_t)n_preds);
  for (int i = 0; i < n_preds; i++) {
    const float th = preds_xyt[3 * i + 2];
    int rs = static_cast<int>(std::round(th * ctx->n_theta / 2 / M_PI));                  // :32
    while (rs >= ctx->n_theta) rs -= ctx->n_theta;
    while (rs < 0) rs += ctx->n_theta;
    shifts[i] = rs;
  }
  for (int d = 0; d < D; d++)
    for (int t = 0; t < T; t++)
      for (int i = 0; i < n_preds; i++) {
        const float* pr = preds_xyt + 3 * i;
        const size_t k = ((size_t)(d * T + t) * n_preds + i) * 2;
        centers[k] = pr[0] + dists[d] * std::cos(thetas[t] + pr[2]);                       // :67
        centers[k + 1] = pr[1] + dists[d] * std::sin(thetas[t] + pr[2]);
      }
  const size_t maps_b = (size_t)n_total * C * P * 4, mask_b = (size_t)n_total * P;
  const size_t off_mask = ((maps_b + 255) / 256) * 256, off_cent = off_mask + ((mask_b + 255) / 256) * 256;
  const size_t off_shift = off_cent + (size_t)n_total * 8, off_tot = off_shift + ((size_t)n_preds * 4 + 255) / 256 * 256;
  if (int e = ctx->scratch.reserve(off_tot + (size_t)n_cfg * 4)) return e;
  unsigned char* b = ctx->scratch.as<unsigned char>();
  TDR_CUDA(cudaMemcpyAsync(b + off_cent, centers.data(), (size_t)n_total * 8, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(b + off_shift, shifts.data(), (size_t)n_preds * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (int e = local_polar(ctx, reinterpret_cast<float*>(b + off_cent), n_total, 1.f, 2.f, reinterpret_cast<float*>(b), b + off_mask)) return e;   // :29
  if (int e = active_pairwise(ctx, reinterpret_cast<float*>(b), reinterpret_cast<int*>(b + off_shift), n_cfg, n_preds,
                              reinterpret_cast<float*>(b + off_tot))) return e;
  std::vector<float> totals((size_t)n_cfg);
  TDR_CUDA(cudaMemcpyAsync(totals.data(), b + off_tot, (size_t)n_cfg * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  const int cnt = n_preds * (n_preds - 1) / 2 * C;
  float best_diff = 0, best0 = 0, best1 = 0;
  for (int d = 0; d < D && best_diff < 6000; d++)
    for (int t = 0; t < T; t++) {
      const float diff = totals[(size_t)d * T + t] / cnt;                                  // 0 / 0 = NaN for one prediction: never best
      if (diff > best_diff) { best_diff = diff; best0 = dists[d]; best1 = thetas[t]; }
    }
  rel_pos[0] = best0; rel_pos[1] = best1;
  if (best_diff_out) *best_diff_out = best_diff;
  return TDR_OK;
}

int tdr_map_local_cart(tdr_ctx* ctx, float cx, float cy, float rot, float res, int rows, int cols, float* dists, uint8_t* mask) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->have_map, TDR_ESTATE, "no map");
  TDR_REQUIRE(rows > 0 && cols > 0 && dists && mask, TDR_EINVAL, "bad arguments");
  const size_t P = (size_t)rows * cols;
  size_t db = P * ctx->C * 4;
  if (int e = ctx->scratch.reserve(db + P + 256)) return e;
  float* dd = ctx->scratch.as<float>();
  uint8_t* dm = ctx->scratch.as<uint8_t>() + db;
  if (int e = local_cart(ctx, cx, cy, rot, res, rows, cols, dd, dm)) return e;
  TDR_CUDA(cudaMemcpyAsync(dists, dd, db, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(mask, dm, P, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

// ---- scan -----------------------------------------------------------------------------------------
int tdr_scan_set_points(tdr_ctx* ctx, const void* pts, int stride_bytes, int intensity_off, int64_t n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(n >= 0 && (n == 0 || pts), TDR_EINVAL, "null points");
  TDR_REQUIRE(stride_bytes >= 8 && stride_bytes % 4 == 0 && intensity_off % 4 == 0 && intensity_off >= 0 &&
                  intensity_off + 4 <= stride_bytes, TDR_EINVAL, "bad point layout (stride %d, intensity offset %d)", stride_bytes, intensity_off);
  size_t bytes = (size_t)n * stride_bytes;
  if (int e = ctx->pts.reserve(bytes + 16)) return e;
  if (n) TDR_CUDA(cudaMemcpyAsync(ctx->pts.p, pts, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->pts_stride = stride_bytes; ctx->pts_ioff = intensity_off; ctx->n_pts = n;
  return TDR_OK;   // asynchronous: the caller keeps `pts` alive until the next tdr_sync (pinned memory recommended)
}

int tdr_scan_set_lut(tdr_ctx* ctx, const int32_t* lut, int n_lut, int num_classes) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(lut && n_lut > 0 && num_classes >= 1 && num_classes <= TDR_MAX_CLASSES, TDR_EINVAL, "bad lut");
  if (int e = ctx->lut.reserve((size_t)n_lut * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->lut.p, lut, (size_t)n_lut * 4, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->n_lut = n_lut; ctx->lut_classes = num_classes;
  return TDR_OK;
}

static int render_polar_resident(tdr_ctx* ctx, float res, float ang_res, int n_theta, int n_r) {
  const int C = ctx->lut_classes;
  if (int e = ctx->scan_img.reserve((size_t)C * n_theta * n_r * 4)) return e;
  if (int e = scan_render(ctx, true, res, ang_res, n_theta, n_r, ctx->scan_img.as<float>())) return e;
  ctx->scan_theta = n_theta; ctx->scan_r = n_r; ctx->scan_C = C; ctx->have_scan = true;
  return TDR_OK;
}

int tdr_scan_render_polar(tdr_ctx* ctx, float res, float ang_res, int n_theta, int n_r, float* imgs) {
  CTX_CHECK(ctx);
  stage_mark(ctx, TDR_STAGE_RENDER);
  if (int e = render_polar_resident(ctx, res, ang_res, n_theta, n_r)) return e;
  if (imgs) {
    TDR_CUDA(cudaMemcpyAsync(imgs, ctx->scan_img.p, (size_t)ctx->scan_C * n_theta * n_r * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return TDR_OK;
}

int tdr_scan_render_cart(tdr_ctx* ctx, float res, int rows, int cols, float* imgs) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(imgs, TDR_EINVAL, "null output");
  const int C = ctx->lut_classes;
  size_t bytes = (size_t)C * rows * cols * 4;
  if (int e = ctx->scratch2.reserve(bytes)) return e;
  if (int e = scan_render(ctx, false, res, 0.f, rows, cols, ctx->scratch2.as<float>())) return e;
  TDR_CUDA(cudaMemcpyAsync(imgs, ctx->scratch2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int tdr_scan_render_geometric_polar(tdr_ctx* ctx, int width, int height, float res, float ang_res, int n_theta, int n_r, float* imgs) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(imgs && n_theta > 0 && n_r > 0 && n_r <= 1024, TDR_EINVAL, "bad geometric image shape");
  const size_t bytes = (size_t)2 * n_theta * n_r * 4;
  if (int e = ctx->scratch2.reserve(bytes)) return e;
  if (int e = scan_render_geometric(ctx, true, res, ang_res, n_theta, n_r, width, height, ctx->scratch2.as<float>())) return e;
  TDR_CUDA(cudaMemcpyAsync(imgs, ctx->scratch2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int tdr_scan_render_geometric_cart(tdr_ctx* ctx, int width, int height, float res, int rows, int cols, float* imgs) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(imgs && rows > 0 && cols > 0, TDR_EINVAL, "bad geometric image shape");
  const size_t bytes = (size_t)2 * rows * cols * 4;
  if (int e = ctx->scratch2.reserve(bytes)) return e;
  if (int e = scan_render_geometric(ctx, false, res, 0.f, rows, cols, width, height, ctx->scratch2.as<float>())) return e;
  TDR_CUDA(cudaMemcpyAsync(imgs, ctx->scratch2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int tdr_refine_bin(tdr_ctx* ctx, const float* xy, const int32_t* cls, int64_t n, float res, float center_x, float center_y,
                   int width, int height, int num_classes, uint8_t* maps_out) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(n == 0 || (xy && cls), TDR_EINVAL, "null points");
  return refine_bin(ctx, xy, cls, n, res, center_x, center_y, width, height, num_classes, maps_out);
}

int tdr_refine_begin(tdr_ctx* ctx, float res, float center_x, float center_y, int width, int height, int num_classes) {
  CTX_CHECK(ctx);
  return refine_begin(ctx, res, center_x, center_y, width, height, num_classes);
}
int tdr_refine_add(tdr_ctx* ctx, const float* xy, const int32_t* cls, int64_t n) { CTX_CHECK(ctx); return refine_add(ctx, xy, cls, n, false); }
int tdr_refine_add_dev(tdr_ctx* ctx, const void* dev_xy, const void* dev_cls, int64_t n) {
  CTX_CHECK(ctx);
  return refine_add(ctx, reinterpret_cast<const float*>(dev_xy), reinterpret_cast<const int32_t*>(dev_cls), n, true);
}
int tdr_refine_counts(tdr_ctx* ctx, uint8_t* maps_out) { CTX_CHECK(ctx); return refine_counts(ctx, maps_out); }
int tdr_refine_rebuild_map(tdr_ctx* ctx, float resolution) { CTX_CHECK(ctx); return refine_rebuild_map(ctx, resolution); }

int tdr_scan_set_polar_images(tdr_ctx* ctx, const float* imgs, int n_theta, int n_r, int num_classes) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(imgs && n_theta > 0 && n_r > 0 && num_classes >= 1 && num_classes <= TDR_MAX_CLASSES, TDR_EINVAL, "bad images");
  size_t bytes = (size_t)num_classes * n_theta * n_r * 4;
  if (int e = ctx->scan_img.reserve(bytes)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->scan_img.p, imgs, bytes, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->scan_theta = n_theta; ctx->scan_r = n_r; ctx->scan_C = num_classes; ctx->have_scan = true;
  return TDR_OK;
}

// ---- filter ---------------------------------------------------------------------------------------
int tdr_pf_set_params(tdr_ctx* ctx, const tdr_filter_params* p) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(p && p->num_classes >= 1 && p->num_classes <= TDR_MAX_CLASSES, TDR_EINVAL, "bad filter params");
  ctx->fp = *p;
  TDR_CUDA(cudaMemcpyAsync(ctx->d_cw.p, ctx->fp.class_weights, 64, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_params = true; ctx->map16_valid = false; ctx->map8_valid = false; ctx->map16g_log2 = -1;   // class weights are folded into the fp16 map copy
  return TDR_OK;
}

int tdr_pf_set_search(tdr_ctx* ctx, const float* thetas, const int32_t* shifts, int n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(thetas && shifts && n > 0 && n <= TDR_MAX_SHIFTS, TDR_EINVAL, "bad search list");
  ctx->search_thetas.assign(thetas, thetas + n);
  ctx->search_shifts.assign(shifts, shifts + n);
  if (int e = ctx->d_search_thetas.reserve((size_t)n * 4)) return e;
  if (int e = ctx->d_search_shifts.reserve((size_t)n * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->d_search_thetas.p, thetas, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(ctx->d_search_shifts.p, shifts, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int tdr_pf_set_states(tdr_ctx* ctx, const tdr_state* states, const float* last_dist, int64_t n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(states && n > 0 && n < (1ll << 31), TDR_EINVAL, "bad states");
  static_assert(sizeof(tdr_state) == 28, "State must be 28 bytes");
  Particles& pt = ctx->part[ctx->cur];
  if (int e = pt.reserve(n)) return e;
  if (int e = ctx->scratch.reserve((size_t)n * 28)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->scratch.p, states, (size_t)n * 28, cudaMemcpyHostToDevice, ctx->stream));
  k_aos_to_soa<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(ctx->scratch.as<uint32_t>(), n, soa_of(pt));
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (last_dist) TDR_CUDA(cudaMemcpyAsync(pt.last_dist.p, last_dist, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  else TDR_CUDA(cudaMemsetAsync(pt.last_dist.p, 0, (size_t)n * 4, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t un = 0;
  for (int64_t i = 0; i < n; i++) un += states[i].have_init ? 0 : 1;
  pt.n = n; ctx->n_uninit = un; ctx->uninit_pending = false; ctx->have_argmax = false;
  return TDR_OK;
}

int tdr_pf_get_states(tdr_ctx* ctx, tdr_state* states, int64_t n) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(states && n > 0 && n <= pt.n, TDR_EINVAL, "bad state count %lld (have %lld)", (long long)n, (long long)pt.n);
  if (int e = ctx->scratch.reserve((size_t)n * 28)) return e;
  k_soa_to_aos<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(soa_of(pt), n, ctx->scratch.as<uint32_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(states, ctx->scratch.p, (size_t)n * 28, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

static int propagate_host_z(tdr_ctx* ctx, float tx, float ty, float omega, int scale_freeze, float pos_cov, float theta_cov,
                            const float* z, bool z_on_device, int64_t n) {
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(z && n == pt.n, TDR_EINVAL, "need 4 variates for each of the %lld resident particles", (long long)pt.n);
  const float* dz = z;
  if (!z_on_device) {
    if (int e = ctx->scratch.reserve((size_t)n * 16)) return e;
    TDR_CUDA(cudaMemcpyAsync(ctx->scratch.p, z, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    dz = ctx->scratch.as<float>();
  }
  if (int e = propagate(ctx, tx, ty, omega, scale_freeze, pos_cov, theta_cov, dz, false, 0, 0, nullptr)) return e;
  if (!z_on_device) TDR_CUDA(cudaStreamSynchronize(ctx->stream));      // the caller's pageable buffer may go away
  return TDR_OK;
}
int tdr_pf_propagate(tdr_ctx* ctx, float trans_x, float trans_y, float omega, int scale_freeze, float pos_cov, float theta_cov,
                     const float* z, int64_t n) {
  CTX_CHECK(ctx);
  return propagate_host_z(ctx, trans_x, trans_y, omega, scale_freeze, pos_cov, theta_cov, z, false, n);
}
int tdr_pf_propagate_dev(tdr_ctx* ctx, float trans_x, float trans_y, float omega, int scale_freeze, float pos_cov,
                         float theta_cov, const void* dev_z, int64_t n) {
  CTX_CHECK(ctx);
  return propagate_host_z(ctx, trans_x, trans_y, omega, scale_freeze, pos_cov, theta_cov, reinterpret_cast<const float*>(dev_z), true, n);
}
int tdr_pf_propagate_rng(tdr_ctx* ctx, float trans_x, float trans_y, float omega, int scale_freeze, float pos_cov,
                         float theta_cov, uint64_t seed, uint64_t step, float* z_out) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur];
  float* dz = nullptr;
  if (z_out) { if (int e = ctx->scratch.reserve((size_t)pt.n * 16)) return e; dz = ctx->scratch.as<float>(); }
  if (int e = propagate(ctx, trans_x, trans_y, omega, scale_freeze, pos_cov, theta_cov, nullptr, true, seed, step, dz)) return e;
  if (z_out) {
    TDR_CUDA(cudaMemcpyAsync(z_out, dz, (size_t)pt.n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return TDR_OK;
}
int tdr_pf_gmm_samples(tdr_ctx* ctx, int num_samples, double* samples) { CTX_CHECK(ctx); return gmm_samples(ctx, num_samples, samples); }
int tdr_pf_get_last_dist(tdr_ctx* ctx, float* last_dist, int64_t n) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(last_dist && n == pt.n, TDR_EINVAL, "bad last_dist buffer");
  TDR_CUDA(cudaMemcpyAsync(last_dist, pt.last_dist.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

// the set the last resampling READ from (particles that were scored): the other half of the ping-pong pair.  What a
// lazily refreshed host mirror shows as new_particles_ after ParticleFilter::update's swap (particle_filter.cpp:187).
int tdr_pf_get_prev_states(tdr_ctx* ctx, tdr_state* states, float* last_dist, int64_t n) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur ^ 1];
  TDR_REQUIRE(states && n > 0 && n <= pt.n, TDR_EINVAL, "bad state count %lld (the previous set holds %lld)", (long long)n, (long long)pt.n);
  if (int e = ctx->scratch.reserve((size_t)n * 28)) return e;
  k_soa_to_aos<<<grid_for(ctx, n), 256, 0, ctx->stream>>>(soa_of(pt), n, ctx->scratch.as<uint32_t>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(states, ctx->scratch.p, (size_t)n * 28, cudaMemcpyDeviceToHost, ctx->stream));
  if (last_dist) TDR_CUDA(cudaMemcpyAsync(last_dist, pt.last_dist.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}
// raw weights (StateParticle::weight(), before particle_filter.cpp:107-147 rewrites the vector) survive the normalisation
// in a device-side copy when asked for: 4 B / particle device-to-device per update instead of a D2H inside it
int tdr_pf_keep_raw_weights(tdr_ctx* ctx, int on) { CTX_CHECK(ctx); ctx->keep_raw = on != 0; return TDR_OK; }
int tdr_pf_get_raw_weights(tdr_ctx* ctx, float* weights, int64_t n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(weights && n > 0 && n <= ctx->n_raw, TDR_EINVAL, "bad raw weight count %lld (kept: %lld)", (long long)n, (long long)ctx->n_raw);
  TDR_CUDA(cudaMemcpyAsync(weights, ctx->raw_weights.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}
static int snapshot_raw(tdr_ctx* ctx) {
  if (!ctx->keep_raw) return TDR_OK;
  if (int e = ctx->raw_weights.reserve((size_t)ctx->n_weights * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->raw_weights.p, ctx->weights.p, (size_t)ctx->n_weights * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->n_raw = ctx->n_weights;
  return TDR_OK;
}

static int copy_particles(tdr_ctx* ctx, Particles& dst, Particles& src) {
  if (int e = dst.reserve(src.n)) return e;
  DevBuf* d[] = {&dst.init_x, &dst.init_y, &dst.dx, &dst.dy, &dst.theta, &dst.scale, &dst.last_dist};
  DevBuf* s[] = {&src.init_x, &src.init_y, &src.dx, &src.dy, &src.theta, &src.scale, &src.last_dist};
  for (int k = 0; k < 7; k++)
    TDR_CUDA(cudaMemcpyAsync(d[k]->p, s[k]->p, (size_t)src.n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(dst.have_init.p, src.have_init.p, (size_t)src.n, cudaMemcpyDeviceToDevice, ctx->stream));
  dst.n = src.n;
  return TDR_OK;
}

int tdr_pf_checkpoint(tdr_ctx* ctx) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(pt.n > 0, TDR_ESTATE, "no particles");
  if (int e = copy_particles(ctx, ctx->ckpt, pt)) return e;
  if (int e = sync_uninit(ctx)) return e;
  ctx->ckpt_uninit = ctx->n_uninit;
  return TDR_OK;
}

int tdr_pf_restore(tdr_ctx* ctx) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->ckpt.n > 0, TDR_ESTATE, "no checkpoint");
  if (int e = copy_particles(ctx, ctx->part[ctx->cur], ctx->ckpt)) return e;
  ctx->n_uninit = ctx->ckpt_uninit; ctx->uninit_pending = false; ctx->have_argmax = false;
  return TDR_OK;
}

int tdr_pf_count(tdr_ctx* ctx, int64_t* n) { TDR_REQUIRE(ctx && n, TDR_EINVAL, "null argument"); *n = ctx->part[ctx->cur].n; return TDR_OK; }

static int copy_out_floats(tdr_ctx* ctx, float* dst, const void* src, int64_t n) {
  TDR_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  return TDR_OK;
}

int tdr_pf_score(tdr_ctx* ctx, float res, float* weights_out) {
  CTX_CHECK(ctx);
  if (int e = scan_pack(ctx)) return e;
  stage_mark(ctx, TDR_STAGE_SCORE);
  if (int e = score_particles(ctx, res)) return e;
  ctx->ld_override = nullptr;
  if (int e = snapshot_raw(ctx)) return e;
  if (weights_out) return copy_out_floats(ctx, weights_out, ctx->weights.p, ctx->n_weights);
  return TDR_OK;
}

int tdr_pf_set_weights(tdr_ctx* ctx, const float* weights, int64_t n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(weights && n > 0, TDR_EINVAL, "bad weights");
  if (int e = ctx->weights.reserve((size_t)n * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->weights.p, weights, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->n_weights = n;
  return TDR_OK;
}

int tdr_pf_set_weights_dev(tdr_ctx* ctx, const void* dev_weights, int64_t n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(dev_weights && n > 0, TDR_EINVAL, "bad weights");
  if (int e = ctx->weights.reserve((size_t)n * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->weights.p, dev_weights, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->n_weights = n;
  return TDR_OK;
}

int tdr_pf_get_weights(tdr_ctx* ctx, float* weights, int64_t n) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(weights && n > 0 && n <= ctx->n_weights, TDR_EINVAL, "bad weight count");
  return copy_out_floats(ctx, weights, ctx->weights.p, n);
}

int tdr_pf_normalize(tdr_ctx* ctx, int64_t* argmax_out, float* stats) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->ld_override || ctx->n_weights == ctx->part[ctx->cur].n, TDR_ESTATE,
              "weights (%lld) and particles (%lld) differ", (long long)ctx->n_weights, (long long)ctx->part[ctx->cur].n);
  if (int e = normalize(ctx)) return e;
  if (!ctx->ld_override) { if (int e = cache_ml_state(ctx, ctx->part[ctx->cur])) return e; }
  if (argmax_out || stats) {
    float host[16];
    TDR_CUDA(cudaMemcpyAsync(host, ctx->scal.p, 16 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
    int a; memcpy(&a, &host[SC_ARGMAX], 4);
    ctx->argmax = a;
    if (argmax_out) *argmax_out = a;
    if (stats) for (int k = 0; k < 6; k++) stats[k] = host[k];
  }
  return TDR_OK;
}

int tdr_pf_resample(tdr_ctx* ctx, float u, int64_t M, int32_t* idx_out) {
  CTX_CHECK(ctx);
  if (int e = resample(ctx, u, M, 0, M, &ctx->part[ctx->cur], &ctx->part[ctx->cur ^ 1])) return e;
  ctx->cur ^= 1;
  if (idx_out) {
    TDR_CUDA(cudaMemcpyAsync(idx_out, ctx->idx.p, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return TDR_OK;
}

int tdr_pf_pose(tdr_ctx* ctx, float mean[4], float cov_mean[16], float ml[4], float cov_ml[16]) {
  CTX_CHECK(ctx);
  return pose_of(ctx, ctx->part[ctx->cur], mean, cov_mean, ml, cov_ml);
}

// a11 + a12 on the resident raw weights: one fused single-CTA kernel for small sets, the tiled kernels otherwise
static int normalize_resample(tdr_ctx* ctx, float u, int64_t M) {
  bool used = false;
  if (int e = small_update(ctx, u, M, true, &used)) return e;
  if (used) {
    if (int e = cache_ml_state(ctx, ctx->part[ctx->cur])) return e;
    stage_mark(ctx, TDR_STAGE_RESAMPLE);
    ctx->cur ^= 1;
    return TDR_OK;
  }
  if (int e = normalize(ctx, true)) return e;
  if (int e = cache_ml_state(ctx, ctx->part[ctx->cur])) return e;
  stage_mark(ctx, TDR_STAGE_RESAMPLE);
  if (int e = resample(ctx, u, M, 0, M, &ctx->part[ctx->cur], &ctx->part[ctx->cur ^ 1])) return e;
  ctx->cur ^= 1;
  return TDR_OK;
}

int tdr_pf_normalize_resample(tdr_ctx* ctx, float u, int64_t M, int64_t* argmax_out, int32_t* idx_out) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(ctx->n_weights == ctx->part[ctx->cur].n, TDR_ESTATE, "weights (%lld) and particles (%lld) differ",
              (long long)ctx->n_weights, (long long)ctx->part[ctx->cur].n);
  ctx->ld_override = nullptr;
  if (int e = normalize_resample(ctx, u, M)) return e;
  if (argmax_out) {
    int a = 0;
    TDR_CUDA(cudaMemcpyAsync(&a, ctx->scal.as<float>() + SC_ARGMAX, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
    *argmax_out = a;
  }
  if (idx_out) {
    TDR_CUDA(cudaMemcpyAsync(idx_out, ctx->idx.p, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return TDR_OK;
}

static int update_resident(tdr_ctx* ctx, float res, float u, int64_t M) {
  if (int e = scan_pack(ctx)) return e;
  stage_mark(ctx, TDR_STAGE_SCORE);
  if (int e = score_particles(ctx, res)) return e;
  ctx->ld_override = nullptr;
  if (int e = snapshot_raw(ctx)) return e;
  stage_mark(ctx, TDR_STAGE_NORMALIZE);
  if (int e = normalize_resample(ctx, u, M)) return e;
  stage_mark(ctx, TDR_N_STAGES);
  ctx->stage_valid = ctx->profiling;
  return TDR_OK;
}

int tdr_pf_update(tdr_ctx* ctx, float res, float u, int64_t M) {
  CTX_CHECK(ctx);
  stage_mark(ctx, TDR_STAGE_RENDER);   // empty render stage
  return update_resident(ctx, res, u, M);
}

int tdr_step(tdr_ctx* ctx, float res, float ang_res, int n_theta, int n_r, float u, int64_t M) {
  CTX_CHECK(ctx);
  stage_mark(ctx, TDR_STAGE_RENDER);
  if (int e = sync_uninit(ctx)) return e;
  if (int e = score_i8_prepare_async(ctx)) return e;      // the particle sort of a theta search starts beside the rasteriser
  if (int e = render_polar_resident(ctx, res, ang_res, n_theta, n_r)) return e;
  return update_resident(ctx, res, u, M);
}

// ---- multi-GPU (particle shards; the collective itself is issued by the host layer) ---------------------
int tdr_pf_export_shard(tdr_ctx* ctx, void* dev_out, int64_t capacity_floats, int with_weights) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(pt.n > 0, TDR_ESTATE, "no particles");
  TDR_REQUIRE(dev_out && capacity_floats >= pt.n * TDR_SHARD_ROWS, TDR_EINVAL, "shard buffer too small");
  TDR_REQUIRE(!with_weights || ctx->n_weights == pt.n, TDR_ESTATE, "weights (%lld) and particles (%lld) differ",
              (long long)ctx->n_weights, (long long)pt.n);
  ShardSrc s; s.w = with_weights ? ctx->weights.as<float>() : nullptr;
  s.ix = pt.init_x.as<float>(); s.iy = pt.init_y.as<float>(); s.dx = pt.dx.as<float>(); s.dy = pt.dy.as<float>();
  s.th = pt.theta.as<float>(); s.sc = pt.scale.as<float>(); s.ld = pt.last_dist.as<float>(); s.hi = pt.have_init.as<uint8_t>();
  k_pack_shard<<<grid_for(ctx, pt.n), 256, 0, ctx->stream>>>(s, pt.n, reinterpret_cast<float*>(dev_out));
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}

static int unpack_all(tdr_ctx* ctx, const void* dev_all, int n_ranks, int64_t n_local, bool weights) {
  const int64_t N = (int64_t)n_ranks * n_local;
  TDR_REQUIRE(dev_all && n_ranks >= 1 && n_local > 0 && N < (1ll << 31), TDR_EINVAL, "bad gathered block");
  if (int e = ctx->all.reserve(N)) return e;
  if (weights) { if (int e = ctx->weights.reserve((size_t)N * 4)) return e; }
  Particles& a = ctx->all;
  ShardDst d; d.w = weights ? ctx->weights.as<float>() : nullptr;
  d.ix = a.init_x.as<float>(); d.iy = a.init_y.as<float>(); d.dx = a.dx.as<float>(); d.dy = a.dy.as<float>();
  d.th = a.theta.as<float>(); d.sc = a.scale.as<float>(); d.ld = a.last_dist.as<float>(); d.hi = a.have_init.as<uint8_t>();
  k_unpack_all<<<grid_for(ctx, N), 256, 0, ctx->stream>>>(reinterpret_cast<const float*>(dev_all), n_ranks, n_local, d);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  a.n = N;
  if (weights) ctx->n_weights = N;
  return TDR_OK;
}

int tdr_pf_update_gathered(tdr_ctx* ctx, const void* dev_all, int n_ranks, int64_t n_local, float u, int64_t M,
                           int64_t i0, int64_t i1) {
  CTX_CHECK(ctx);
  stage_mark(ctx, TDR_STAGE_NORMALIZE);
  if (int e = unpack_all(ctx, dev_all, n_ranks, n_local, true)) return e;
  ctx->ld_override = ctx->all.last_dist.as<float>();
  int e = normalize(ctx, true);
  ctx->ld_override = nullptr;
  if (e) return e;
  if (int e2 = cache_ml_state(ctx, ctx->all)) return e2;
  stage_mark(ctx, TDR_STAGE_RESAMPLE);
  if (int e2 = resample(ctx, u, M, i0, i1, &ctx->all, &ctx->part[ctx->cur ^ 1])) return e2;
  ctx->cur ^= 1;
  stage_mark(ctx, TDR_N_STAGES);
  ctx->stage_valid = ctx->profiling;
  return TDR_OK;
}

// ---- the same update as two collectives: weights + last_dist first (8 B/particle), the states (28 B/particle) while
// the normalisation of the N weights runs
int tdr_pf_export_split(tdr_ctx* ctx, void* dev_wl, void* dev_states) {
  CTX_CHECK(ctx);
  Particles& pt = ctx->part[ctx->cur];
  TDR_REQUIRE(pt.n > 0 && dev_wl && dev_states, TDR_EINVAL, "no particles / null buffers");
  TDR_REQUIRE(ctx->n_weights == pt.n, TDR_ESTATE, "weights (%lld) and particles (%lld) differ", (long long)ctx->n_weights, (long long)pt.n);
  ShardSrc s; s.w = ctx->weights.as<float>();
  s.ix = pt.init_x.as<float>(); s.iy = pt.init_y.as<float>(); s.dx = pt.dx.as<float>(); s.dy = pt.dy.as<float>();
  s.th = pt.theta.as<float>(); s.sc = pt.scale.as<float>(); s.ld = pt.last_dist.as<float>(); s.hi = pt.have_init.as<uint8_t>();
  k_pack_split<<<grid_for(ctx, pt.n), 256, 0, ctx->stream>>>(s, pt.n, reinterpret_cast<float*>(dev_wl), reinterpret_cast<float*>(dev_states));
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  return TDR_OK;
}
int tdr_pf_set_shard_count(tdr_ctx* ctx, int n_ranks) {
  TDR_REQUIRE(ctx && n_ranks >= 1 && n_ranks <= TDR_MAX_PEERS, TDR_EINVAL, "bad rank count");
  ctx->count_scale = n_ranks;
  return TDR_OK;
}
int tdr_pf_normalize_gathered(tdr_ctx* ctx, const void* dev_wl_all, int n_ranks, int64_t n_local) {
  CTX_CHECK(ctx);
  const int64_t N = (int64_t)n_ranks * n_local;
  TDR_REQUIRE(dev_wl_all && n_ranks >= 1 && n_local > 0 && N < (1ll << 31), TDR_EINVAL, "bad gathered block");
  stage_mark(ctx, TDR_STAGE_NORMALIZE);
  if (int e = ctx->all.reserve(N)) return e;
  if (int e = ctx->weights.reserve((size_t)N * 4)) return e;
  k_unpack_wl<<<grid_for(ctx, N), 256, 0, ctx->stream>>>(reinterpret_cast<const float*>(dev_wl_all), n_ranks, n_local,
                                                        ctx->weights.as<float>(), ctx->all.last_dist.as<float>());
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  ctx->all.n = N; ctx->n_weights = N;
  ctx->ld_override = ctx->all.last_dist.as<float>();
  const int e = normalize(ctx, true);
  ctx->ld_override = nullptr;
  return e;
}
int tdr_pf_resample_gathered(tdr_ctx* ctx, const void* dev_states_all, int n_ranks, int64_t n_local, float u, int64_t M,
                             int64_t i0, int64_t i1) {
  CTX_CHECK(ctx);
  const int64_t N = (int64_t)n_ranks * n_local;
  TDR_REQUIRE(dev_states_all && N == ctx->all.n && N == ctx->n_weights, TDR_ESTATE, "tdr_pf_normalize_gathered has not run on this set");
  Particles& a = ctx->all;
  ShardDst d; d.w = nullptr;
  d.ix = a.init_x.as<float>(); d.iy = a.init_y.as<float>(); d.dx = a.dx.as<float>(); d.dy = a.dy.as<float>();
  d.th = a.theta.as<float>(); d.sc = a.scale.as<float>(); d.ld = a.last_dist.as<float>(); d.hi = a.have_init.as<uint8_t>();
  k_unpack_states<<<grid_for(ctx, N), 256, 0, ctx->stream>>>(reinterpret_cast<const float*>(dev_states_all), n_ranks, n_local, d);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  if (int e = cache_ml_state(ctx, ctx->all)) return e;
  stage_mark(ctx, TDR_STAGE_RESAMPLE);
  if (int e = resample(ctx, u, M, i0, i1, &ctx->all, &ctx->part[ctx->cur ^ 1])) return e;
  ctx->cur ^= 1;
  stage_mark(ctx, TDR_N_STAGES);
  ctx->stage_valid = ctx->profiling;
  return TDR_OK;
}

int tdr_pf_pose_gathered(tdr_ctx* ctx, const void* dev_all, int n_ranks, int64_t n_local, float mean[4],
                         float cov_mean[16], float ml[4], float cov_ml[16]) {
  CTX_CHECK(ctx);
  if (int e = unpack_all(ctx, dev_all, n_ranks, n_local, false)) return e;
  return pose_of(ctx, ctx->all, mean, cov_mean, ml, cov_ml);
}

// ---- exhaustive grid --------------------------------------------------------------------------------
int tdr_grid_costs(tdr_ctx* ctx, const float* centers_xy, int64_t n, float scale, float res, const int32_t* shifts,
                   int n_shifts, float* costs_out) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(n > 0 && shifts && n_shifts > 0 && n_shifts <= TDR_MAX_SHIFTS, TDR_EINVAL, "bad grid arguments");
  TDR_REQUIRE(centers_xy || ctx->grid_n == n, TDR_ESTATE, "no resident centres for a grid of %lld", (long long)n);
  if (centers_xy) {
    if (int e = ctx->grid_centers.reserve((size_t)n * 8)) return e;
    TDR_CUDA(cudaMemcpyAsync(ctx->grid_centers.p, centers_xy, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->perm_grid_n = -1;
    // a lattice with an x stride of 2, 4 or 8 px gathers from the phase-split map copy (a layout hint only: any
    // value gives the same results); judged on the first pairs of centres
    int votes[4] = {0, 0, 0, 0};
    const int64_t probe = n - 1 < 256 ? n - 1 : 256;
    for (int64_t i = 0; i < probe; i++) {
      const float d = (centers_xy[2 * i + 2] - centers_xy[2 * i]) / ctx->resolution;
      if (centers_xy[2 * i + 3] != centers_xy[2 * i + 1]) continue;
      for (int k = 1; k <= 3; k++) if (d == (float)(1 << k)) votes[k]++;
    }
    ctx->grid_phase_log2 = 0;
    for (int k = 1; k <= 3; k++) if (votes[k] * 2 > probe) ctx->grid_phase_log2 = k;
    if (const char* e = getenv("TDR_GRID_PHASE_LOG2")) { int v = atoi(e); if (v >= 0 && v <= 3) ctx->grid_phase_log2 = v; }
  }
  if (ctx->grid_n_peers) { /* costs go straight into the peer-mapped full arrays */ }
  else if (!ctx->grid_costs_ext) { if (int e = ctx->grid_costs.reserve((size_t)n * n_shifts * 4)) return e; }
  else TDR_REQUIRE(ctx->grid_costs_ext_cap >= n * n_shifts, TDR_EINVAL, "external cost buffer too small");
  if (int e = ctx->grid_shifts.reserve((size_t)n_shifts * 4)) return e;
  TDR_CUDA(cudaMemcpyAsync(ctx->grid_shifts.p, shifts, (size_t)n_shifts * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->grid_n = n; ctx->grid_shifts_n = n_shifts; ctx->grid_shifts_host.assign(shifts, shifts + n_shifts);
  if (int e = scan_pack(ctx)) return e;
  stage_mark(ctx, TDR_STAGE_SCORE);
  if (int e = score_grid(ctx, n, scale, res)) return e;
  stage_mark(ctx, TDR_STAGE_NORMALIZE); stage_mark(ctx, TDR_STAGE_RESAMPLE); stage_mark(ctx, TDR_N_STAGES);
  ctx->stage_valid = ctx->profiling;
  if (centers_xy || costs_out) {      // caller-owned pageable buffers: finish before returning
    TDR_REQUIRE(!(costs_out && ctx->grid_n_peers), TDR_ESTATE, "with peer buffers registered the costs live in the full arrays");
    if (costs_out) TDR_CUDA(cudaMemcpyAsync(costs_out, grid_costs_ptr(ctx), (size_t)n * n_shifts * 4, cudaMemcpyDeviceToHost, ctx->stream));
    TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return TDR_OK;
}

int tdr_grid_set_costs_buffer(tdr_ctx* ctx, void* dev_costs, int64_t capacity_floats) {
  TDR_REQUIRE(ctx, TDR_EINVAL, "null context");
  ctx->grid_costs_ext = reinterpret_cast<float*>(dev_costs);
  ctx->grid_costs_ext_cap = dev_costs ? capacity_floats : 0;
  return TDR_OK;
}

int tdr_grid_best(tdr_ctx* ctx, float* best_cost, int64_t* best_index) {
  CTX_CHECK(ctx);
  long long idx = -1;
  int e = grid_best(ctx, grid_costs_ptr(ctx), ctx->grid_n * ctx->grid_shifts_n, best_cost, &idx);
  if (best_index) *best_index = idx;
  return e;
}

int tdr_grid_best_dev(tdr_ctx* ctx, const void* dev_costs, int64_t n, float* best_cost, int64_t* best_index) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(dev_costs && n > 0, TDR_EINVAL, "bad arguments");
  long long idx = -1;
  int e = grid_best(ctx, reinterpret_cast<const float*>(dev_costs), n, best_cost, &idx);
  if (best_index) *best_index = idx;
  return e;
}

int tdr_grid_best_key(tdr_ctx* ctx, uint64_t* key) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(key, TDR_EINVAL, "null argument");
  TDR_REQUIRE(ctx->grid_key_valid, TDR_ESTATE, "the last grid did not run on the tensor-core kernel: use tdr_grid_best");
  int bailed = 0;
  TDR_CUDA(cudaMemcpyAsync(key, ctx->grid_key.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaMemcpyAsync(&bailed, ctx->scal.as<float>() + SC_MMA_BAILED, 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  // the fp16 exactness of the scan counts is checked on the device; this read-back is where the host learns about it
  TDR_REQUIRE(!bailed, TDR_ESTATE, "scan counts above 2048: the tensor-core kernel left this grid to the CUDA cores%s",
              ctx->grid_n_peers ? " and the fused peer all-gather did not run" : " (use tdr_grid_best)");
  return TDR_OK;
}
int tdr_grid_key_decode(uint64_t key, float* best_cost, int64_t* best_index) {
  if (key == ~0ull) { if (best_cost) *best_cost = NAN; if (best_index) *best_index = -1; return TDR_OK; }
  uint32_t u = (uint32_t)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float v; memcpy(&v, &u, 4);
  if (best_cost) *best_cost = v;
  if (best_index) *best_index = (int64_t)(key & 0xffffffffull);
  return TDR_OK;
}

// ---- the cross-rank step of the fused grid: arg-min reduction + barrier over peer memory, no NCCL -------------------
// Behind every rank's full cost array sits a mailbox its peers have mapped with the array: two slots of
// (key: u64, arrivals: u32).  Exchange e uses slot e & 1.  Every rank atomically MINs its own (cost, flat index) key into
// the slot of EVERY rank (its own included) and then bumps that rank's arrival counter, both with system-scope atomics
// over NVLink; it then spins until its own counter shows all ranks.  At that point its slot holds the global minimum and —
// because each rank publishes from a kernel launched behind its score kernel, after a system-scope fence — every peer's
// cost stores into this rank's array have landed.  Counters never reset (target = ranks x uses of the slot); a key slot
// is reset by its owner right after reading it, which is safe because no peer can start exchange e + 2 before this rank
// has arrived in exchange e + 1.
static const size_t GRID_MAILBOX_BYTES = 256;
static size_t mailbox_offset(int64_t n_floats) { return (((size_t)n_floats * 4) + 255) / 256 * 256; }
struct GridBoxes { unsigned char* box[TDR_MAX_PEERS]; };
__global__ void k_grid_exchange(GridBoxes peers, int n_peers, unsigned char* own, int slot, unsigned int target,
                                unsigned long long* local_key, unsigned long long* pin, int* timed_out) {
  const int d = threadIdx.x;
  const unsigned long long mine = *local_key;
  if (d < n_peers) {
    unsigned long long* key = reinterpret_cast<unsigned long long*>(peers.box[d]) + slot;
    unsigned int* arrive = reinterpret_cast<unsigned int*>(peers.box[d] + 16) + slot;
    atomicMin_system(key, mine);
    __threadfence_system();
    atomicAdd_system(arrive, 1u);
  }
  __syncwarp();
  if (d == 0) {
    volatile unsigned int* arrive = reinterpret_cast<volatile unsigned int*>(own + 16) + slot;
    const long long t0 = clock64();
    bool ok = true;
    while (*arrive < target) {
      if (clock64() - t0 > 6000000000ll) { ok = false; break; }       // ~3 s: a peer never arrived
      __nanosleep(200);
    }
    __threadfence_system();
    unsigned long long* key = reinterpret_cast<unsigned long long*>(own) + slot;
    const unsigned long long best = ok ? *reinterpret_cast<volatile unsigned long long*>(key) : ~0ull;
    *key = ~0ull;                                                     // ready for exchange e + 2
    __threadfence_system();
    *local_key = best;
    *pin = best;
    *timed_out = ok ? 0 : 1;
  }
}

int tdr_grid_peer_exchange(tdr_ctx* ctx, uint64_t* key) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(key, TDR_EINVAL, "null argument");
  TDR_REQUIRE(ctx->grid_n_peers >= 1 && ctx->grid_full.p && ctx->grid_full_floats > 0, TDR_ESTATE, "no peer arrays registered");
  TDR_REQUIRE(ctx->grid_key_valid, TDR_ESTATE, "the last grid did not run on the tensor-core kernel");
  const size_t box = mailbox_offset(ctx->grid_full_floats);
  GridBoxes peers;
  for (int d = 0; d < TDR_MAX_PEERS; d++)
    peers.box[d] = d < ctx->grid_n_peers ? reinterpret_cast<unsigned char*>(ctx->grid_peers[d]) + box : nullptr;
  const int slot = (int)(ctx->grid_epoch & 1);
  const unsigned int target = (unsigned int)(ctx->grid_n_peers * (ctx->grid_epoch / 2 + 1));
  ctx->grid_epoch++;
  int* flags = reinterpret_cast<int*>(ctx->grid_key_pin + 1);
  k_grid_exchange<<<1, 32, 0, ctx->stream>>>(peers, ctx->grid_n_peers, ctx->grid_full.as<unsigned char>() + box, slot, target,
                                             ctx->grid_key.as<unsigned long long>(), ctx->grid_key_pin, flags);
  count_launch(ctx);
  TDR_CUDA(cudaGetLastError());
  TDR_CUDA(cudaMemcpyAsync(flags + 1, ctx->scal.as<float>() + SC_MMA_BAILED, 4, cudaMemcpyDeviceToHost, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  TDR_REQUIRE(!flags[0], TDR_ESTATE, "grid exchange: a peer did not arrive within 3 s");
  TDR_REQUIRE(!flags[1], TDR_ESTATE, "scan counts above 2048: the fused peer all-gather did not run");
  *key = *ctx->grid_key_pin;
  return TDR_OK;
}

int tdr_grid_peer_alloc(tdr_ctx* ctx, int64_t n_floats, void** dev_ptr, uint8_t handle[TDR_IPC_HANDLE_BYTES]) {
  CTX_CHECK(ctx);
  static_assert(sizeof(cudaIpcMemHandle_t) == TDR_IPC_HANDLE_BYTES, "IPC handle size");
  TDR_REQUIRE(n_floats > 0 && dev_ptr && handle, TDR_EINVAL, "bad arguments");
  const size_t box = mailbox_offset(n_floats);
  if (int e = ctx->grid_full.reserve(box + GRID_MAILBOX_BYTES)) return e;
  ctx->grid_full_floats = n_floats; ctx->grid_epoch = 0;
  // mailbox: two (key, arrivals) slots; keys start at "no result", arrival counters at 0 and only ever grow
  unsigned long long init[GRID_MAILBOX_BYTES / 8];
  memset(init, 0, sizeof(init));
  init[0] = init[1] = ~0ull;
  TDR_CUDA(cudaMemcpyAsync(ctx->grid_full.as<unsigned char>() + box, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  TDR_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  TDR_CUDA(cudaIpcGetMemHandle(&h, ctx->grid_full.p));
  memcpy(handle, &h, TDR_IPC_HANDLE_BYTES);
  *dev_ptr = ctx->grid_full.p;
  return TDR_OK;
}
int tdr_grid_peer_open(tdr_ctx* ctx, const uint8_t handle[TDR_IPC_HANDLE_BYTES], void** dev_ptr) {
  CTX_CHECK(ctx);
  TDR_REQUIRE(handle && dev_ptr, TDR_EINVAL, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, TDR_IPC_HANDLE_BYTES);
  void* p = nullptr;
  TDR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->grid_opened.push_back(p);
  *dev_ptr = p;
  return TDR_OK;
}
int tdr_grid_peer_set(tdr_ctx* ctx, void* const* peer_ptrs, int n_peers, int64_t row_offset) {
  TDR_REQUIRE(ctx && peer_ptrs && n_peers >= 1 && n_peers <= TDR_MAX_PEERS && row_offset >= 0, TDR_EINVAL, "bad peer registration");
  for (int k = 0; k < n_peers; k++) { TDR_REQUIRE(peer_ptrs[k], TDR_EINVAL, "null peer pointer %d", k); ctx->grid_peers[k] = reinterpret_cast<float*>(peer_ptrs[k]); }
  ctx->grid_n_peers = n_peers; ctx->grid_peer_row0 = row_offset;
  return TDR_OK;
}
int tdr_grid_peer_clear(tdr_ctx* ctx) {
  CTX_CHECK(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->grid_opened) cudaIpcCloseMemHandle(p);
  ctx->grid_opened.clear(); ctx->grid_n_peers = 0; ctx->grid_epoch = 0;
  return TDR_OK;
}

int tdr_dev_ptr(tdr_ctx* ctx, int which, void** ptr, int64_t* n_elems) {
  TDR_REQUIRE(ctx && ptr, TDR_EINVAL, "null argument");
  switch (which) {
    case TDR_BUF_WEIGHTS: *ptr = ctx->weights.p; if (n_elems) *n_elems = ctx->n_weights; break;
    case TDR_BUF_GRID_COSTS: *ptr = grid_costs_ptr(ctx); if (n_elems) *n_elems = ctx->grid_n * ctx->grid_shifts_n; break;
    case TDR_BUF_GRID_BEST_KEY: if (int e = ctx->grid_key.reserve(8)) return e; *ptr = ctx->grid_key.p; if (n_elems) *n_elems = 1; break;
    case TDR_BUF_SCAN_IMAGES: *ptr = ctx->scan_img.p; if (n_elems) *n_elems = (int64_t)ctx->scan_C * ctx->scan_theta * ctx->scan_r; break;
    default: tdr::set_error("unknown buffer id %d", which); return TDR_EINVAL;
  }
  return TDR_OK;
}

}  // extern "C"
