// tdr_cpu_standin.cpp — TEST INFRASTRUCTURE ONLY: the subset of the C ABI (include/tdr.h) that the adapters
// (top_down_renderer_b200/adapters/*.cpp) call, answered by the CPU oracle (oracle/tdr_oracle.cpp).
//
// It exists so that the HOST logic of the adapter classes — the RNG streams of initializeParticles / propagate /
// update, the lazy host mirror, the map-centre shift of updateMap, freezeScale, the map cache files — can be exercised
// by the `-m "not gpu"` suite (tests/test_adapters.py) and so that the checks of the GPU tests are themselves proven on
// a run whose every number comes from the oracle.  It is linked into the adapters' CPU test library only (oracle/Makefile,
// `_adapters`); the package, the product library libtdr_b200.so, bench.py and smoke() never see it, and
// libtdr_b200.so keeps failing with TDR_ENOGPU when no device is usable.
#include <math.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tdr.h"

extern "C" {
struct OrcState { float init_x_px, init_y_px, dx_m, dy_m, theta, scale; uint8_t have_init; uint8_t pad[3]; };
struct OrcFilterParams {
  float regularization; int force_on_map; float fixed_scale, scale_log_min, scale_log_max;
  float map_width, map_height; int num_classes; float class_weights[16];
};
void orc_render_polar(const uint8_t* pts, int stride, int intensity_off, long n, float res, float ang_res, int n_theta, int n_r,
                      const int* lut, int n_lut, int C, float* imgs);
void orc_render_geometric_polar(const uint8_t* pts, int stride, int width, int height, float res, float ang_res, int n_theta, int n_r, float* imgs);
void orc_render_geometric_cart(const uint8_t* pts, int stride, int width, int height, float res, int rows, int cols, float* imgs);
void orc_render_cart(const uint8_t* pts, int stride, int intensity_off, long n, float res, int rows, int cols, const int* lut, int n_lut,
                     int C, float* imgs);
void orc_local_map_cart(const float* layers, const uint8_t* mask, int rows, int cols, int C, float resolution, float cx, float cy,
                        float rot, float res, int out_rows, int out_cols, float* dists, uint8_t* mask_out);
void orc_map_dims(int h_img, int w_img, float res, int* rows, int* cols);
void orc_class_image_to_layers(const uint8_t* img, int h_img, int w_img, int stride, const int* lut, int n_lut, int C, float res,
                               float* layers);
void orc_compute_dists(float* layers, int rows, int cols, int C, float resolution, uint8_t* mask);
void orc_geo_raster(const float* class_layers, int rows, int cols, int C, float* geo);
void orc_local_map_polar(const float* layers, const uint8_t* mask, int rows, int cols, int C, float resolution, const float* tab,
                         int P, float cx, float cy, float scale, float res, float* dists, uint8_t* mask_out);
void orc_score_all(OrcState* states, long n, const OrcFilterParams* fp, const float* layers, const uint8_t* mask,
                   const float* geo_layers, int rows, int cols, float resolution, const float* tab, int n_theta, int n_r,
                   const float* scan, float res, const float* search_thetas, const int* search_shifts, int n_search, float* weights,
                   int n_threads);
long orc_normalize(float* w, const float* last_dist, long n, float* stats);
void orc_resample_fast(const float* w, long n, float shift, int M, int* idx, float* prefix_out);
void orc_mean_cov(const OrcState* st, long n, float mean[4], float cov[16]);
void orc_ml_cov(const OrcState* st, long n, long argmax, float ml[4], float cov[16]);
void orc_active_best_rel_pos(const float* layers, const uint8_t* mask, int rows, int cols, int C, float resolution, const float* tab,
                             int n_theta, int n_r, const float* preds, int n, float rel[2], float* best_diff_out);
void orc_gmm_samples(const OrcState* st, long n, int num_samples, double* samples);
void orc_raster_polygons(const float* verts, const int* poly_start, const int* poly_class, int n_poly, int map_w, int map_h, float rot,
                         float resolution, int C, const int* exclusive, int n_excl, float* layers);
}
static_assert(sizeof(OrcState) == sizeof(tdr_state), "State layouts differ");

struct tdr_ctx {
  int rows = 0, cols = 0, C = 0, n_theta = 0, n_r = 0, scan_C = 0, n_lut = 0;
  float resolution = 1;
  std::vector<float> layers, seeds, geo, tab, scan, weights, last_dist, thetas;
  std::vector<uint8_t> mask, pts;
  std::vector<int> lut, shifts;
  int pts_stride = 0, pts_off = 0; long n_pts = 0;
  std::vector<OrcState> states, prev_states;
  std::vector<float> prev_last_dist, raw;
  bool keep_raw = false;
  OrcState ml_state{};
  bool have_ml = false, have_geo = false;
  tdr_filter_params fp{};
};
static thread_local std::string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }
#define REQ(c, code, msg) do { if (!(c)) return fail(code, msg); } while (0)

extern "C" {
int tdr_abi_version(void) { return TDR_ABI_VERSION; }
const char* tdr_last_error(void) { return g_err.c_str(); }
int tdr_create(tdr_ctx** out, int) { *out = new tdr_ctx(); return TDR_OK; }
void tdr_destroy(tdr_ctx* c) { delete c; }
int tdr_sync(tdr_ctx*) { return TDR_OK; }

static void ensure_geo(tdr_ctx* c) {
  if (c->have_geo) return;
  c->geo.resize((size_t)2 * c->rows * c->cols);
  orc_geo_raster(c->seeds.data(), c->rows, c->cols, c->C, c->geo.data());
  std::vector<uint8_t> tmp((size_t)c->rows * c->cols);
  orc_compute_dists(c->geo.data(), c->rows, c->cols, 2, c->resolution, tmp.data());
  c->have_geo = true;
}
int tdr_map_set_class_image(tdr_ctx* c, const uint8_t* img, int h, int w, int stride, const int32_t* lut, int n_lut, int C, float res) {
  REQ(img && lut && C > 0 && C <= TDR_MAX_CLASSES, TDR_EINVAL, "bad map arguments");
  orc_map_dims(h, w, res, &c->rows, &c->cols);
  c->C = C; c->resolution = res;
  c->seeds.assign((size_t)C * c->rows * c->cols, 0.f);
  orc_class_image_to_layers(img, h, w, stride, lut, n_lut, C, res, c->seeds.data());
  c->layers = c->seeds;
  c->mask.assign((size_t)c->rows * c->cols, 0);
  orc_compute_dists(c->layers.data(), c->rows, c->cols, C, res, c->mask.data());
  c->have_geo = false;
  return TDR_OK;
}
static int dists_from_seeds(tdr_ctx* c) {
  c->layers = c->seeds;
  c->mask.assign((size_t)c->rows * c->cols, 0);
  orc_compute_dists(c->layers.data(), c->rows, c->cols, c->C, c->resolution, c->mask.data());
  c->have_geo = false;
  return TDR_OK;
}
int tdr_map_set_binary_layers(tdr_ctx* c, const float* layers, int rows, int cols, int C, float res) {
  REQ(layers && rows > 0 && cols > 0 && C > 0 && C <= TDR_MAX_CLASSES, TDR_EINVAL, "bad layers");
  c->rows = rows; c->cols = cols; c->C = C; c->resolution = res;
  c->seeds.assign(layers, layers + (size_t)C * rows * cols);
  return dists_from_seeds(c);
}
int tdr_map_set_polygons(tdr_ctx* c, const float* verts, const int32_t* start, const int32_t* cls, int n_poly, int map_w, int map_h,
                         float rot, int C, float res, const int32_t* excl, int n_excl, float* layers_out) {
  REQ(verts && start && cls && C > 0 && C <= TDR_MAX_CLASSES, TDR_EINVAL, "bad polygons");
  c->rows = (int)((float)map_h / res); c->cols = (int)((float)map_w / res); c->C = C; c->resolution = res;
  c->seeds.assign((size_t)C * c->rows * c->cols, 0.f);
  orc_raster_polygons(verts, start, cls, n_poly, map_w, map_h, rot, res, C, excl, n_excl, c->seeds.data());
  if (layers_out) std::copy(c->seeds.begin(), c->seeds.end(), layers_out);
  return dists_from_seeds(c);
}
int tdr_map_set_dist_layers(tdr_ctx* c, const float* layers, const uint8_t* mask, int rows, int cols, int C, float res) {
  REQ(layers && mask && rows > 0 && cols > 0, TDR_EINVAL, "bad layers");
  c->rows = rows; c->cols = cols; c->C = C; c->resolution = res;
  c->layers.assign(layers, layers + (size_t)C * rows * cols);
  c->mask.assign(mask, mask + (size_t)rows * cols);
  c->have_geo = false; c->seeds.clear();
  return TDR_OK;
}
int tdr_map_set_geo_dist_layers(tdr_ctx* c, const float* geo) {
  REQ(c->rows > 0 && geo, TDR_ESTATE, "no map");
  c->geo.assign(geo, geo + (size_t)2 * c->rows * c->cols);
  c->have_geo = true;
  return TDR_OK;
}
int tdr_map_get_layers(tdr_ctx* c, float* layers, uint8_t* mask) {
  REQ(c->rows > 0, TDR_ESTATE, "no map");
  if (layers) std::copy(c->layers.begin(), c->layers.end(), layers);
  if (mask) std::copy(c->mask.begin(), c->mask.end(), mask);
  return TDR_OK;
}
int tdr_map_get_geo_layers(tdr_ctx* c, float* geo) {
  REQ(c->rows > 0 && (c->have_geo || !c->seeds.empty()), TDR_ESTATE, "no map seeds");
  ensure_geo(c);
  std::copy(c->geo.begin(), c->geo.end(), geo);
  return TDR_OK;
}
int tdr_map_info(tdr_ctx* c, int* rows, int* cols, int* C, float* res) {
  if (rows) *rows = c->rows;
  if (cols) *cols = c->cols;
  if (C) *C = c->C;
  if (res) *res = c->resolution;
  return TDR_OK;
}
int tdr_map_set_polar_table(tdr_ctx* c, const float* tab, int n_theta, int n_r) {
  c->tab.assign(tab, tab + (size_t)2 * n_theta * n_r); c->n_theta = n_theta; c->n_r = n_r;
  return TDR_OK;
}
int tdr_map_local_polar(tdr_ctx* c, const float* xy, int n, float scale, float res, float* dists, uint8_t* mask) {
  REQ(c->rows > 0 && !c->tab.empty(), TDR_ESTATE, "no map / table");
  const int P = c->n_theta * c->n_r;
  for (int i = 0; i < n; i++)
    orc_local_map_polar(c->layers.data(), c->mask.data(), c->rows, c->cols, c->C, c->resolution, c->tab.data(), P, xy[2 * i],
                        xy[2 * i + 1], scale, res, dists + (size_t)i * c->C * P, mask + (size_t)i * P);
  return TDR_OK;
}
int tdr_map_local_cart(tdr_ctx* c, float cx, float cy, float rot, float res, int rows, int cols, float* dists, uint8_t* mask) {
  REQ(c->rows > 0, TDR_ESTATE, "no map");
  orc_local_map_cart(c->layers.data(), c->mask.data(), c->rows, c->cols, c->C, c->resolution, cx, cy, rot, res, rows, cols, dists, mask);
  return TDR_OK;
}
int tdr_map_local_geo_polar(tdr_ctx* c, const float* xy, int n, float scale, float res, float* geo) {
  REQ(c->rows > 0 && !c->tab.empty(), TDR_ESTATE, "no map / table");
  ensure_geo(c);
  const int P = c->n_theta * c->n_r;
  std::vector<uint8_t> zero((size_t)c->rows * c->cols, 0), m(P);
  for (int i = 0; i < n; i++)
    orc_local_map_polar(c->geo.data(), zero.data(), c->rows, c->cols, 2, c->resolution, c->tab.data(), P, xy[2 * i], xy[2 * i + 1],
                        scale, res, geo + (size_t)i * 2 * P, m.data());
  return TDR_OK;
}
int tdr_active_best_rel_pos(tdr_ctx* c, const float* preds, int n, float rel[2], float* diff) {
  REQ(c->rows > 0 && !c->tab.empty(), TDR_ESTATE, "no map / table");
  float d = 0;
  orc_active_best_rel_pos(c->layers.data(), c->mask.data(), c->rows, c->cols, c->C, c->resolution, c->tab.data(), c->n_theta, c->n_r,
                          preds, n, rel, &d);
  if (diff) *diff = d;
  return TDR_OK;
}

int tdr_scan_set_lut(tdr_ctx* c, const int32_t* lut, int n_lut, int C) { c->lut.assign(lut, lut + n_lut); c->n_lut = n_lut; c->scan_C = C; return TDR_OK; }
int tdr_scan_set_points(tdr_ctx* c, const void* pts, int stride, int off, int64_t n) {
  c->pts.assign((const uint8_t*)pts, (const uint8_t*)pts + (size_t)n * stride); c->pts_stride = stride; c->pts_off = off; c->n_pts = (long)n;
  return TDR_OK;
}
int tdr_scan_render_polar(tdr_ctx* c, float res, float ang_res, int n_theta, int n_r, float* imgs) {
  REQ(!c->lut.empty(), TDR_ESTATE, "no lut");
  c->scan.assign((size_t)c->scan_C * n_theta * n_r, 0.f);
  orc_render_polar(c->pts.data(), c->pts_stride, c->pts_off, c->n_pts, res, ang_res, n_theta, n_r, c->lut.data(), c->n_lut, c->scan_C,
                   c->scan.data());
  if (imgs) std::copy(c->scan.begin(), c->scan.end(), imgs);
  return TDR_OK;
}
int tdr_scan_render_cart(tdr_ctx* c, float res, int rows, int cols, float* imgs) {
  REQ(!c->lut.empty() && imgs, TDR_ESTATE, "no lut");
  orc_render_cart(c->pts.data(), c->pts_stride, c->pts_off, c->n_pts, res, rows, cols, c->lut.data(), c->n_lut, c->scan_C, imgs);
  return TDR_OK;
}
int tdr_scan_render_geometric_polar(tdr_ctx* c, int width, int height, float res, float ang_res, int n_theta, int n_r, float* imgs) {
  REQ((long)width * height == c->n_pts && imgs, TDR_EINVAL, "organised cloud expected");
  orc_render_geometric_polar(c->pts.data(), c->pts_stride, width, height, res, ang_res, n_theta, n_r, imgs);
  return TDR_OK;
}
int tdr_scan_render_geometric_cart(tdr_ctx* c, int width, int height, float res, int rows, int cols, float* imgs) {
  REQ((long)width * height == c->n_pts && imgs, TDR_EINVAL, "organised cloud expected");
  orc_render_geometric_cart(c->pts.data(), c->pts_stride, width, height, res, rows, cols, imgs);
  return TDR_OK;
}
int tdr_scan_set_polar_images(tdr_ctx* c, const float* imgs, int n_theta, int n_r, int C) {
  c->scan.assign(imgs, imgs + (size_t)C * n_theta * n_r); c->scan_C = C;
  return TDR_OK;
}

int tdr_pf_set_params(tdr_ctx* c, const tdr_filter_params* p) { c->fp = *p; return TDR_OK; }
int tdr_pf_set_search(tdr_ctx* c, const float* thetas, const int32_t* shifts, int n) {
  c->thetas.assign(thetas, thetas + n); c->shifts.assign(shifts, shifts + n);
  return TDR_OK;
}
int tdr_pf_set_states(tdr_ctx* c, const tdr_state* st, const float* ld, int64_t n) {
  c->states.resize((size_t)n);
  std::memcpy(c->states.data(), st, (size_t)n * sizeof(tdr_state));
  if (ld) c->last_dist.assign(ld, ld + n); else c->last_dist.assign((size_t)n, 0.f);
  return TDR_OK;
}
int tdr_pf_get_states(tdr_ctx* c, tdr_state* st, int64_t n) {
  REQ(n == (int64_t)c->states.size(), TDR_EINVAL, "bad state count");
  std::memcpy(st, c->states.data(), (size_t)n * sizeof(tdr_state));
  return TDR_OK;
}
int tdr_pf_count(tdr_ctx* c, int64_t* n) { *n = (int64_t)c->states.size(); return TDR_OK; }
int tdr_pf_get_last_dist(tdr_ctx* c, float* ld, int64_t n) {
  REQ(n == (int64_t)c->last_dist.size(), TDR_EINVAL, "bad last_dist buffer");
  std::copy(c->last_dist.begin(), c->last_dist.end(), ld);
  return TDR_OK;
}
int tdr_pf_get_weights(tdr_ctx* c, float* w, int64_t n) {
  REQ(n > 0 && n <= (int64_t)c->weights.size(), TDR_EINVAL, "bad weight count");
  std::copy(c->weights.begin(), c->weights.begin() + n, w);
  return TDR_OK;
}
// StateParticle::propagate (state_particle.cpp:57-78) with the noise given as standard variates: `z * stddev + mean`
int tdr_pf_propagate(tdr_ctx* c, float tx, float ty, float omega, int freeze, float pos_cov, float theta_cov, const float* z, int64_t n) {
  REQ(n == (int64_t)c->states.size() && z, TDR_EINVAL, "bad variates");
  for (int64_t i = 0; i < n; i++) {
    OrcState& s = c->states[i];
    const float cs = std::cos(s.theta), sn = std::sin(s.theta);
    const float gx = cs * tx - sn * ty, gy = sn * tx + cs * ty;
    const float lx = s.dx_m, ly = s.dy_m;
    s.dx_m += gx; s.dy_m += gy;
    const float dist = std::sqrt(gx * gx + gy * gy);
    const float sd_pos = pos_cov * dist, sd_th = theta_cov * dist;
    s.theta += (z[4 * i] * sd_th + 0.f) + omega;
    s.dx_m += z[4 * i + 1] * sd_pos + 0.f;
    s.dy_m += z[4 * i + 2] * sd_pos + 0.f;
    if (!freeze) s.scale *= z[4 * i + 3] * static_cast<float>(std::min(2. / dist, 0.02)) + 1.f;
    const float mx = lx - s.dx_m, my = ly - s.dy_m;
    c->last_dist[i] = std::sqrt(mx * mx + my * my);
  }
  return TDR_OK;
}
int tdr_pf_propagate_rng(tdr_ctx*, float, float, float, int, float, float, uint64_t, uint64_t, float*) {
  g_err = "the CPU stand-in has no device RNG (TDR_ADAPTER_DEVICE_RNG needs the real library)";
  return TDR_EUNSUPPORTED;
}
int tdr_pf_gmm_samples(tdr_ctx* c, int num, double* out) {
  REQ(!c->states.empty() && num > 0, TDR_ESTATE, "no particles");
  orc_gmm_samples(c->states.data(), (long)c->states.size(), num, out);
  return TDR_OK;
}
// ParticleFilter::update (particle_filter.cpp:94-189) in its three stages, and in one call
int tdr_pf_score(tdr_ctx* c, float res, float* weights_out) {
  REQ(c->rows > 0 && !c->tab.empty() && !c->scan.empty() && !c->states.empty() && !c->thetas.empty(), TDR_ESTATE, "score before setup");
  const long n = (long)c->states.size();
  OrcFilterParams fp{};
  fp.regularization = c->fp.regularization; fp.force_on_map = c->fp.force_on_map; fp.fixed_scale = c->fp.fixed_scale;
  fp.scale_log_min = c->fp.scale_log_min; fp.scale_log_max = c->fp.scale_log_max; fp.num_classes = c->fp.num_classes;
  fp.map_width = (float)c->cols * c->resolution; fp.map_height = (float)c->rows * c->resolution;
  std::copy(c->fp.class_weights, c->fp.class_weights + 16, fp.class_weights);
  c->weights.assign((size_t)n, 0.f);
  const int nt = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  orc_score_all(c->states.data(), n, &fp, c->layers.data(), c->mask.data(), nullptr, c->rows, c->cols, c->resolution, c->tab.data(),
                c->n_theta, c->n_r, c->scan.data(), res, c->thetas.data(), c->shifts.data(), (int)c->thetas.size(), c->weights.data(), nt);
  if (c->keep_raw) c->raw = c->weights;
  if (weights_out) std::copy(c->weights.begin(), c->weights.end(), weights_out);
  return TDR_OK;
}
int tdr_pf_set_weights(tdr_ctx* c, const float* w, int64_t n) { c->weights.assign(w, w + n); return TDR_OK; }
int tdr_pf_normalize(tdr_ctx* c, int64_t* argmax_out, float* stats) {
  REQ(c->weights.size() == c->states.size() && !c->states.empty(), TDR_ESTATE, "weights and particles differ");
  const long arg = orc_normalize(c->weights.data(), c->last_dist.data(), (long)c->states.size(), stats);
  c->ml_state = c->states[arg]; c->have_ml = true;
  if (argmax_out) *argmax_out = arg;
  return TDR_OK;
}
int tdr_pf_resample(tdr_ctx* c, float u, int64_t M, int32_t* idx_out) {
  REQ(c->weights.size() == c->states.size() && !c->states.empty() && M > 0, TDR_ESTATE, "resample before normalise");
  std::vector<int> idx((size_t)M);
  orc_resample_fast(c->weights.data(), (long)c->states.size(), u, (int)M, idx.data(), nullptr);
  std::vector<OrcState> ns((size_t)M); std::vector<float> nl((size_t)M);
  for (int64_t i = 0; i < M; i++) { ns[i] = c->states[idx[i]]; nl[i] = c->last_dist[idx[i]]; }
  c->states.swap(ns); c->last_dist.swap(nl);
  c->prev_states.swap(ns); c->prev_last_dist.swap(nl);          // the set the resampling read from
  if (idx_out) std::copy(idx.begin(), idx.end(), idx_out);
  return TDR_OK;
}
int tdr_pf_get_prev_states(tdr_ctx* c, tdr_state* st, float* ld, int64_t n) {
  REQ(st && n > 0 && n <= (int64_t)c->prev_states.size(), TDR_EINVAL, "bad state count");
  std::memcpy(st, c->prev_states.data(), (size_t)n * sizeof(tdr_state));
  if (ld) std::copy(c->prev_last_dist.begin(), c->prev_last_dist.begin() + n, ld);
  return TDR_OK;
}
int tdr_pf_keep_raw_weights(tdr_ctx* c, int on) { c->keep_raw = on != 0; return TDR_OK; }
int tdr_pf_get_raw_weights(tdr_ctx* c, float* w, int64_t n) {
  REQ(w && n > 0 && n <= (int64_t)c->raw.size(), TDR_EINVAL, "bad raw weight count");
  std::copy(c->raw.begin(), c->raw.begin() + n, w);
  return TDR_OK;
}
int tdr_pf_update(tdr_ctx* c, float res, float u, int64_t M) {
  if (int e = tdr_pf_score(c, res, nullptr)) return e;
  if (int e = tdr_pf_normalize(c, nullptr, nullptr)) return e;
  return tdr_pf_resample(c, u, M, nullptr);
}
int tdr_pf_pose(tdr_ctx* c, float mean[4], float cov_mean[16], float ml[4], float cov_ml[16]) {
  REQ(!c->states.empty(), TDR_ESTATE, "no particles");
  const long n = (long)c->states.size();
  float m[4], cv[16];
  if (mean || cov_mean) {
    orc_mean_cov(c->states.data(), n, m, cv);
    if (mean) std::copy(m, m + 4, mean);
    if (cov_mean) std::copy(cv, cv + 16, cov_mean);
  }
  if (ml || cov_ml) {
    // the arg-max particle of the last update (max_likelihood_particle_, particle_filter.cpp:147), particle 0 before one;
    // orc_ml_cov takes it by index, so it rides along as entry 0 of a copy whose other entries are the current set
    std::vector<OrcState> tmp(1, c->have_ml ? c->ml_state : c->states[0]);
    float mlv[4];
    orc_ml_cov(tmp.data(), 1, 0, mlv, cv);
    if (ml) std::copy(mlv, mlv + 4, ml);
    if (cov_ml) return fail(TDR_EUNSUPPORTED, "cov about the ML pose is not part of the CPU stand-in");
  }
  return TDR_OK;
}
}  // extern "C"
