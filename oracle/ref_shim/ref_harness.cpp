// ref_harness.cpp — TEST INFRASTRUCTURE ONLY: a C interface over the reference's OWN classes, built from the reference's
// unmodified sources (scan_renderer.cpp, scan_renderer_polar.cpp, top_down_map.cpp, top_down_map_polar.cpp,
// state_particle.cpp, particle_filter.cpp, active_localizer.cpp under /root/reference/src) against the stand-in headers of this directory.
// tests/test_ref_build.py drives it beside the oracle.  See README.md here for what this does and does not pin.
#include "top_down_render/scan_renderer_polar.h"
// TDR_ADAPTER_BUILD: the same harness over the ADAPTERS (top_down_renderer_b200/adapters/*.cpp) instead of the reference's own
// bodies (active_localizer.cpp stays the reference's, on top of the adapter map); see `make -C oracle _adapters`
#include "top_down_render/particle_filter.h"

#define REF_API extern "C" __attribute__((visibility("default")))

static pcl::PointCloud<pcl::PointXYZI>::ConstPtr make_cloud(const float* aos, long n) {
  auto c = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
  c->points.resize((size_t)n);
  std::memcpy(c->points.data(), aos, (size_t)n * 32);
  c->width = (uint32_t)n; c->height = 1;
  return c;
}
static Eigen::VectorXi make_lut(const int* lut, int n) { Eigen::VectorXi v(n); for (int i = 0; i < n; i++) v[i] = lut[i]; return v; }
static std::vector<Eigen::ArrayXXf> make_imgs(int C, int rows, int cols, const float* src = nullptr) {
  std::vector<Eigen::ArrayXXf> v(C, Eigen::ArrayXXf(rows, cols));
  if (src) for (int c = 0; c < C; c++) std::memcpy(v[c].data(), src + (size_t)c * rows * cols, (size_t)rows * cols * 4);
  return v;
}
static void copy_out(const std::vector<Eigen::ArrayXXf>& v, float* dst) {
  for (size_t c = 0; c < v.size(); c++) std::memcpy(dst + c * v[c].size(), v[c].data(), (size_t)v[c].size() * 4);
}

// ---- a1 / a2: the reference's renderers --------------------------------------------------------------------------
REF_API void ref_render_polar(const float* pts, long n, float res, float ang_res, int n_theta, int n_r, const int* lut, int n_lut,
                              int C, float* imgs) {
  ScanRendererPolar r(make_lut(lut, n_lut));
  auto v = make_imgs(C, n_theta, n_r);
  r.renderSemanticTopDown(make_cloud(pts, n), res, ang_res, v);
  copy_out(v, imgs);
}
REF_API void ref_render_cart(const float* pts, long n, float res, int rows, int cols, const int* lut, int n_lut, int C, float* imgs) {
  ScanRenderer r(make_lut(lut, n_lut));
  auto v = make_imgs(C, rows, cols);
  r.renderSemanticTopDown(make_cloud(pts, n), res, v);
  copy_out(v, imgs);
}

// ---- f2: the reference's geometric renderers on an organised cloud (width columns x height rows)
static pcl::PointCloud<pcl::PointXYZI>::ConstPtr make_cloud_wh(const float* aos, int width, int height) {
  auto c = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
  c->points.resize((size_t)width * height);
  std::memcpy(c->points.data(), aos, (size_t)width * height * 32);
  c->width = (uint32_t)width; c->height = (uint32_t)height;
  return c;
}
REF_API void ref_render_geometric_polar(const float* pts, int width, int height, float res, float ang_res, int n_theta, int n_r, float* imgs) {
  ScanRendererPolar r(make_lut(nullptr, 0));
  auto v = make_imgs(2, n_theta, n_r);
  r.renderGeometricTopDown(make_cloud_wh(pts, width, height), res, ang_res, v);
  copy_out(v, imgs);
}
REF_API void ref_render_geometric_cart(const float* pts, int width, int height, float res, int rows, int cols, float* imgs) {
  ScanRenderer r(make_lut(nullptr, 0));
  auto v = make_imgs(2, rows, cols);
  r.renderGeometricTopDown(make_cloud_wh(pts, width, height), res, v);
  copy_out(v, imgs);
}

// ---- a7: TopDownMapPolar on installed layers ---------------------------------------------------------------------------
REF_API void* ref_map_create(const float* layers, const uint8_t* mask, const float* geo, int rows, int cols, int C, float resolution,
                             const float* tab, int n_theta, int n_r, int center_x, int center_y) {
  TopDownMap::Params p;
  p.num_classes = C; p.resolution = resolution;
  auto* m = new TopDownMapPolar(p);
  m->class_maps_ = make_imgs(C, rows, cols, layers);
  m->class_mask_ = Eigen::ArrayXXc(rows, cols);
  std::memcpy(m->class_mask_.data(), mask, (size_t)rows * cols);
  if (geo) m->geo_maps_ = make_imgs(2, rows, cols, geo); else m->geo_maps_ = make_imgs(2, rows, cols);
  m->ang_sample_pts_ = Eigen::Array2Xf(2, n_theta * n_r);
  std::memcpy(m->ang_sample_pts_.data(), tab, (size_t)2 * n_theta * n_r * 4);
  m->map_center_ = Eigen::Vector2i(center_x, center_y);
  m->have_map_ = true;
  return m;
}
REF_API void ref_map_local_polar(void* map, float cx, float cy, float scale, float res, int n_theta, int n_r, float* dists, uint8_t* mask) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  auto v = make_imgs(m->numClasses(), n_theta, n_r);
  Eigen::ArrayXXc k(n_theta, n_r);
  m->getLocalMap(Eigen::Vector2f(cx, cy), scale, res, v, k);
  copy_out(v, dists);
  std::memcpy(mask, k.data(), (size_t)n_theta * n_r);
}
REF_API void ref_map_local_geo_polar(void* map, float cx, float cy, float scale, float res, int n_theta, int n_r, float* geo) {
  auto v = make_imgs(2, n_theta, n_r);
  static_cast<TopDownMapPolar*>(map)->getLocalGeoMap(Eigen::Vector2f(cx, cy), scale, res, v);
  copy_out(v, geo);
}
REF_API void ref_active_best_rel_pos(void* map, const float* preds, int n, float rel[2]) {
  ActiveLocalizer al(static_cast<TopDownMapPolar*>(map));
  std::vector<Eigen::Vector3f> p;
  for (int i = 0; i < n; i++) p.push_back(Eigen::Vector3f(preds[3 * i], preds[3 * i + 1], preds[3 * i + 2]));
  Eigen::Vector2f best = al.getBestRelPos(p);
  rel[0] = best[0]; rel[1] = best[1];
}

// ---- a3 - a6, a8, the vector map and the caches: src/top_down_map.cpp itself -------------------------------------------
static TopDownMap::Params make_params(int C, float resolution, const int* lut, int n_lut) {
  TopDownMap::Params p;
  p.num_classes = C; p.resolution = resolution;
  if (lut) p.flatten_lut.assign(lut, lut + n_lut);
  return p;
}
// the dynamic-map path: TopDownMapPolar(params) then updateMap(image, centre) (top_down_render.cpp:81, :591)
REF_API void* ref_map_from_class_image(const uint8_t* img, int h, int w, int stride, const int* lut, int n_lut, int C, float resolution,
                                       int center_x, int center_y) {
  auto* m = new TopDownMapPolar(make_params(C, resolution, lut, n_lut));
  std::vector<uint8_t> packed((size_t)h * w);
  for (int r = 0; r < h; r++) std::memcpy(&packed[(size_t)r * w], img + (size_t)r * stride, (size_t)w);
  m->updateMap(cv::Mat(h, w, CV_8UC1, packed.data()), Eigen::Vector2i(center_x, center_y));
  return m;
}
// the static-map constructor (top_down_map.cpp:9-64): cache hit -> loadCachedMaps; `.svg` -> loadSvg (nanosvg, vendored in
// the reference), getRasterMap, saveRasterizedMaps; a directory -> loadRasterizedMaps; then geo maps, computeDists,
// saveCachedMaps.  home: what $HOME is set to for the cache directory.  class_colors: packed fill colour per class index.
REF_API void* ref_map_from_path(const char* home, const char* map_path, const int* lut, int n_lut, int C, float resolution,
                                const uint32_t* class_colors, const int* exclusive, int n_excl) {
  setenv("HOME", home, 1);
  TopDownMap::Params p = make_params(C, resolution, lut, n_lut);
  p.map_path = map_path;
  p.color_lut.packed.assign(class_colors, class_colors + n_lut);
  p.exclusive_classes.assign(exclusive, exclusive + n_excl);
  return new TopDownMapPolar(p);
}
REF_API void ref_map_info(void* map, int* rows, int* cols, int* C, int* have_map, int* center_xy) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  *rows = m->class_maps_.empty() ? 0 : (int)m->class_maps_[0].rows();
  *cols = m->class_maps_.empty() ? 0 : (int)m->class_maps_[0].cols();
  *C = (int)m->class_maps_.size(); *have_map = m->haveMap() ? 1 : 0;
  center_xy[0] = m->mapCenter()[0]; center_xy[1] = m->mapCenter()[1];
}
REF_API void ref_map_get(void* map, float* layers, uint8_t* mask, float* geo) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  if (layers) copy_out(m->class_maps_, layers);
  if (mask) std::memcpy(mask, m->class_mask_.data(), (size_t)m->class_mask_.size());
  if (geo && m->geo_maps_.size() == 2) copy_out(m->geo_maps_, geo);
}
#ifndef TDR_ADAPTER_BUILD
// what the static constructor does after loading (:47-58), for a map that came through updateMap: geo maps + their distances
REF_API void ref_map_build_geo(void* map) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  // updateMap leaves class_maps_ as DISTANCE fields; getGeoRasterMap wants the binary maps, which the caller re-installs
  m->geo_maps_.clear();
  for (size_t i = 0; i < 2; i++) m->geo_maps_.push_back(Eigen::ArrayXXf(m->class_maps_[0].rows(), m->class_maps_[0].cols()));
  m->getGeoRasterMap(m->geo_maps_);
  Eigen::ArrayXXc tmp;
  m->computeDists(m->geo_maps_, tmp);
}
#endif
REF_API void ref_map_set_class_maps(void* map, const float* layers, int rows, int cols, int C) {
  static_cast<TopDownMapPolar*>(map)->class_maps_ = make_imgs(C, rows, cols, layers);
}
REF_API void ref_map_polar_table(void* map, int n_theta, int n_r, float ang_res, float* tab) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  m->samplePtsPolar(Eigen::Vector2i(n_theta, n_r), ang_res);
  std::memcpy(tab, m->ang_sample_pts_.data(), (size_t)2 * n_theta * n_r * 4);
}
REF_API void ref_map_set_polar_table(void* map, const float* tab, int n_theta, int n_r) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  m->ang_sample_pts_ = Eigen::Array2Xf(2, n_theta * n_r);
  std::memcpy(m->ang_sample_pts_.data(), tab, (size_t)2 * n_theta * n_r * 4);
}
REF_API unsigned ref_map_classes_at(void* map, int as_float, float x, float y) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  std::vector<int> cls;
  if (as_float) m->TopDownMap::getClassesAtPoint(Eigen::Vector2f(x, y), cls);
  else m->TopDownMap::getClassesAtPoint(Eigen::Vector2i((int)x, (int)y), cls);
  unsigned bits = 0;
  for (int c : cls) bits |= 1u << c;
  return bits;
}
// a8: the Cartesian TopDownMap::getLocalMap (hidden by the polar overloads: called through the base class)
REF_API void ref_map_local_cart(void* map, float cx, float cy, float rot, float res, int rows, int cols, float* dists, uint8_t* mask) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  auto v = make_imgs(m->numClasses(), rows, cols);
  Eigen::ArrayXXc k(rows, cols);
  static_cast<TopDownMap*>(m)->getLocalMap(Eigen::Vector2f(cx, cy), rot, res, v, k);
  copy_out(v, dists);
  std::memcpy(mask, k.data(), (size_t)rows * cols);
}

// ---- a9 - a13 and the rows around them: ParticleFilter ------------------------------------------------------------------
struct RefFilterParams {      // FilterParams, state_particle.h:19-38
  float pos_cov, theta_cov, regularization;
  float init_pos_px_x, init_pos_px_y, init_pos_px_cov, init_pos_m_x, init_pos_m_y, init_pos_deg_theta, init_pos_deg_cov;
  int force_on_map;
  float fixed_scale, scale_log_min, scale_log_max;
  float class_weights[16];
  int n_class_weights;
};
static_assert(sizeof(State) == 28, "State is 28 bytes");

// The constructor seeds its engine from std::random_device (particle_filter.cpp:4-5) and initialises at once when the map
// exists; here the engine is re-seeded before initializeParticles runs so that a run can be repeated.
REF_API void* ref_filter_create(void* map, int N, const RefFilterParams* rp, uint32_t seed, int initialize) {
  auto* m = static_cast<TopDownMapPolar*>(map);
  FilterParams fp;
  fp.pos_cov = rp->pos_cov; fp.theta_cov = rp->theta_cov; fp.regularization = rp->regularization;
  fp.init_pos_px_x = rp->init_pos_px_x; fp.init_pos_px_y = rp->init_pos_px_y; fp.init_pos_px_cov = rp->init_pos_px_cov;
  fp.init_pos_m_x = rp->init_pos_m_x; fp.init_pos_m_y = rp->init_pos_m_y;
  fp.init_pos_deg_theta = rp->init_pos_deg_theta; fp.init_pos_deg_cov = rp->init_pos_deg_cov;
  fp.force_on_map = rp->force_on_map != 0; fp.fixed_scale = rp->fixed_scale;
  fp.scale_log_min = rp->scale_log_min; fp.scale_log_max = rp->scale_log_max;
  fp.class_weights.assign(rp->class_weights, rp->class_weights + rp->n_class_weights);
  const bool had = m->have_map_;
  m->have_map_ = false;
  auto* f = new ParticleFilter(N, m, fp);
  m->have_map_ = had;
  *f->gen_ = std::mt19937(seed);
  if (initialize) f->initializeParticles();
  return f;
}
// the adapters keep particles_ / new_particles_ / weights_ as a LAZY mirror of the device set: visualize is the reader
// that refreshes it (adapters/particle_filter_adapter.cpp); the reference's own classes need nothing
#ifdef TDR_ADAPTER_BUILD
extern "C" void tdr_adapter_mark_host_ahead(const void* filter);
#endif
static void sync_mirror(ParticleFilter* f) {
#ifdef TDR_ADAPTER_BUILD
  cv::Mat none;
  f->visualize(none);
#else
  (void)f;
#endif
}
REF_API long ref_filter_count(void* fv) { sync_mirror(static_cast<ParticleFilter*>(fv)); return (long)static_cast<ParticleFilter*>(fv)->particles_.size(); }
REF_API int ref_filter_num_particles(void* fv) { return static_cast<ParticleFilter*>(fv)->numParticles(); }
// which: 0 = particles_ (the current set), 1 = new_particles_ (after an update: the set that was scored)
REF_API long ref_filter_get(void* fv, int which, State* st, float* last_dist, float* raw_weight, long cap) {
  sync_mirror(static_cast<ParticleFilter*>(fv));
  auto* f = static_cast<ParticleFilter*>(fv);
  std::lock_guard<std::mutex> g(f->particle_lock_);
  auto& v = which ? f->new_particles_ : f->particles_;
  long n = std::min<long>(cap, (long)v.size());
  for (long i = 0; i < n; i++) {
    if (st) st[i] = v[i]->state();
    if (last_dist) last_dist[i] = v[i]->lastDist();
    if (raw_weight) raw_weight[i] = v[i]->weight();
  }
  return (long)v.size();
}
REF_API void ref_filter_set(void* fv, const State* st, const float* last_dist, long n) {
  sync_mirror(static_cast<ParticleFilter*>(fv));
  auto* f = static_cast<ParticleFilter*>(fv);
  std::lock_guard<std::mutex> g(f->particle_lock_);
  for (long i = 0; i < n && i < (long)f->particles_.size(); i++) {
    f->particles_[i]->setState(st[i]);
    if (last_dist) f->particles_[i]->last_dist_ = last_dist[i];
  }
#ifdef TDR_ADAPTER_BUILD
  tdr_adapter_mark_host_ahead(f);
#endif
}
REF_API long ref_filter_weights(void* fv, float* w, long cap) {
  sync_mirror(static_cast<ParticleFilter*>(fv));
  auto* f = static_cast<ParticleFilter*>(fv);
  long n = std::min<long>(cap, (long)f->weights_.size());
  for (long i = 0; i < n; i++) w[i] = f->weights_[i];
  return (long)f->weights_.size();
}
REF_API void ref_filter_propagate(void* fv, float tx, float ty, float omega) {
  Eigen::Vector2f t(tx, ty);
  static_cast<ParticleFilter*>(fv)->propagate(t, omega);
}
REF_API void ref_filter_update(void* fv, const float* scan, int n_theta, int n_r, int C, float res) {
  auto v = make_imgs(C, n_theta, n_r, scan);
  auto geo = make_imgs(2, n_theta, n_r);
  static_cast<ParticleFilter*>(fv)->update(v, geo, res);
}
REF_API void ref_filter_pose(void* fv, float mean[4], float cov_mean[16], float ml[4], float cov_ml[16]) {
  auto* f = static_cast<ParticleFilter*>(fv);
  Eigen::Vector4f v; Eigen::Matrix4f c;
  if (mean) { f->meanLikelihood(v); std::memcpy(mean, v.data(), 16); }
  if (cov_mean) { f->computeMeanCov(c); std::memcpy(cov_mean, c.data(), 64); }
  if (ml) { f->maxLikelihood(v); std::memcpy(ml, v.data(), 16); }
  if (cov_ml) { f->computeCov(c); std::memcpy(cov_ml, c.data(), 64); }
}
REF_API void ref_filter_freeze_scale(void* fv) { static_cast<ParticleFilter*>(fv)->freezeScale(); }
REF_API float ref_filter_scale(void* fv) { return static_cast<ParticleFilter*>(fv)->scale(); }
REF_API int ref_filter_scale_frozen(void* fv) { return static_cast<ParticleFilter*>(fv)->isScaleFrozen() ? 1 : 0; }
// ParticleFilter::updateMap (:320-341) with a class-index image; the filter's map must carry a flatten_lut
REF_API void ref_filter_update_map(void* fv, const uint8_t* img, int h, int w, int center_x, int center_y) {
  std::vector<uint8_t> copy(img, img + (size_t)h * w);
  static_cast<ParticleFilter*>(fv)->updateMap(cv::Mat(h, w, CV_8UC1, copy.data()), Eigen::Vector2i(center_x, center_y));
}
REF_API void ref_filter_init_px(void* fv, float px[2]) {
  auto* f = static_cast<ParticleFilter*>(fv);
  px[0] = f->params_.init_pos_px_x; px[1] = f->params_.init_pos_px_y;
}
// the engine's next 32-bit output, from a copy: tells how many outputs the calls so far consumed
REF_API uint32_t ref_filter_engine_peek(void* fv) { std::mt19937 copy = *static_cast<ParticleFilter*>(fv)->gen_; return copy(); }
// computeGMM on the caller's thread: the sample matrix it built (rows x 4 doubles) and the mixture it stored
REF_API int ref_filter_gmm(void* fv, double* samples, int cap_rows, float* means3, float* covs9, int cap_clusters) {
  auto* f = static_cast<ParticleFilter*>(fv);
  f->computeGMM();
  const cv::Mat& s = cv::ml::EM::lastSamples();
  for (int i = 0; i < s.rows && i < cap_rows; i++) for (int j = 0; j < 4; j++) samples[4 * i + j] = s.at<double>(i, j);
  std::vector<Eigen::Vector3f> means; std::vector<Eigen::Matrix3f> covs;
  f->getGMM(means, covs);
  for (size_t k = 0; k < means.size() && (int)k < cap_clusters; k++) {
    for (int j = 0; j < 3; j++) means3[3 * k + j] = means[k][j];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) covs9[9 * k + 3 * r + c] = covs[k](r, c);
  }
  return s.rows;
}
