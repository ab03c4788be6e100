// tdr_host.hpp — C++ host side above the C ABI (include/tdr.h): the reference's class interfaces for the hot path,
// with the SAME names, argument meaning and (silent) error behaviour, on dependency-free types.
//
// The reference's classes use Eigen / PCL / OpenCV / ROS types, none of which exist in this image, so
// this mirror swaps them for layout-compatible plain types (column-major ArrayXXf, 32-byte PointXYZI, raw uint8
// image); INTEGRATION.md shows the one-line conversions the real adapters add.  Everything with arithmetic lives
// behind the C ABI on the device; what stays here is what the reference also keeps on the host: particle
// initialisation and propagation (libstdc++ RNG), the polar offset table and the theta-search list.
//
//   ScanRenderer        include/top_down_render/scan_renderer.h:14-23
//   ScanRendererPolar   include/top_down_render/scan_renderer_polar.h:15-22
//   TopDownMap          include/top_down_render/top_down_map.h:52-101
//   TopDownMapPolar     include/top_down_render/top_down_map_polar.h:6-22
//   ActiveLocalizer     include/top_down_render/active_localizer.h:7-16
//   ParticleFilter      include/top_down_render/particle_filter.h:22-41
#pragma once
#include <math.h>
#include <sys/stat.h>
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "../../include/tdr.h"
#include "png_gray.hpp"

namespace tdrhost {

// Eigen::ArrayXXf / ArrayXXc stand-ins: column-major, element (r, c) at c*rows + r
template <typename T> struct Array2D {
  int rows_ = 0, cols_ = 0;
  std::vector<T> d;
  Array2D() {}
  Array2D(int r, int c) : rows_(r), cols_(c), d((size_t)r * c, T(0)) {}
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
  T& operator()(int r, int c) { return d[(size_t)c * rows_ + r]; }
  const T& operator()(int r, int c) const { return d[(size_t)c * rows_ + r]; }
};
using ArrayXXf = Array2D<float>;
using ArrayXXc = Array2D<uint8_t>;
struct Vector2f { float x = 0, y = 0; };
struct Vector2i { int x = 0, y = 0; };
struct Vector3f { float x = 0, y = 0, z = 0; };   // ActiveLocalizer's (x, y, theta) predictions

// pcl::PointXYZI layout (32 bytes; intensity at byte 16)
struct PointXYZI { float x, y, z, pad0, intensity, pad1, pad2, pad3; };
static_assert(sizeof(PointXYZI) == 32, "PointXYZI must be 32 bytes");

using State = tdr_state;   // state_particle.h:9-17, identical layout

// state_particle.h:19-38
struct FilterParams {
  float pos_cov = 0.3f, theta_cov = 0.0314f, regularization = 0.15f;
  float init_pos_px_x = -1, init_pos_px_y = -1, init_pos_px_cov = -1;
  float init_pos_m_x = -1, init_pos_m_y = -1, init_pos_deg_theta = -1, init_pos_deg_cov = -1;   // the node passes +inf for "unset" (top_down_render.cpp:214-231)
  bool force_on_map = false;
  float fixed_scale = -1, scale_log_min = -0.1f, scale_log_max = 1;
  std::vector<float> class_weights;
};

// one process-wide device context, like the reference's single ROS-thread ownership
inline tdr_ctx* ctx() {
  static tdr_ctx* c = nullptr;
  if (!c) {
    const char* dev = getenv("TDR_DEVICE");
    if (tdr_create(&c, dev ? atoi(dev) : 0) != TDR_OK) {
      fprintf(stderr, "[XView] libtdr_b200: %s\n", tdr_last_error());   // the reference logs and carries on
      c = nullptr;
    }
  }
  return c;
}
inline bool ok(int rc) {
  if (rc != TDR_OK) fprintf(stderr, "[XView] libtdr_b200: %s\n", tdr_last_error());
  return rc == TDR_OK;
}

// scan_renderer.h:14-23
class ScanRenderer {
 public:
  explicit ScanRenderer(const std::vector<int>& flatten_lut) : flatten_lut_(flatten_lut) {}
  // scan_renderer.cpp:55-78.  imgs: pre-sized (rows x cols) images; their size defines the raster, centred on the sensor.
  void renderSemanticTopDown(const std::vector<PointXYZI>& cloud, float res, std::vector<ArrayXXf>& imgs) {
    if (imgs.size() < 1 || !ctx()) return;                                  // :57
    int max_cls = -1;
    for (int v : flatten_lut_) max_cls = std::max(max_cls, v);
    if ((int)imgs.size() < max_cls + 1) return;                             // the reference indexes imgs[lut] unchecked
    const int C = (int)imgs.size(), rows = imgs[0].rows(), cols = imgs[0].cols();
    if (!ok(tdr_scan_set_lut(ctx(), flatten_lut_.data(), (int)flatten_lut_.size(), C))) return;
    if (!ok(tdr_scan_set_points(ctx(), cloud.data(), sizeof(PointXYZI), 16, (int64_t)cloud.size()))) return;
    stage_.resize((size_t)C * rows * cols);
    if (!ok(tdr_scan_render_cart(ctx(), res, rows, cols, stage_.data()))) return;
    for (int c = 0; c < C; c++)
      std::copy(stage_.begin() + (size_t)c * rows * cols, stage_.begin() + (size_t)(c + 1) * rows * cols, imgs[c].d.begin());
  }

 protected:
  std::vector<int> flatten_lut_;
  std::vector<float> stage_;
};

// scan_renderer_polar.h:15-22: the 4-argument overload hides the Cartesian one, as in the reference
class ScanRendererPolar : public ScanRenderer {
 public:
  explicit ScanRendererPolar(const std::vector<int>& flatten_lut) : ScanRenderer(flatten_lut) {}
  // scan_renderer_polar.cpp:83-109.  imgs: C pre-sized (n_theta x n_r) images; their size defines the raster.
  void renderSemanticTopDown(const std::vector<PointXYZI>& cloud, float res, float ang_res, std::vector<ArrayXXf>& imgs) {
    int max_cls = -1;
    for (int v : flatten_lut_) max_cls = std::max(max_cls, v);
    if ((int)imgs.size() < max_cls + 1 || !ctx()) return;                   // the reference's silent early-out (:85)
    const int C = (int)imgs.size(), n_theta = imgs[0].rows(), n_r = imgs[0].cols();
    if (!ok(tdr_scan_set_lut(ctx(), flatten_lut_.data(), (int)flatten_lut_.size(), C))) return;
    if (!ok(tdr_scan_set_points(ctx(), cloud.data(), sizeof(PointXYZI), 16, (int64_t)cloud.size()))) return;
    stage_.resize((size_t)C * n_theta * n_r);
    if (!ok(tdr_scan_render_polar(ctx(), res, ang_res, n_theta, n_r, stage_.data()))) return;
    for (int c = 0; c < C; c++)
      std::copy(stage_.begin() + (size_t)c * n_theta * n_r, stage_.begin() + (size_t)(c + 1) * n_theta * n_r, imgs[c].d.begin());
  }
};

// top_down_map.h:52-101.  TopDownMapPolar derives from it and HIDES getLocalMap / getLocalGeoMap with the polar versions
// of the same signatures (non-virtual, as in the reference): through a TopDownMap* the Cartesian ones are called.
class TopDownMap {
 public:
  struct Params {                       // top_down_map.h:54-62 (the dynamic-map fields)
    std::vector<int> flatten_lut;
    int num_classes = 0;
    float resolution = 1;
  };
  explicit TopDownMap(const Params& params) : params_(params) {}

  // top_down_map.cpp:146-157: class-index image (row-major, row 0 = top) -> layers -> distance fields
  void updateMap(const uint8_t* img, int rows, int cols, int stride, const Vector2i& map_center) {
    if (!ctx()) return;
    map_center_ = map_center;
    std::vector<int32_t> lut(256, -1);                                     // padded like top_down_render.cpp:57
    std::copy(params_.flatten_lut.begin(), params_.flatten_lut.end(), lut.begin());
    if (!ok(tdr_map_set_class_image(ctx(), img, rows, cols, stride, lut.data(), 256, params_.num_classes, params_.resolution))) return;
    int r = 0, c = 0, k = 0; float res = 1;
    tdr_map_info(ctx(), &r, &c, &k, &res);
    rows_ = r; cols_ = c;
    layers_.resize((size_t)k * r * c); mask_.resize((size_t)r * c);
    ok(tdr_map_get_layers(ctx(), layers_.data(), mask_.data()));           // host mirror for getClassesAtPoint
    // have_map_ is set (and stays set) unless the BINARY layer 1 is entirely zero, i.e. every cell is road — the
    // isZero(0) test of :150 runs before computeDists; cells sample the image as loadCompressedRasterMap does (:137-138)
    bool all_road = k > 1;
    for (int xi = 0; xi < c && all_road; xi++)
      for (int yi = 0; yi < r && all_road; yi++) {
        const int ir = std::max<int>((int)((float)rows - (float)yi * params_.resolution - 1), 0);
        const int ic = std::min<int>((int)((float)xi * params_.resolution), cols - 1);
        all_road = lut[img[(size_t)ir * stride + ic]] == 1;
      }
    if (!all_road) have_map_ = true;
    if (!tab_.empty()) ok(tdr_map_set_polar_table(ctx(), tab_.data(), n_theta_, n_r_));
  }
  // ---- the map cache, top_down_map.cpp:226-286 with write_binary / read_binary of top_down_map.h:29-50: same files, same
  // bytes (Eigen::Index = int64 rows, cols, then column-major scalars), in `dir` instead of $HOME/.ros/xview_cache
  bool loadCacheMetaData(const std::string& dir, const std::string& map_path) const {
    std::ifstream f(dir + "/cached_data.txt");
    if (!f) return false;
    std::string line;
    std::getline(f, line); if (line != map_path) return false;
    std::getline(f, line); if (std::stoi(line) != params_.num_classes) return false;
    std::getline(f, line); if (std::abs(std::stof(line) - params_.resolution) > 0.01) return false;
    return true;
  }
  void saveCachedMaps(const std::string& dir, const std::string& map_path) {
    if (!ctx() || rows_ == 0) return;
    std::ofstream f(dir + "/cached_data.txt", std::ofstream::out | std::ofstream::trunc);
    f << map_path << std::endl << params_.num_classes << std::endl << params_.resolution << std::endl;
    const size_t L = (size_t)rows_ * cols_;
    for (int cls = 0; cls < params_.num_classes; cls++) write_eig(dir + "/class_map" + std::to_string(cls) + ".eig", layers_.data() + cls * L, 4);
    std::vector<float> geo(2 * L);
    if (ok(tdr_map_get_geo_layers(ctx(), geo.data())))
      for (int cls = 0; cls < 2; cls++) write_eig(dir + "/geo_map" + std::to_string(cls) + ".eig", geo.data() + cls * L, 4);
    write_eig(dir + "/class_mask.eig", mask_.data(), 1);
  }
  // a cache hit skips the distance transform: the cached fields go straight to the device (tdr_map_set_dist_layers)
  bool loadCachedMaps(const std::string& dir) {
    if (!ctx()) return false;
    std::vector<float> all;
    int64_t r = 0, c = 0;
    for (int cls = 0; cls < params_.num_classes; cls++) {
      std::vector<char> raw;
      int64_t rr = 0, cc = 0;
      if (!read_eig(dir + "/class_map" + std::to_string(cls) + ".eig", 4, raw, rr, cc) || (cls > 0 && (rr != r || cc != c))) return false;
      r = rr; c = cc;
      all.insert(all.end(), reinterpret_cast<float*>(raw.data()), reinterpret_cast<float*>(raw.data()) + (size_t)r * c);
    }
    std::vector<char> m;
    int64_t mr = 0, mc = 0;
    if (!read_eig(dir + "/class_mask.eig", 1, m, mr, mc) || mr != r || mc != c) return false;
    if (!ok(tdr_map_set_dist_layers(ctx(), all.data(), reinterpret_cast<const uint8_t*>(m.data()), (int)r, (int)c, params_.num_classes,
                                    params_.resolution))) return false;
    rows_ = (int)r; cols_ = (int)c; layers_ = all; mask_.assign(m.begin(), m.end());
    std::vector<float> geo;                                                // geo_map0/1.eig (:252-257)
    for (int cls = 0; cls < 2; cls++) {
      std::vector<char> raw;
      int64_t rr = 0, cc = 0;
      if (!read_eig(dir + "/geo_map" + std::to_string(cls) + ".eig", 4, raw, rr, cc) || rr != r || cc != c) return false;
      geo.insert(geo.end(), reinterpret_cast<float*>(raw.data()), reinterpret_cast<float*>(raw.data()) + (size_t)r * c);
    }
    if (!ok(tdr_map_set_geo_dist_layers(ctx(), geo.data()))) return false;
    have_map_ = true;
    if (!tab_.empty()) ok(tdr_map_set_polar_table(ctx(), tab_.data(), n_theta_, n_r_));
    return true;
  }
  // ---- the vector-map constructor path (top_down_map.cpp:22-31 with getRasterMap :391-408): the polygons of every flattened
  // class as loadSvg leaves them (y already flipped, :89) -> binary class maps (kept for saveRasterizedMaps) -> distance
  // fields, all on the device; exclusive_classes as Params holds them (top_down_render.cpp:177-181)
  void setVectorMap(const std::vector<std::vector<std::vector<Vector2f>>>& poly, const Vector2i& map_size,
                    const std::vector<int>& exclusive_classes) {
    if (poly.size() < 1 || !ctx()) return;                                  // :394
    std::vector<float> verts; std::vector<int32_t> start(1, 0), cls;
    for (size_t c = 0; c < poly.size(); c++)
      for (const auto& path : poly[c]) {
        for (const Vector2f& v : path) { verts.push_back(v.x); verts.push_back(v.y); }
        start.push_back((int32_t)(verts.size() / 2)); cls.push_back((int32_t)c);
      }
    const int r = (int)((float)map_size.y / params_.resolution), c = (int)((float)map_size.x / params_.resolution);   // :399-400
    binary_.assign((size_t)params_.num_classes * r * c, 0.f);
    std::vector<int32_t> excl(exclusive_classes.begin(), exclusive_classes.end());
    if (!ok(tdr_map_set_polygons(ctx(), verts.data(), start.data(), cls.data(), (int)cls.size(), map_size.x, map_size.y, 0.f,
                                 params_.num_classes, params_.resolution, excl.data(), (int)excl.size(), binary_.data()))) return;
    adoptDeviceMap();
  }
  // saveRasterizedMaps :197-211: class<i>.png = the binary class map x 255 (saturate_cast<uchar>), flipped vertically
  void saveRasterizedMaps(const std::string& path) const {
    if (binary_.empty()) return;
    ::mkdir(path.c_str(), S_IRWXU);
    std::vector<uint8_t> img((size_t)rows_ * cols_);
    for (int cls = 0; cls < params_.num_classes; cls++) {
      const float* m = binary_.data() + (size_t)cls * rows_ * cols_;
      for (int r = 0; r < rows_; r++)
        for (int c = 0; c < cols_; c++) {
          const float v = std::nearbyint(m[(size_t)c * rows_ + r] * 255.f);
          img[(size_t)(rows_ - 1 - r) * cols_ + c] = (uint8_t)std::min(std::max(v, 0.f), 255.f);
        }
      png::write_gray(path + "/class" + std::to_string(cls) + ".png", img.data(), cols_, rows_);
    }
  }
  // loadRasterizedMaps :213-224 and what the constructor does next (:47-60): flip back, scale by float(1/255), distance
  // fields on the device.  false when a file is missing or is not an 8-bit gray PNG (the reference crashes there).
  bool loadRasterizedMaps(const std::string& path) {
    if (!ctx()) return false;
    std::vector<float> all;
    int r = 0, c = 0;
    for (int cls = 0; cls < params_.num_classes; cls++) {
      std::vector<uint8_t> img; int w = 0, h = 0;
      if (!png::read_gray(path + "/class" + std::to_string(cls) + ".png", img, w, h) || (cls > 0 && (w != c || h != r))) return false;
      r = h; c = w;
      const size_t at = all.size();
      all.resize(at + (size_t)r * c);
      for (int y = 0; y < r; y++)
        for (int x = 0; x < c; x++) all[at + (size_t)x * r + y] = (float)img[(size_t)(r - 1 - y) * c + x] * (float)(1. / 255);
    }
    if (!ok(tdr_map_set_binary_layers(ctx(), all.data(), r, c, params_.num_classes, params_.resolution))) return false;
    binary_.swap(all);
    adoptDeviceMap();
    return true;
  }
  // top_down_map.cpp:159-170 (integer point)
  void getClassesAtPoint(const Vector2i& center_ind, std::vector<int>& classes) const {
    classes.clear();
    const int cx = (int)((float)center_ind.x / params_.resolution), cy = (int)((float)center_ind.y / params_.resolution);
    for (int cls = 0; cls < params_.num_classes; cls++)
      if (cx < cols_ && cy < rows_ && cx >= 0 && cy >= 0 && layers_[(size_t)cls * rows_ * cols_ + (size_t)cx * rows_ + cy] < 1)
        classes.push_back(cls);
  }
  // top_down_map.cpp:172-175 (float point: divided by the resolution here AND again in the integer overload)
  void getClassesAtPoint(const Vector2f& center, std::vector<int>& classes) const {
    getClassesAtPoint(Vector2i{(int)(center.x / params_.resolution), (int)(center.y / params_.resolution)}, classes);
  }
  // TopDownMap::getLocalMap (Cartesian, top_down_map.cpp:429-459): a rotated rows x cols lattice around `center`; the
  // size of the caller's images defines the lattice.  Only the node's debug view calls it (top_down_render.cpp:315).
  void getLocalMap(Vector2f center, float rot, float res, std::vector<ArrayXXf>& dists, ArrayXXc& mask) {
    if (dists.size() < 1 || !ctx()) return;                                // :432
    const int r = dists[0].rows(), c = dists[0].cols();
    stage_.resize((size_t)params_.num_classes * r * c);
    if (!ok(tdr_map_local_cart(ctx(), center.x, center.y, rot, res, r, c, stage_.data(), mask.data()))) return;
    for (int k = 0; k < params_.num_classes && k < (int)dists.size(); k++)
      std::copy(stage_.begin() + (size_t)k * r * c, stage_.begin() + (size_t)(k + 1) * r * c, dists[k].d.begin());
  }
  Vector2i size() const { return Vector2i{cols_, rows_}; }
  Vector2i mapCenter() const { return map_center_; }
  int numClasses() const { return params_.num_classes; }
  float resolution() const { return params_.resolution; }
  bool haveMap() const { return have_map_; }

 protected:
  // after a constructor-path load: host copies of the distance fields for getClassesAtPoint, have_map_ (:62), table
  void adoptDeviceMap() {
    int r = 0, c = 0, k = 0; float res = 1;
    tdr_map_info(ctx(), &r, &c, &k, &res);
    rows_ = r; cols_ = c;
    layers_.resize((size_t)k * r * c); mask_.resize((size_t)r * c);
    ok(tdr_map_get_layers(ctx(), layers_.data(), mask_.data()));
    have_map_ = true;
    if (!tab_.empty()) ok(tdr_map_set_polar_table(ctx(), tab_.data(), n_theta_, n_r_));
  }
  void write_eig(const std::string& name, const void* data, size_t scalar_bytes) const {
    std::ofstream out(name, std::ios::out | std::ios::binary | std::ios::trunc);
    const int64_t rows = rows_, cols = cols_;
    out.write(reinterpret_cast<const char*>(&rows), 8);
    out.write(reinterpret_cast<const char*>(&cols), 8);
    out.write(reinterpret_cast<const char*>(data), (std::streamsize)((size_t)rows * cols * scalar_bytes));
  }
  static bool read_eig(const std::string& name, size_t scalar_bytes, std::vector<char>& data, int64_t& rows, int64_t& cols) {
    std::ifstream in(name, std::ios::in | std::ios::binary);
    if (!in) return false;
    in.read(reinterpret_cast<char*>(&rows), 8);
    in.read(reinterpret_cast<char*>(&cols), 8);
    if (!in || rows <= 0 || cols <= 0 || rows * cols > (int64_t)1 << 32) return false;
    data.resize((size_t)rows * cols * scalar_bytes);
    in.read(data.data(), (std::streamsize)data.size());
    return (bool)in;
  }
  Params params_;
  bool have_map_ = false;
  Vector2i map_center_;
  int rows_ = 0, cols_ = 0, n_theta_ = 0, n_r_ = 0;
  std::vector<float> layers_, tab_, stage_, binary_;
  std::vector<uint8_t> mask_;
};

// top_down_map_polar.h:6-22
class TopDownMapPolar : public TopDownMap {
 public:
  explicit TopDownMapPolar(const Params& params) : TopDownMap(params) { samplePtsPolar(100, 50, (float)(2 * M_PI / 100)); }
  // top_down_map_polar.cpp:21-53
  void getLocalMap(Vector2f center, float scale, float res, std::vector<ArrayXXf>& dists, ArrayXXc& mask) {
    if ((int)dists.size() < params_.num_classes || !ctx()) return;         // :25
    const int P = n_theta_ * n_r_;
    stage_.resize((size_t)params_.num_classes * P);
    float cxy[2] = {center.x, center.y};
    if (!ok(tdr_map_local_polar(ctx(), cxy, 1, scale, res, stage_.data(), mask.data()))) return;
    for (int c = 0; c < params_.num_classes; c++) std::copy(stage_.begin() + (size_t)c * P, stage_.begin() + (size_t)(c + 1) * P, dists[c].d.begin());
  }
  void getLocalMap(Vector2f center, float res, std::vector<ArrayXXf>& dists, ArrayXXc& mask) { getLocalMap(center, 1, res, dists, mask); }   // :78-82
  // top_down_map_polar.cpp:55-76 (+ the 3-argument overload :84-87): the two geometric distance layers, no mask
  void getLocalGeoMap(Vector2f center, float scale, float res, std::vector<ArrayXXf>& dists) {
    if (dists.size() < 1 || !ctx()) return;                                // :58
    const int P = n_theta_ * n_r_;
    stage_.resize((size_t)2 * P);
    const float c[2] = {center.x, center.y};
    if (!ok(tdr_map_local_geo_polar(ctx(), c, 1, scale, res, stage_.data()))) return;
    for (size_t k = 0; k < dists.size() && k < 2; k++) std::copy(stage_.begin() + k * P, stage_.begin() + (k + 1) * P, dists[k].d.begin());
  }
  void getLocalGeoMap(Vector2f center, float res, std::vector<ArrayXXf>& dists) { getLocalGeoMap(center, 1, res, dists); }
  // top_down_map_polar.cpp:7-19 (+ samplePts, top_down_map.cpp:367-389): the offset table is computed on the host
  void samplePtsPolar(int n_theta, int n_r, float ang_res) {
    n_theta_ = n_theta; n_r_ = n_r;
    tab_.resize((size_t)2 * n_theta * n_r);
    for (int p = 0; p < n_theta * n_r; p++) {
      const int row = p % n_theta, col = p / n_theta;
      const float a0 = ((float)row - (float)(n_theta - 1) / 2.f) * ang_res;
      const float a1 = (float)col * (1.f / params_.resolution);
      tab_[2 * p] = cosf(a0) * a1;
      tab_[2 * p + 1] = sinf(a0) * a1;
    }
    if (have_map_ && ctx()) ok(tdr_map_set_polar_table(ctx(), tab_.data(), n_theta, n_r));
  }
  int nTheta() const { return n_theta_; }
  int nR() const { return n_r_; }
};

// active_localizer.h:7-16 / active_localizer.cpp:45-82: the whole candidate search is one call
class ActiveLocalizer {
 public:
  explicit ActiveLocalizer(TopDownMapPolar* map) : map_(map) {}
  Vector2f getBestRelPos(std::vector<Vector3f>& preds) {
    Vector2f best;                                                          // (dist, theta); (0, 0) when nothing beats 0
    if (!ctx() || !map_ || !map_->haveMap() || preds.empty()) return best;
    float rel[2] = {0, 0}, diff = 0;
    static_assert(sizeof(Vector3f) == 12, "predictions are packed (x, y, theta)");
    if (ok(tdr_active_best_rel_pos(ctx(), &preds[0].x, (int)preds.size(), rel, &diff))) { best.x = rel[0]; best.y = rel[1]; }
    return best;
  }
 private:
  TopDownMapPolar* map_;
};

class ParticleFilter {
 public:
  // particle_filter.cpp:3-17.  seed: the reference seeds from std::random_device; a fixed seed makes runs repeatable.
  ParticleFilter(int N, TopDownMapPolar* map, FilterParams& params, uint32_t seed = std::random_device{}())
      : gen_(seed), map_(map), params_(params), max_num_particles_(N) {
    if (map_->haveMap()) initializeParticles();
  }
  // particle_filter.cpp:86-92 + StateParticle::propagate (state_particle.cpp:57-78).  The noise comes from the
  // reference's own RNG calls in particle order — per particle a fresh theta distribution drawn once, a fresh
  // displacement distribution drawn twice (the second draw is the polar method's saved value), a fresh scale
  // distribution unless frozen — taken as STANDARD variates (the uniforms consumed are the same for any stddev);
  // the states stay on the device and tdr_pf_propagate applies `z * stddev + mean` and the motion there.
  void propagate(const Vector2f& trans, float omega) {
    if (!ctx()) return;
    int64_t n = 0; tdr_pf_count(ctx(), &n);
    if (n <= 0) return;
    z_.resize((size_t)n * 4);
    for (int64_t i = 0; i < n; i++) {
      std::normal_distribution<float> theta_dist{0, 1}, disp_dist{0, 1};
      z_[4 * i] = theta_dist(gen_);
      z_[4 * i + 1] = disp_dist(gen_);
      z_[4 * i + 2] = disp_dist(gen_);
      z_[4 * i + 3] = 0.f;
      if (!scale_frozen_) { std::normal_distribution<float> scale_dist{0, 1}; z_[4 * i + 3] = scale_dist(gen_); }
    }
    if (ok(tdr_pf_propagate(ctx(), trans.x, trans.y, omega, scale_frozen_ ? 1 : 0, params_.pos_cov, params_.theta_cov, z_.data(), n)))
      host_dirty_ = true;
  }
  // particle_filter.cpp:262-272: the sample matrix the GMM thread fits (num_samples x 4 doubles), straight off the device
  std::vector<double> gmmSamples() {
    int64_t n = 0;
    if (!ctx() || tdr_pf_count(ctx(), &n) != 0 || n <= 0) return {};
    const int num_samples = (int)std::min<int64_t>(1000, n);
    std::vector<double> s((size_t)num_samples * 4);
    if (!ok(tdr_pf_gmm_samples(ctx(), num_samples, s.data()))) s.clear();
    return s;
  }
  // particle_filter.cpp:151-158: particle count of the next resampling from the GMM covariances (the top-left 2x2 block of
  // every row-major 3x3 Matrix3f): sum of sqrt(l0) * sqrt(l1) over the clusters, bounded below by 3/4 of the last count + 10
  using Matrix3f = std::array<float, 9>;
  static int adaptiveCount(const std::vector<Matrix3f>& covs, int last_num_particles, int max_num_particles) {
    int num = 0;
    for (const auto& cv : covs) {
      const float a = cv[0], b = cv[1], c = cv[3], d = cv[4];
      const float tr = a + d, det = a * d - b * c, disc = tr * tr - 4 * det;
      float e0, e1;
      if (disc >= 0) { const float sq = std::sqrt(disc); e0 = (tr - sq) / 2; e1 = (tr + sq) / 2; } else { e0 = e1 = tr / 2; }
      num += static_cast<int>(std::sqrt(e0) * std::sqrt(e1));
    }
    return std::min(std::max(num, 3 * last_num_particles / 4 + 10), max_num_particles);
  }
  // particle_filter.cpp:238-243.  The mixture itself is fitted by OpenCV's cv::ml::EM in the reference's GMM thread
  // (:252-318) on the matrix gmmSamples() returns; the adapter keeps that thread and hands the result over with setGMM.
  // Until a mixture has been set, update() keeps the particle count (the reference fits one before the first update).
  void setGMM(const std::vector<Vector3f>& means, const std::vector<Matrix3f>& covs) { means_ = means; covs_ = covs; }
  void getGMM(std::vector<Vector3f>& means, std::vector<Matrix3f>& covs) const { means = means_; covs = covs_; }
  // particle_filter.cpp:94-189.  top_down_geo is accepted and ignored, as in the reference's cost (F10).
  void update(std::vector<ArrayXXf>& top_down_scan, std::vector<ArrayXXf>& /*top_down_geo*/, float res) {
    if (num_particles_ == 0 || !ctx()) return;                             // :96-99
    const int C = (int)top_down_scan.size(), n_theta = top_down_scan[0].rows(), n_r = top_down_scan[0].cols();
    stage_.resize((size_t)C * n_theta * n_r);
    for (int c = 0; c < C; c++) std::copy(top_down_scan[c].d.begin(), top_down_scan[c].d.end(), stage_.begin() + (size_t)c * n_theta * n_r);
    if (!ok(tdr_scan_set_polar_images(ctx(), stage_.data(), n_theta, n_r, C))) return;
    std::uniform_real_distribution<float> dist(0., 1.);
    if (!covs_.empty()) num_particles_ = adaptiveCount(covs_, num_particles_, max_num_particles_);   // :151-158
    const float u = last_u_ = dist(gen_);                                  // the ONE draw of :172-173
    if (!ok(tdr_pf_update(ctx(), res, u, num_particles_))) return;
    host_dirty_ = true;
  }
  void meanLikelihood(float state[4]) { if (ctx()) ok(tdr_pf_pose(ctx(), state, nullptr, nullptr, nullptr)); }
  void computeMeanCov(float cov[16]) { float m[4]; if (ctx()) ok(tdr_pf_pose(ctx(), m, cov, nullptr, nullptr)); }
  void maxLikelihood(float state[4]) { if (ctx()) ok(tdr_pf_pose(ctx(), nullptr, nullptr, state, nullptr)); }
  void computeCov(float cov[16]) { float m[4]; if (ctx()) ok(tdr_pf_pose(ctx(), nullptr, nullptr, m, cov)); }
  int numParticles() const { return num_particles_; }
  // particle_filter.cpp:358-366
  float scale() {
    if (params_.fixed_scale > 0) return params_.fixed_scale;
    if (scale_frozen_) { pull(); return states_.empty() ? -1.f : states_[0].scale; }
    return -1;
  }
  bool isScaleFrozen() const { return scale_frozen_; }
  // particle_filter.cpp:343-357: lock every particle's scale to the geometric mean (float accumulator, double pow)
  void freezeScale() {
    if (scale_frozen_) return;
    pull();
    float geo_mean = 1;
    for (const State& s : states_) geo_mean *= std::pow(s.scale, 1. / states_.size());
    for (State& s : states_) s.scale = geo_mean;
    scale_frozen_ = true;
    push();
  }
  // particle_filter.cpp:320-341: new aerial map -> distance fields on the device, every particle's init position
  // shifted by the map-centre delta; particles are initialised if there were none yet
  void updateMap(const uint8_t* img, int rows, int cols, int stride, const Vector2i& map_center) {
    map_->updateMap(img, rows, cols, stride, map_center);
    const int dx = map_center.x - last_map_center_.x, dy = map_center.y - last_map_center_.y;
    if (num_particles_ > 0 && (dx != 0 || dy != 0)) {
      pull();
      for (State& s : states_) { s.init_x_px += dx; s.init_y_px += dy; }
      push();
    }
    last_map_center_ = map_center;
    if (num_particles_ == 0 && map_->haveMap()) initializeParticles();   // the reference does not test haveMap here and spins in the road search on a road-less map
  }
  // host views (the reference's visualize / GMM thread read the particles on the host)
  const std::vector<State>& states() { pull(); return states_; }
  const std::vector<float>& lastDist() { pull(); return last_dist_; }
  float lastUniform() const { return last_u_; }
  std::vector<float> weights() {
    std::vector<float> w(num_particles_);
    if (ctx() && num_particles_) ok(tdr_pf_get_weights(ctx(), w.data(), num_particles_));
    return w;
  }

 private:
  // StateParticle::StateParticle(gen, map, params, init = true)  state_particle.cpp:3-49
  State randomState() {
    std::uniform_real_distribution<float> uniform_dist(0., 1.);
    std::normal_distribution<float> normal_dist(0., 1.);
    State s{};
    const Vector2i sz = map_->size();
    const float mw = (float)sz.x * map_->resolution(), mh = (float)sz.y * map_->resolution();
    if (params_.fixed_scale < 0) s.scale = (float)std::pow(10, (uniform_dist(gen_) - 0.5) * 2);
    else s.scale = params_.fixed_scale;
    std::vector<int> cls_vec;
    while (true) {
      if (params_.init_pos_px_x > 0) {
        s.init_x_px = std::clamp<float>(normal_dist(gen_) * params_.init_pos_px_cov + params_.init_pos_px_x, 0, mw);
        s.init_y_px = std::clamp<float>(normal_dist(gen_) * params_.init_pos_px_cov + params_.init_pos_px_y, 0, mh);
      } else {
        s.init_x_px = uniform_dist(gen_) * mw;
        s.init_y_px = uniform_dist(gen_) * mh;
      }
      map_->getClassesAtPoint(Vector2i{(int)s.init_x_px, (int)s.init_y_px}, cls_vec);
      if (std::find(cls_vec.begin(), cls_vec.end(), 1) != cls_vec.end()) break;   // particle is on the road
    }
    if (params_.init_pos_deg_theta != std::numeric_limits<float>::infinity()) {
      s.theta = normal_dist(gen_) * params_.init_pos_deg_cov + params_.init_pos_deg_theta;
      s.theta *= M_PI / 180;
      s.have_init = 1;
    } else { s.theta = 0; s.have_init = 0; }
    return s;
  }
  // particle_filter.cpp:19-84 (the GMM thread and visualisation are out of scope)
  void initializeParticles() {
    size_t num_at_scale = 1;
    if (params_.fixed_scale < 0) num_at_scale = 10; else scale_frozen_ = true;
    // :27-54: a metric initial position (relative to the map centre) overrides the pixel one; the filter stays empty
    // when that position is off the map or has no road within 4 px
    if (scale_frozen_ && params_.init_pos_m_x != std::numeric_limits<float>::infinity()) {
      const Vector2i mc = map_->mapCenter(), sz = map_->size();
      params_.init_pos_px_x = (params_.init_pos_m_x * params_.fixed_scale) + mc.x;
      params_.init_pos_px_y = (params_.init_pos_m_y * params_.fixed_scale) + mc.y;
      if (params_.init_pos_px_x < 0 || params_.init_pos_px_x >= sz.x || params_.init_pos_px_y < 0 || params_.init_pos_px_y >= sz.y) {
        fprintf(stderr, "[XView] No map received for input loc\n");
        return;
      }
      bool good_init = false;
      std::vector<int> cls_vec;
      for (int dx = -4; dx <= 4 && !good_init; dx++)
        for (int dy = -4; dy <= 4 && !good_init; dy++) {
          map_->getClassesAtPoint(Vector2i{(int)(params_.init_pos_px_x + dx), (int)(params_.init_pos_px_y + dy)}, cls_vec);
          good_init = std::find(cls_vec.begin(), cls_vec.end(), 1) != cls_vec.end();
        }
      if (!good_init) {
        fprintf(stderr, "[XView] No road in map at init location\n");
        return;
      }
    }
    for (int i = 0; i < max_num_particles_ / (int)num_at_scale; i++) {
      State proto = randomState();
      for (float scale = 0; scale < 1; scale += 1. / num_at_scale) {
        State p = randomState();                                            // the reference constructs (and draws) again
        if (params_.fixed_scale < 0) { p = proto; p.scale = (float)std::pow(10., scale); }
        states_.push_back(p);
        (void)randomState();                                                // new_particles_ entry: same RNG consumption
      }
    }
    num_particles_ = (int)states_.size();
    last_dist_.assign(states_.size(), 0.f);
    tdr_filter_params fp{};
    fp.regularization = params_.regularization; fp.force_on_map = params_.force_on_map ? 1 : 0;
    fp.fixed_scale = params_.fixed_scale; fp.scale_log_min = params_.scale_log_min; fp.scale_log_max = params_.scale_log_max;
    fp.num_classes = map_->numClasses();
    for (int c = 0; c < fp.num_classes && c < 16; c++) fp.class_weights[c] = c < (int)params_.class_weights.size() ? params_.class_weights[c] : 1.f;
    if (!ctx() || !ok(tdr_pf_set_params(ctx(), &fp))) return;
    // theta-search candidates with the reference's own float loop (state_particle.cpp:197, :123-128)
    std::vector<float> thetas; std::vector<int32_t> shifts;
    const int nb = map_->nTheta();
    for (float t = 0; t < 2 * M_PI; t += 2 * M_PI / 40) {
      thetas.push_back(t);
      int s = (int)std::round((double)(t * nb / 2) / M_PI);
      while (s >= nb) s -= nb;
      while (s < 0) s += nb;
      shifts.push_back(s);
    }
    ok(tdr_pf_set_search(ctx(), thetas.data(), shifts.data(), (int)thetas.size()));
    push();
  }
  void push() { if (ctx() && !states_.empty()) ok(tdr_pf_set_states(ctx(), states_.data(), last_dist_.data(), (int64_t)states_.size())); host_dirty_ = false; }
  void pull() {
    if (!host_dirty_ || !ctx()) return;
    int64_t n = 0; tdr_pf_count(ctx(), &n);
    states_.resize((size_t)n); last_dist_.resize((size_t)n, 0.f);
    if (n) { ok(tdr_pf_get_states(ctx(), states_.data(), n)); ok(tdr_pf_get_last_dist(ctx(), last_dist_.data(), n)); }
    num_particles_ = (int)n; host_dirty_ = false;
  }

  std::mt19937 gen_;
  TopDownMapPolar* map_;
  FilterParams params_;
  int max_num_particles_ = 0, num_particles_ = 0;
  bool scale_frozen_ = false, host_dirty_ = false;
  Vector2i last_map_center_;
  float last_u_ = 0.f;
  std::vector<State> states_;
  std::vector<float> last_dist_, stage_, z_;
  std::vector<Vector3f> means_;
  std::vector<Matrix3f> covs_;
};

}  // namespace tdrhost
