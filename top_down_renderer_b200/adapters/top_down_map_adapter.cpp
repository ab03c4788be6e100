// adapters/top_down_map_adapter.cpp — the bodies that REPLACE src/top_down_map.cpp and src/top_down_map_polar.cpp of the reference.
// The class declarations are the reference's own, unchanged (include/top_down_render/top_down_map.h:52-101,
// top_down_map_polar.h:6-22); rasterising the vector map, the distance fields and every gather run on the device through
// the C ABI (include/tdr.h).  What stays host code is what is file / glue work in the reference too: the svg parse
// (nanosvg, vendored in the reference's include directory), the two cache formats (through the reference's own
// write_binary / read_binary templates and cv::imwrite / cv::imread), and host copies of the distance fields in
// class_maps_ / class_mask_ / geo_maps_ — the members the header declares — for getClassesAtPoint, size() and the caches.
//
// One device context holds ONE map; a process with several TopDownMap objects gets the right one re-installed from its
// host copies on demand (`make_resident`).  In this repository the file is compiled against the stand-in headers of
// oracle/ref_shim (`make -C oracle _adapters`) and run by tests/test_adapters.py beside the reference's own bodies.
#include <cstring>
#include <filesystem>
#include <vector>

#include "top_down_render/top_down_map_polar.h"

#define NANOSVG_IMPLEMENTATION
#include "top_down_render/nanosvg.h"

#include "tdr_adapter_common.h"

using tdr_adapter::ok;
using tdr_adapter::tdr;

namespace {
const void* g_resident = nullptr;        // the TopDownMap whose layers the device holds
const void* g_fresh = nullptr;           // ... and for which it has JUST run the transform (no second upload in computeDists)
std::vector<float> g_table;              // the polar offset table the device holds (n_theta * n_r leading entries)
int g_table_theta = 0, g_table_r = 0;

std::string cache_dir() { return std::string(getenv("HOME")) + "/.ros/xview_cache/"; }

std::vector<float> pack(const std::vector<Eigen::ArrayXXf>& maps) {
  std::vector<float> all;
  for (const auto& m : maps) all.insert(all.end(), m.data(), m.data() + m.size());
  return all;
}
void unpack(const std::vector<float>& all, int rows, int cols, size_t count, std::vector<Eigen::ArrayXXf>& maps) {
  maps.clear();
  for (size_t k = 0; k < count; k++) {
    Eigen::ArrayXXf m(rows, cols);
    std::memcpy(m.data(), all.data() + k * (size_t)rows * cols, (size_t)rows * cols * sizeof(float));
    maps.push_back(m);
  }
}
}  // namespace

// ---- construction (replaces :9-64) ----------------------------------------------------------------------------------------
TopDownMap::TopDownMap(const TopDownMap::Params& params) {
  params_ = params;
  map_center_ = Eigen::Vector2i::Zero();
  have_map_ = false;
  if (g_resident == this) g_resident = nullptr;                             // an earlier object lived at this address
  if (g_fresh == this) g_fresh = nullptr;
  if (params_.map_path == "" || !tdr()) return;                             // dynamic map: updateMap brings it
  if (loadCacheMetaData(params_.map_path)) {
    loadCachedMaps();                                                        // distance fields straight to the device, no transform
  } else {
    const std::string ext = params_.map_path.size() >= 4 ? params_.map_path.substr(params_.map_path.size() - 4) : "";
    if (ext == ".svg") {
      std::vector<std::vector<std::vector<Eigen::Vector2f>>> poly;
      Eigen::Vector2i map_size = loadSvg(params_.map_path, poly);
      getRasterMap(map_size, 0, params_.resolution, poly);                   // polygons -> binary class maps, on the device
      saveRasterizedMaps(params_.map_path.substr(0, params_.map_path.size() - 4) + "_raster_cache");
    } else if (ext == ".png" || ext == ".jpg") {
      cv::Mat color_map = cv::imread(params_.map_path);
      if (color_map.empty()) { ROS_ERROR("[XView] Compressed raster map loading failed"); return; }
      cv::Mat class_map;
      params_.color_lut.color2Ind(color_map, class_map);
      loadCompressedRasterMap(class_map);
    } else {
      loadRasterizedMaps(params_.map_path);
    }
    if (class_maps_.empty()) return;
    Eigen::ArrayXXc unused;
    computeDists(class_maps_, class_mask_);                                  // seeds -> distance fields (and the geo pair) on the device
    computeDists(geo_maps_, unused);
    saveCachedMaps(params_.map_path);
  }
  have_map_ = true;
}

// replaces :66-114.  Same polygons in the same order: per class index the shapes whose fill colour matches, every third
// point of each cubic path (the anchors), y measured from the bottom.
Eigen::Vector2i TopDownMap::loadSvg(const std::string& svg_path, std::vector<std::vector<std::vector<Eigen::Vector2f>>>& poly) {
  NSVGimage* image = nsvgParseFromFile(svg_path.c_str(), "px", 96);
  if (image == NULL) { ROS_ERROR("[XView] Svg map loading failed"); return Eigen::Vector2i(0, 0); }
  poly.resize(params_.num_classes);
  for (size_t cls = 0; cls < params_.flatten_lut.size(); cls++) {
    const auto rgb = SemanticColorLut::unpackColor(params_.color_lut.ind2Color(cls));
    const uint32_t wanted = (uint32_t)rgb[0] << 16 | (uint32_t)rgb[1] << 8 | (uint32_t)rgb[2];   // nanosvg packs 0xBBGGRR
    auto& dst = poly[params_.flatten_lut[cls]];
    for (NSVGshape* shape = image->shapes; shape != NULL; shape = shape->next) {
      if ((shape->fill.color & 0xFFFFFF) != wanted) continue;
      for (NSVGpath* path = shape->paths; path != NULL; path = path->next) {
        std::vector<Eigen::Vector2f> ring;
        for (int i = 0; i < path->npts - 1; i += 3) ring.push_back(Eigen::Vector2f(path->pts[2 * i], image->height - path->pts[2 * i + 1]));
        dst.push_back(ring);
      }
    }
  }
  Eigen::Vector2i size((int)image->width, (int)image->height);
  nsvgDelete(image);
  return size;
}

// replaces :391-408 together with samplePts (:367-389) and getClasses (:328-365): tdr_map_set_polygons
void TopDownMap::getRasterMap(const Eigen::Vector2i& map_size, float rot, float res, std::vector<std::vector<std::vector<Eigen::Vector2f>>>& poly) {
  tdr_adapter::Guard dev_guard;
  if (poly.size() < 1 || !tdr()) return;
  std::vector<float> verts;
  std::vector<int32_t> start(1, 0), cls;
  for (size_t c = 0; c < poly.size(); c++)
    for (const auto& ring : poly[c]) {
      for (const auto& v : ring) { verts.push_back(v[0]); verts.push_back(v[1]); }
      start.push_back((int32_t)(verts.size() / 2));
      cls.push_back((int32_t)c);
    }
  const int rows = static_cast<int>(map_size[1] / params_.resolution), cols = static_cast<int>(map_size[0] / params_.resolution);
  std::vector<float> layers((size_t)params_.num_classes * rows * cols);
  std::vector<int32_t> excl(params_.exclusive_classes.begin(), params_.exclusive_classes.end());
  if (!ok(tdr_map_set_polygons(tdr(), verts.data(), start.data(), cls.data(), (int)cls.size(), map_size[0], map_size[1], rot,
                               params_.num_classes, res, excl.data(), (int)excl.size(), layers.data()))) return;
  unpack(layers, rows, cols, params_.num_classes, class_maps_);              // the BINARY maps, what saveRasterizedMaps writes
  g_resident = g_fresh = this;
}

// replaces :289-326 (and :410-427 for the geometric pair): the transform runs on the device; `classes` comes back as
// distance fields exactly as the reference leaves them.  class_maps_ -> upload the seeds; geo_maps_ -> the device derives
// the pair from the class seeds it already holds.
void TopDownMap::computeDists(std::vector<Eigen::ArrayXXf>& classes, Eigen::ArrayXXc& mask) {
  tdr_adapter::Guard dev_guard;
  if (!tdr() || class_maps_.empty()) return;
  const int rows = (int)class_maps_[0].rows(), cols = (int)class_maps_[0].cols();
  if (&classes == &class_maps_) {
    if (g_fresh != this) {                                                   // e.g. maps read from the raster cache: seeds go up first
      std::vector<float> seeds = pack(class_maps_);
      if (!ok(tdr_map_set_binary_layers(tdr(), seeds.data(), rows, cols, (int)class_maps_.size(), params_.resolution))) return;
    }
    g_resident = this;
    g_fresh = nullptr;
    g_table_theta = 0;                                                       // a new map: the table is set again on the next gather
    std::vector<float> dist(class_maps_.size() * (size_t)rows * cols);
    mask = Eigen::ArrayXXc(rows, cols);
    if (!ok(tdr_map_get_layers(tdr(), dist.data(), mask.data()))) return;
    unpack(dist, rows, cols, class_maps_.size(), class_maps_);
  } else {
    std::vector<float> geo((size_t)2 * rows * cols);
    if (g_resident != this || !ok(tdr_map_get_geo_layers(tdr(), geo.data()))) return;
    unpack(geo, rows, cols, 2, classes);
  }
}
void TopDownMap::getGeoRasterMap(std::vector<Eigen::ArrayXXf>&) {}           // folded into computeDists(geo_maps_) above

// replaces :116-144 for the static .png / .jpg path: class-index image -> device; class_maps_ receives the binary maps
void TopDownMap::loadCompressedRasterMap(const cv::Mat& map) {
  if (!tdr() || map.empty()) return;
  std::vector<int32_t> lut(256, -1);
  std::copy(params_.flatten_lut.begin(), params_.flatten_lut.end(), lut.begin());
  if (!ok(tdr_map_set_class_image(tdr(), map.data, map.rows, map.cols, (int)map.step, lut.data(), 256, params_.num_classes, params_.resolution))) return;
  g_resident = g_fresh = this;
  int rows = 0, cols = 0, k = 0; float r = 1;
  tdr_map_info(tdr(), &rows, &cols, &k, &r);
  class_maps_.clear();
  for (int c = 0; c < k; c++) class_maps_.push_back(Eigen::ArrayXXf::Constant(rows, cols, 1.0f));
  for (int xi = 0; xi < cols; xi++)                                          // the binary maps on the host too (have_map_ test, raster cache)
    for (int yi = 0; yi < rows; yi++) {
      const int cls = lut[map.at<uint8_t>(std::max<int>(map.rows - yi * params_.resolution - 1, 0), std::min<int>(xi * params_.resolution, map.cols - 1))];
      if (cls >= 0 && cls < k) class_maps_[cls](yi, xi) = 0;
    }
}

// replaces :146-157
void TopDownMap::updateMap(const cv::Mat& map, const Eigen::Vector2i& map_center) {
  map_center_ = map_center;
  loadCompressedRasterMap(map);
  if (class_maps_.size() < 2) return;
  if (!class_maps_[1].isZero(0)) have_map_ = true;                          // the reference's test, on the binary layer (:150)
  else ROS_WARN("[XView] Received map with no road");
  computeDists(class_maps_, class_mask_);
}

// ---- caches -----------------------------------------------------------------------------------------------------------------
// replaces :197-211
void TopDownMap::saveRasterizedMaps(const std::string& path) {
  mkdir(path.c_str(), S_IRWXU);
  size_t ind = 0;
  for (const auto& map : class_maps_) {
    cv::Mat img((int)map.rows(), (int)map.cols(), CV_8UC1);
    for (int r = 0; r < img.rows; r++)
      for (int c = 0; c < img.cols; c++) img.at<uint8_t>(img.rows - 1 - r, c) = map(r, c) >= 0.5f ? 255 : 0;   // x255, flipped to look like the input
    cv::imwrite(path + "/class" + std::to_string(ind++) + ".png", img);
  }
}
// replaces :213-224
void TopDownMap::loadRasterizedMaps(const std::string& map_path) {
  tdr_adapter::Guard dev_guard;
  for (size_t i = 0; i < (size_t)params_.num_classes; i++) {
    cv::Mat img = cv::imread(map_path + "/class" + std::to_string(i) + ".png", cv::IMREAD_GRAYSCALE);
    if (img.empty()) { class_maps_.clear(); return; }
    Eigen::ArrayXXf map(img.rows, img.cols);
    for (int r = 0; r < img.rows; r++)
      for (int c = 0; c < img.cols; c++) map(r, c) = (float)img.at<uint8_t>(img.rows - 1 - r, c) * (float)(1. / 255);
    class_maps_.push_back(map);
  }
}
// replaces :226-242
bool TopDownMap::loadCacheMetaData(const std::string& map_path) {
  std::ifstream data_file(cache_dir() + "cached_data.txt");
  if (!data_file) return false;
  std::string line;
  std::getline(data_file, line);
  if (line != map_path) return false;
  std::getline(data_file, line);
  if (std::stoi(line) != params_.num_classes) return false;
  std::getline(data_file, line);
  return std::abs(std::stof(line) - params_.resolution) <= 0.01;
}
// replaces :244-261: the reference's own read_binary template (top_down_map.h:40-50), then the fields go to the device as they are
void TopDownMap::loadCachedMaps() {
  tdr_adapter::Guard dev_guard;
  for (int cls = 0; cls < params_.num_classes; cls++) {
    Eigen::ArrayXXf m;
    std::string name = cache_dir() + "class_map" + std::to_string(cls) + ".eig";
    read_binary(name, m);
    class_maps_.push_back(m);
  }
  for (int cls = 0; cls < 2; cls++) {
    Eigen::ArrayXXf m;
    std::string name = cache_dir() + "geo_map" + std::to_string(cls) + ".eig";
    read_binary(name, m);
    geo_maps_.push_back(m);
  }
  std::string name = cache_dir() + "class_mask.eig";
  read_binary(name, class_mask_);
  const int rows = (int)class_maps_[0].rows(), cols = (int)class_maps_[0].cols();
  if (!ok(tdr_map_set_dist_layers(tdr(), pack(class_maps_).data(), class_mask_.data(), rows, cols, params_.num_classes, params_.resolution))) return;
  ok(tdr_map_set_geo_dist_layers(tdr(), pack(geo_maps_).data()));
  g_resident = this;
  g_table_theta = 0;
}
// replaces :263-286
void TopDownMap::saveCachedMaps(const std::string& map_path) {
  std::filesystem::create_directory(cache_dir());
  std::ofstream data_file(cache_dir() + "cached_data.txt", std::ofstream::out | std::ofstream::trunc);
  data_file << map_path << std::endl << params_.num_classes << std::endl << params_.resolution << std::endl;
  for (int cls = 0; cls < params_.num_classes; cls++) {
    std::string name = cache_dir() + "class_map" + std::to_string(cls) + ".eig";
    write_binary(name, class_maps_[cls]);
  }
  for (int cls = 0; cls < 2 && cls < (int)geo_maps_.size(); cls++) {
    std::string name = cache_dir() + "geo_map" + std::to_string(cls) + ".eig";
    write_binary(name, geo_maps_[cls]);
  }
  std::string name = cache_dir() + "class_mask.eig";
  write_binary(name, class_mask_);
}

// ---- queries ----------------------------------------------------------------------------------------------------------------
// :159-175: host copies of the distance fields answer
void TopDownMap::getClassesAtPoint(const Eigen::Vector2i& center_ind, std::vector<int>& classes) {
  const int cx = (int)((float)center_ind[0] / params_.resolution), cy = (int)((float)center_ind[1] / params_.resolution);
  classes.clear();
  for (int cls = 0; cls < params_.num_classes; cls++)
    if (cx < class_maps_[cls].cols() && cy < class_maps_[cls].rows() && cx >= 0 && cy >= 0 && class_maps_[cls](cy, cx) < 1) classes.push_back(cls);
}
void TopDownMap::getClassesAtPoint(const Eigen::Vector2f& center, std::vector<int>& classes) {
  getClassesAtPoint(Eigen::Vector2i((int)(center[0] / params_.resolution), (int)(center[1] / params_.resolution)), classes);
}
int TopDownMap::numClasses() const { return params_.num_classes; }
Eigen::Vector2i TopDownMap::size() const { return Eigen::Vector2i(class_maps_[0].cols(), class_maps_[0].rows()); }
Eigen::Vector2i TopDownMap::mapCenter() const { return map_center_; }
float TopDownMap::resolution() const { return params_.resolution; }
bool TopDownMap::haveMap() const { return have_map_; }

namespace {
// the device holds one map: re-install this object's distance fields from its host copies when another map was loaded since
bool make_resident(TopDownMap* m, const std::vector<Eigen::ArrayXXf>& class_maps, const Eigen::ArrayXXc& mask,
                   const std::vector<Eigen::ArrayXXf>& geo, int num_classes, float resolution) {
  if (!tdr() || class_maps.empty()) return false;
  if (g_resident == m) return true;
  const int rows = (int)class_maps[0].rows(), cols = (int)class_maps[0].cols();
  if (!ok(tdr_map_set_dist_layers(tdr(), pack(class_maps).data(), mask.data(), rows, cols, num_classes, resolution))) return false;
  if (geo.size() == 2) ok(tdr_map_set_geo_dist_layers(tdr(), pack(geo).data()));
  g_resident = m;
  g_table_theta = 0;
  return true;
}
void scatter(const std::vector<float>& stage, std::vector<Eigen::ArrayXXf>& dists, size_t count) {
  size_t at = 0;
  for (size_t k = 0; k < count && k < dists.size(); k++) { std::memcpy(dists[k].data(), stage.data() + at, (size_t)dists[k].size() * 4); at += (size_t)dists[k].size(); }
}
}  // namespace

// replaces :429-459 (a8)
void TopDownMap::getLocalMap(Eigen::Vector2f center, float rot, float res, std::vector<Eigen::ArrayXXf>& dists, Eigen::ArrayXXc& mask) {
  tdr_adapter::Guard dev_guard;
  if (dists.size() < 1 || !make_resident(this, class_maps_, class_mask_, geo_maps_, params_.num_classes, params_.resolution)) return;
  const int rows = (int)dists[0].rows(), cols = (int)dists[0].cols();
  std::vector<float> stage((size_t)params_.num_classes * rows * cols);
  if (!ok(tdr_map_local_cart(tdr(), center[0], center[1], rot, res, rows, cols, stage.data(), mask.data()))) return;
  scatter(stage, dists, params_.num_classes);
}

// ---- TopDownMapPolar (replaces top_down_map_polar.cpp) -----------------------------------------------------------------------
TopDownMapPolar::TopDownMapPolar(const Params& params) : TopDownMap(params) { samplePtsPolar(Eigen::Vector2i(100, 50), 2 * M_PI / 100); }

// :7-19 (with samplePts at rot = 0): offset p = (row p % n_theta, column p / n_theta) -> (cos, sin)(angle) * radius
void TopDownMapPolar::samplePtsPolar(Eigen::Vector2i shape, float ang_res) {
  tdr_adapter::Guard dev_guard;
  const int n_theta = shape[0], n_r = shape[1];
  ang_sample_pts_ = Eigen::Array2Xf(2, n_theta * n_r);
  for (int p = 0; p < n_theta * n_r; p++) {
    const float angle = ((float)(p % n_theta) - (float)(n_theta - 1) / 2.f) * ang_res;
    const float radius = (float)(p / n_theta) * (float)(1. / params_.resolution);
    ang_sample_pts_(0, p) = cosf(angle) * radius;
    ang_sample_pts_(1, p) = sinf(angle) * radius;
  }
}

namespace {
// the caller's images define the raster (n_theta x n_r); the reference reads the first n_theta * n_r offsets of its table
// (top_down_map_polar.cpp:33), which the constructor sizes 100 x 50.  Uploaded when it differs from what the device holds.
bool table_resident(const Eigen::Array2Xf& tab, int n_theta, int n_r) {
  const size_t n = (size_t)2 * n_theta * n_r;
  if ((size_t)tab.size() < n) return false;
  if (g_table_theta == n_theta && g_table_r == n_r && g_table.size() == n && std::memcmp(g_table.data(), tab.data(), n * 4) == 0) return true;
  if (!ok(tdr_map_set_polar_table(tdr(), tab.data(), n_theta, n_r))) return false;
  g_table.assign(tab.data(), tab.data() + n);
  g_table_theta = n_theta; g_table_r = n_r;
  return true;
}
}  // namespace

// replaces :21-53 (a7)
void TopDownMapPolar::getLocalMap(Eigen::Vector2f center, float scale, float res, std::vector<Eigen::ArrayXXf>& dists, Eigen::ArrayXXc& mask) {
  tdr_adapter::Guard dev_guard;
  if (dists.size() < 1 || !make_resident(this, class_maps_, class_mask_, geo_maps_, params_.num_classes, params_.resolution)) return;
  if (!table_resident(ang_sample_pts_, (int)dists[0].rows(), (int)dists[0].cols())) return;
  const float c[2] = {center[0], center[1]};
  std::vector<float> stage((size_t)params_.num_classes * dists[0].size());
  if (!ok(tdr_map_local_polar(tdr(), c, 1, scale, res, stage.data(), mask.data()))) return;
  scatter(stage, dists, params_.num_classes);
}
// replaces :55-76
void TopDownMapPolar::getLocalGeoMap(Eigen::Vector2f center, float scale, float res, std::vector<Eigen::ArrayXXf>& dists) {
  tdr_adapter::Guard dev_guard;
  if (dists.size() < 1 || !make_resident(this, class_maps_, class_mask_, geo_maps_, params_.num_classes, params_.resolution)) return;
  if (!table_resident(ang_sample_pts_, (int)dists[0].rows(), (int)dists[0].cols())) return;
  const float c[2] = {center[0], center[1]};
  std::vector<float> stage((size_t)2 * dists[0].size());
  if (!ok(tdr_map_local_geo_polar(tdr(), c, 1, scale, res, stage.data()))) return;
  scatter(stage, dists, 2);
}
void TopDownMapPolar::getLocalMap(Eigen::Vector2f center, float res, std::vector<Eigen::ArrayXXf>& dists, Eigen::ArrayXXc& mask) {
  tdr_adapter::Guard dev_guard;
  getLocalMap(center, 1, res, dists, mask);                                  // :78-82
}
void TopDownMapPolar::getLocalGeoMap(Eigen::Vector2f center, float res, std::vector<Eigen::ArrayXXf>& dists) {
  tdr_adapter::Guard dev_guard;
  getLocalGeoMap(center, 1, res, dists);                                     // :84-87
}
