// adapters/scan_renderer_adapter.cpp — the bodies that REPLACE src/scan_renderer.cpp and src/scan_renderer_polar.cpp of the
// reference: the class declarations come from the reference's own, unchanged headers
// (include/top_down_render/scan_renderer.h:14-23, scan_renderer_polar.h:15-22); every member with arithmetic calls the
// C ABI of libtdr_b200 (include/tdr.h).  Compile inside the reference's catkin package in place of the two files, or —
// in this repository, where ROS / Eigen / PCL are absent — against the stand-in headers of oracle/ref_shim
// (`make -C oracle _adapters`), which is how tests/test_adapters.py links and runs it.
//
// renderGeometricTopDown (scan_renderer.cpp:7-53, scan_renderer_polar.cpp:6-81) runs on the device too
// (tdr_scan_render_geometric_polar / _cart): per angular bin a sort by range and the slope walk, per scan line the
// slope walk with line drawing.  Its only call in the node is commented out (top_down_render.cpp:540).
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "top_down_render/scan_renderer_polar.h"

#include "tdr_adapter_common.h"

using tdr_adapter::ok;
using tdr_adapter::tdr;

namespace {
// scan upload shared by both renderers: pcl::PointXYZI is 32 bytes with the intensity at byte 16
bool upload(const Eigen::VectorXi& lut, const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud, int num_images) {
  static_assert(sizeof(pcl::PointXYZI) == 32, "pcl::PointXYZI layout");
  if (!tdr()) return false;
  if (!ok(tdr_scan_set_lut(tdr(), lut.data(), (int)lut.size(), num_images))) return false;
  return ok(tdr_scan_set_points(tdr(), cloud->points.data(), (int)sizeof(pcl::PointXYZI), (int)offsetof(pcl::PointXYZI, intensity),
                                (int64_t)cloud->height * cloud->width));
}
// the geometric renderers need the points only (no class lut)
bool upload_points(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud) {
  if (!tdr()) return false;
  return ok(tdr_scan_set_points(tdr(), cloud->points.data(), (int)sizeof(pcl::PointXYZI), (int)offsetof(pcl::PointXYZI, intensity),
                                (int64_t)cloud->height * cloud->width));
}
void scatter(const std::vector<float>& stage, std::vector<Eigen::ArrayXXf>& imgs) {
  size_t at = 0;
  for (auto& img : imgs) { std::memcpy(img.data(), stage.data() + at, (size_t)img.size() * sizeof(float)); at += (size_t)img.size(); }
}
}  // namespace

ScanRenderer::ScanRenderer(const Eigen::VectorXi& flatten_lut) { flatten_lut_ = flatten_lut; }   // scan_renderer.cpp:3-5

// replaces scan_renderer.cpp:55-78
void ScanRenderer::renderSemanticTopDown(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud, float res,
                                         std::vector<Eigen::ArrayXXf>& imgs) {
  tdr_adapter::Guard dev_guard;
  if (imgs.size() < 1) return;                                                           // :57
  const int rows = (int)imgs[0].rows(), cols = (int)imgs[0].cols();
  if (!upload(flatten_lut_, cloud, (int)imgs.size())) return;
  std::vector<float> stage(imgs.size() * (size_t)rows * cols);
  if (!ok(tdr_scan_render_cart(tdr(), res, rows, cols, stage.data()))) return;
  scatter(stage, imgs);
}

ScanRendererPolar::ScanRendererPolar(const Eigen::VectorXi& flatten_lut) : ScanRenderer(flatten_lut) {}   // scan_renderer_polar.cpp:3-4

// replaces scan_renderer_polar.cpp:83-109; the class images also stay resident on the device for ParticleFilter::update
void ScanRendererPolar::renderSemanticTopDown(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud, float res, float ang_res,
                                              std::vector<Eigen::ArrayXXf>& imgs) {
  tdr_adapter::Guard dev_guard;
  if (imgs.size() < 1) return;                                                           // :85
  const int n_theta = (int)imgs[0].rows(), n_r = (int)imgs[0].cols();
  if (!upload(flatten_lut_, cloud, (int)imgs.size())) return;
  std::vector<float> stage(imgs.size() * (size_t)n_theta * n_r);
  if (!ok(tdr_scan_render_polar(tdr(), res, ang_res, n_theta, n_r, stage.data()))) return;
  scatter(stage, imgs);
}

// replaces scan_renderer.cpp:7-53
void ScanRenderer::renderGeometricTopDown(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud, float res,
                                          std::vector<Eigen::ArrayXXf>& imgs) {
  tdr_adapter::Guard dev_guard;
  if (imgs.size() < 2) return;                                                           // :9
  const int rows = (int)imgs[0].rows(), cols = (int)imgs[0].cols();
  if (!upload_points(cloud)) return;
  std::vector<float> stage(2 * (size_t)rows * cols);
  if (!ok(tdr_scan_render_geometric_cart(tdr(), (int)cloud->width, (int)cloud->height, res, rows, cols, stage.data()))) return;
  for (size_t i = 2; i < imgs.size(); i++) imgs[i].setZero();                            // :12-14 zero every image
  std::vector<Eigen::ArrayXXf> two(imgs.begin(), imgs.begin() + 2);
  scatter(stage, two);
  imgs[0] = two[0]; imgs[1] = two[1];
}

// replaces scan_renderer_polar.cpp:6-81
void ScanRendererPolar::renderGeometricTopDown(const pcl::PointCloud<pcl::PointXYZI>::ConstPtr& cloud, float res, float ang_res,
                                               std::vector<Eigen::ArrayXXf>& imgs) {
  tdr_adapter::Guard dev_guard;
  if (imgs.size() < 2) return;                                                           // :8
  const int n_theta = (int)imgs[0].rows(), n_r = (int)imgs[0].cols();
  if (!upload_points(cloud)) return;
  std::vector<float> stage(2 * (size_t)n_theta * n_r);
  if (!ok(tdr_scan_render_geometric_polar(tdr(), (int)cloud->width, (int)cloud->height, res, ang_res, n_theta, n_r, stage.data()))) return;
  for (size_t i = 2; i < imgs.size(); i++) imgs[i].setZero();
  std::vector<Eigen::ArrayXXf> two(imgs.begin(), imgs.begin() + 2);
  scatter(stage, two);
  imgs[0] = two[0]; imgs[1] = two[1];
}
