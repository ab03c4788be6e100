// adapter_harness.cpp — TEST INFRASTRUCTURE ONLY: the same C interface as ref_harness.cpp, over the ADAPTERS
// (top_down_renderer_b200/adapters/*.cpp: the reference's unchanged class declarations, bodies over the C ABI) instead
// of the reference's own bodies.  Linked twice by `make -C oracle _adapters`: with the CPU stand-in of the C ABI
// (tests/cpp/tdr_cpu_standin.cpp, the oracle answers) and with libtdr_b200.so (the device answers).
#include "top_down_render/scan_renderer_polar.h"

#define ADP_API extern "C" __attribute__((visibility("default")))

static pcl::PointCloud<pcl::PointXYZI>::ConstPtr make_cloud(const float* aos, long n) {
  auto c = std::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
  c->points.resize((size_t)n);
  std::memcpy(c->points.data(), aos, (size_t)n * 32);
  c->width = (uint32_t)n; c->height = 1;
  return c;
}
static Eigen::VectorXi make_lut(const int* lut, int n) { Eigen::VectorXi v(n); for (int i = 0; i < n; i++) v[i] = lut[i]; return v; }

ADP_API void adp_render_polar(const float* pts, long n, float res, float ang_res, int n_theta, int n_r, const int* lut, int n_lut,
                              int C, float* imgs) {
  ScanRendererPolar r(make_lut(lut, n_lut));
  std::vector<Eigen::ArrayXXf> v(C, Eigen::ArrayXXf(n_theta, n_r));
  for (auto& a : v) a.setConstant(-7.f);                      // must be overwritten, not accumulated into
  r.renderSemanticTopDown(make_cloud(pts, n), res, ang_res, v);
  for (int c = 0; c < C; c++) std::memcpy(imgs + (size_t)c * n_theta * n_r, v[c].data(), (size_t)n_theta * n_r * 4);
}
ADP_API void adp_render_cart(const float* pts, long n, float res, int rows, int cols, const int* lut, int n_lut, int C, float* imgs) {
  ScanRenderer r(make_lut(lut, n_lut));
  std::vector<Eigen::ArrayXXf> v(C, Eigen::ArrayXXf(rows, cols));
  for (auto& a : v) a.setConstant(-7.f);
  r.renderSemanticTopDown(make_cloud(pts, n), res, v);
  for (int c = 0; c < C; c++) std::memcpy(imgs + (size_t)c * rows * cols, v[c].data(), (size_t)rows * cols * 4);
}
