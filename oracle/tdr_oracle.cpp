// =============================================================================
// tdr_oracle.cpp — CPU restatement of the top_down_render localization hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
// top_down_renderer_b200/ or libtdr_b200.so) may include, link or call this
// file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / CPU baseline.
//
// WHAT PINS IT.  The reference (KumarRobotics/top_down_renderer) ships no tests, golden vectors or fixtures, and it
// cannot be built as shipped in this image (every header needs ROS, Eigen, OpenCV C++, PCL — none installed, no
// network).  Two kinds of pins exist:
//   * PINNED AGAINST THE REFERENCE'S OWN SOURCE (oracle/_ref, tests/test_ref_build.py): the seven translation units of
//     the hot path (scan_renderer.cpp, scan_renderer_polar.cpp, top_down_map.cpp, top_down_map_polar.cpp,
//     state_particle.cpp, particle_filter.cpp, active_localizer.cpp) compile unmodified against stand-in headers
//     (oracle/ref_shim/) and run beside this file: every row of SURVEY 8a and the widened rows agree bit for bit, except
//     where a value passes through an Eigen reduction (weights: 1e-6 relative; the stand-in sums sequentially).
//   * NOT PINNED BY THAT BUILD — whatever is decided inside the absent third-party libraries — and pinned instead by
//       - cv2.distanceTransform / cv2.threshold (the real routine the reference calls, top_down_map.cpp:312,315) and
//         cv2.imwrite / imread for the raster cache — tests/test_oracle.py, tests/test_host_math.py,
//       - glibc 2.39 libm (atan2f/sqrtf/roundf — the reference's own libm calls),
//       - a deliberately naive numpy twin (oracle/numpy_twin.py) and hand KATs;
//     Eigen's SIMD reduction order, its packet cos/sin and the assertion-free meaning of samplePts' shape-mismatched
//     assignment are restated from Eigen's published sources and stay unpinned.
// Build contract being restated: g++ -O2, no -march (x86-64 baseline, SSE2
// 4-float Eigen packets, no FMA, no SSE3 hadd), -ffp-contract=off, Eigen
// 3.3/3.4 reduction orders where an order must be chosen.
//
// Citations are file:line relative to the reference checkout.
// Conventions: Eigen::ArrayXXf(rows, cols) is column-major, element (r, c) at
// c*rows + r.  Images/layers here are plain float arrays in that layout.
// =============================================================================
#include <math.h>   // global float overloads, as in the reference TU (ros.h pulls <math.h>)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <climits>
#include <vector>
#include <algorithm>
#include <thread>
#include <random>
#include <limits>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

// std::mt19937 that counts its 32-bit outputs, so that a later call can resume the reference's ONE shared engine
// (particle_filter.cpp:5) where an earlier one left it: same min / max, hence the same values out of every distribution
struct CountingMt {
  using result_type = std::mt19937::result_type;
  std::mt19937 g;
  uint64_t n = 0;
  explicit CountingMt(uint32_t seed, uint64_t discard = 0) : g(seed) { g.discard(discard); }
  static constexpr result_type min() { return std::mt19937::min(); }
  static constexpr result_type max() { return std::mt19937::max(); }
  result_type operator()() { ++n; return g(); }
};

// x86-64 cvttss2si semantics: NaN / out-of-range -> INT_MIN ("integer indefinite").
// The reference relies on this implicitly (float -> int assignments).
inline int f2i_x86(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)v;
}
inline int d2i_x86(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
  return (int)v;
}

// Eigen SSE2 predux<Packet4f>: (a0+a2)+(a1+a3)  (no SSE3 hadd in a flag-less build)
inline float predux4(const float p[4]) { return (p[0] + p[2]) + (p[1] + p[3]); }

// Eigen redux_impl<LinearVectorizedTraversal, NoUnrolling> for a sum over a
// 16-byte-aligned contiguous float vector (Eigen/src/Core/Redux.h).  Used by
// weights_.sum() (particle_filter.cpp:135,142).
float eigen_linear_sum(const float* x, long size) {
  if (size == 0) return 0.f;  // DenseBase::sum() special-cases empty
  const long ps = 4;
  const long alignedSize2 = (size / (2 * ps)) * (2 * ps);
  const long alignedSize = (size / ps) * ps;
  float res;
  if (alignedSize) {
    float p0[4] = {x[0], x[1], x[2], x[3]};
    if (alignedSize > ps) {
      float p1[4] = {x[4], x[5], x[6], x[7]};
      for (long i = 2 * ps; i < alignedSize2; i += 2 * ps) {
        for (int k = 0; k < 4; k++) p0[k] = p0[k] + x[i + k];
        for (int k = 0; k < 4; k++) p1[k] = p1[k] + x[i + ps + k];
      }
      for (int k = 0; k < 4; k++) p0[k] = p0[k] + p1[k];
      if (alignedSize > alignedSize2)
        for (int k = 0; k < 4; k++) p0[k] = p0[k] + x[alignedSize2 + k];
    }
    res = predux4(p0);
    for (long i = alignedSize; i < size; i++) res = res + x[i];
  } else {
    res = x[0];
    for (long i = 1; i < size; i++) res = res + x[i];
  }
  return res;
}

// Eigen redux_impl<SliceVectorizedTraversal> for (A.block * B.block).sum() where
// both blocks are `inner` consecutive rows of column-major arrays with `outer`
// columns and column stride `ld` (state_particle.cpp:136-142).
float eigen_slice_prod_sum(const float* a, const float* b, long inner, long outer, long ld) {
  if (inner == 0 || outer == 0) return 0.f;  // DenseBase::sum() on an empty expression
  const long ps = 4;
  const long pin = (inner / ps) * ps;
  float res;
  if (pin) {
    float p[4];
    for (int k = 0; k < 4; k++) p[k] = a[k] * b[k];
    for (long j = 0; j < outer; j++)
      for (long i = (j == 0 ? ps : 0); i < pin; i += ps)
        for (int k = 0; k < 4; k++) p[k] = p[k] + a[j * ld + i + k] * b[j * ld + i + k];
    res = predux4(p);
    for (long j = 0; j < outer; j++)
      for (long i = pin; i < inner; i++) res = res + a[j * ld + i] * b[j * ld + i];
  } else {
    // DefaultTraversal: coeff(0,0) then column-major order
    res = a[0] * b[0];
    for (long i = 1; i < inner; i++) res = res + a[i] * b[i];
    for (long j = 1; j < outer; j++)
      for (long i = 0; i < inner; i++) res = res + a[j * ld + i] * b[j * ld + i];
  }
  return res;
}

}  // namespace

// -----------------------------------------------------------------------------
// a1  ScanRendererPolar::renderSemanticTopDown   src/scan_renderer_polar.cpp:83-109
// pts: AoS, x,y at byte offsets 0,4, intensity at byte offset `intensity_off`
// (pcl::PointXYZI: 16).  imgs: C images, each n_theta x n_r column-major.
// Deviation: the reference indexes flatten_lut_[pt_class] unchecked (UB when
// out of range); here an out-of-range class drops the point.
// -----------------------------------------------------------------------------
ORC_API void orc_render_polar(const uint8_t* pts, int stride, int intensity_off, long n,
                              float res, float ang_res, int n_theta, int n_r,
                              const int* lut, int n_lut, int C, float* imgs) {
  if (C < 1) return;                                      // :85
  std::memset(imgs, 0, sizeof(float) * (size_t)C * n_theta * n_r);  // :88-90
  for (long idx = 0; idx < n; idx++) {                    // :93
    float x, y, inten;
    std::memcpy(&x, pts + idx * stride + 0, 4);
    std::memcpy(&y, pts + idx * stride + 4, 4);
    std::memcpy(&inten, pts + idx * stride + intensity_off, 4);
    if (x == 0 && y == 0) continue;                       // :95
    float theta = atan2(x, y);                            // :97 (float overload -> atan2f)
    float r = sqrt(x * x + y * y);                        // :98
    int theta_ind = f2i_x86(std::round(theta / ang_res) + n_theta / 2);  // :100
    int r_ind = f2i_x86(std::round(r / res));             // :101
    if (theta_ind >= 0 && theta_ind < n_theta && r_ind >= 0 && r_ind < n_r) {  // :102
      int pt_class = f2i_x86(inten);                      // :103
      if (pt_class < 0 || pt_class >= n_lut) continue;    // (UB in the reference)
      int f = lut[pt_class];
      if (f >= 0 && f < C) imgs[(size_t)f * n_theta * n_r + (size_t)r_ind * n_theta + theta_ind] += 1;  // :104-105
    }
  }
}

// a2  ScanRenderer::renderSemanticTopDown   src/scan_renderer.cpp:55-78
// imgs: C images rows x cols column-major; (y_ind, x_ind) -> x_ind*rows + y_ind.
ORC_API void orc_render_cart(const uint8_t* pts, int stride, int intensity_off, long n, float res,
                             int rows, int cols, const int* lut, int n_lut, int C, float* imgs) {
  if (C < 1) return;
  std::memset(imgs, 0, sizeof(float) * (size_t)C * rows * cols);
  for (long idx = 0; idx < n; idx++) {
    float x, y, inten;
    std::memcpy(&x, pts + idx * stride + 0, 4);
    std::memcpy(&y, pts + idx * stride + 4, 4);
    std::memcpy(&inten, pts + idx * stride + intensity_off, 4);
    if (x == 0 && y == 0) continue;                                   // :67
    int x_ind = f2i_x86(std::round(x / res) + cols / 2);              // :69  img_size[0] = cols
    int y_ind = f2i_x86(std::round(y / res) + rows / 2);              // :70
    if (x_ind >= 0 && x_ind < cols && y_ind >= 0 && y_ind < rows) {   // :71
      int pt_class = f2i_x86(inten);
      if (pt_class < 0 || pt_class >= n_lut) continue;
      int f = lut[pt_class];
      if (f >= 0 && f < C) imgs[(size_t)f * rows * cols + (size_t)x_ind * rows + y_ind] += 1;  // :74
    }
  }
}

// -----------------------------------------------------------------------------
// f2  ScanRendererPolar::renderGeometricTopDown   src/scan_renderer_polar.cpp:6-81
// The cloud is ORGANISED (width columns x height rows, point (col, row) at row * width + col) and walked column by
// column (:27-28).  Points go to angular bins in that order (theta index clamped, :36-37), every bin is sorted by
// descending planar range (std::sort, :50-52 — the order of exactly equal ranges is whatever introsort leaves; here the
// same std::sort on the same sequence), then walked with the slope rule: a step steeper than 1 marks an obstacle cell in
// imgs[1], a step flatter than 0.3 that does not follow an obstacle fills imgs[0] from the previous range bin up to this
// one.  imgs: 2 images n_theta x n_r column-major.
ORC_API void orc_render_geometric_polar(const uint8_t* pts, int stride, int width, int height, float res, float ang_res,
                                        int n_theta, int n_r, float* imgs) {
  std::memset(imgs, 0, sizeof(float) * 2 * (size_t)n_theta * n_r);                 // :12-14
  struct P4 { float x, y, z, r; };
  std::vector<std::vector<P4>> bins((size_t)n_theta);                             // :18-22
  for (int idx = 0; idx < width; idx++) {                                          // :27
    for (int idy = 0; idy < height; idy++) {                                       // :28
      float x, y, z;
      const uint8_t* q = pts + ((size_t)idy * width + idx) * stride;
      std::memcpy(&x, q, 4); std::memcpy(&y, q + 4, 4); std::memcpy(&z, q + 8, 4);
      if (x == 0 && y == 0) continue;                                              // :30
      float theta = atan2(x, y);                                                   // :32
      float r = sqrt(x * x + y * y);                                               // :33
      float t = std::round(theta / ang_res) + n_theta / 2;                         // :36-37, clamp<float>
      if (t != t) continue;                                                        // NaN: the reference indexes out of bounds
      t = (t < 0.f) ? 0.f : (((float)(n_theta - 1) < t) ? (float)(n_theta - 1) : t);
      bins[(size_t)(int)t].push_back(P4{x, y, z, r});                             // :39
    }
  }
  for (int theta_ind = 0; theta_ind < n_theta; theta_ind++) {                      // :47
    auto& bin = bins[(size_t)theta_ind];
    std::sort(bin.begin(), bin.end(), [](P4& a, P4& b) { return a.r > b.r; });    // :50-52
    float lx = 0, ly = 0, lz = 0;                                                  // :55
    bool last_high_grad = false;
    int last_r_ind = 0;
    for (const P4& pt : bin) {                                                     // :58
      const float dx = pt.x - lx, dy = pt.y - ly;
      float s2 = 0.f; s2 = s2 + dx * dx; s2 = s2 + dy * dy;                        // head<2>().norm(): sequential sum of squares
      const float dist = std::sqrt(s2);                                            // :59
      const float slope = abs(pt.z - lz) / dist;                                   // :60 (float abs: <math.h>)
      const int r_ind = f2i_x86(std::round(pt.r / res));                           // :61
      if (slope > 1) {                                                             // :63
        if (r_ind >= 0 && r_ind < n_r) imgs[(size_t)n_theta * n_r + (size_t)r_ind * n_theta + theta_ind] += 1;   // :64-66
        last_high_grad = true;
      } else if (slope < 0.3 && last_high_grad == false) {                         // :68 (double 0.3)
        for (int i = last_r_ind; i <= r_ind; i += 1)                               // :69
          if (i < n_r && i >= 0) imgs[(size_t)i * n_theta + theta_ind] += 1;       // :70-72 (i >= 0 always: ranges are >= 0)
      } else {
        last_high_grad = false;                                                    // :75
      }
      lx = pt.x; ly = pt.y; lz = pt.z;                                             // :77
      last_r_ind = r_ind;
    }
  }
}

// f2  ScanRenderer::renderGeometricTopDown   src/scan_renderer.cpp:7-53
// One pass per vertical scan line (column of the organised cloud); obstacle steps mark imgs[1], flat steps draw the
// segment from the previous cell to this one into imgs[0] (float parameter i += 1. / |diff| with the INTEGER norm of
// the index difference, :38-39).  imgs: 2 images rows x cols column-major, (y_ind, x_ind) at x_ind * rows + y_ind.
ORC_API void orc_render_geometric_cart(const uint8_t* pts, int stride, int width, int height, float res, int rows, int cols,
                                       float* imgs) {
  std::memset(imgs, 0, sizeof(float) * 2 * (size_t)rows * cols);
  const int sx = cols, sy = rows;                                                  // img_size = (cols, rows)  :10
  for (int idx = 0; idx < width; idx++) {                                          // :16
    float lx = 0, ly = 0, lz = 0;
    int last_x = sx / 2, last_y = sy / 2;                                          // :19
    bool last_high_grad = false;
    for (int idy = 0; idy < height; idy++) {                                       // :23
      float x, y, z;
      const uint8_t* q = pts + ((size_t)idy * width + idx) * stride;
      std::memcpy(&x, q, 4); std::memcpy(&y, q + 4, 4); std::memcpy(&z, q + 8, 4);
      if (x == 0 && y == 0) continue;                                              // :26
      const int x_ind = f2i_x86(std::round(x / res) + sx / 2);                     // :27
      const int y_ind = f2i_x86(std::round(y / res) + sy / 2);                     // :28
      const float dx = x - lx, dy = y - ly;
      float s2 = 0.f; s2 = s2 + dx * dx; s2 = s2 + dy * dy;
      const float dist = std::sqrt(s2);                                            // :30
      const float slope = abs(z - lz) / dist;                                      // :31
      if (slope > 1) {
        if (x_ind >= 0 && x_ind < sx && y_ind >= 0 && y_ind < sy) imgs[(size_t)rows * cols + (size_t)x_ind * rows + y_ind] += 1;   // :33-35
        last_high_grad = true;
      } else if (slope < 0.3 && last_high_grad == false) {
        const int ddx = x_ind - last_x, ddy = y_ind - last_y;                      // :38
        const int nrm = (int)std::sqrt((double)((long long)ddx * ddx + (long long)ddy * ddy));   // Vector2i::norm(): integer
        for (float i = 0; i < 1; i += 1. / nrm) {                                  // :39 (float += double)
          const int ix = f2i_x86(round((float)last_x + i * (float)ddx));           // :40
          const int iy = f2i_x86(round((float)last_y + i * (float)ddy));
          if (ix >= 0 && ix < sx && iy >= 0 && iy < sy) imgs[(size_t)ix * rows + iy] += 1;   // :41-44
        }
      } else {
        last_high_grad = false;
      }
      lx = x; ly = y; lz = z;                                                      // :49
      last_x = x_ind; last_y = y_ind;
    }
  }
}

// -----------------------------------------------------------------------------
// a3  TopDownMap::loadCompressedRasterMap   src/top_down_map.cpp:116-144
// img: row-major uint8 (cv::Mat) h_img x w_img with row stride `stride`.
// layers: C layers rows x cols column-major, rows=(int)(h_img/res), cols=(int)(w_img/res).
// Deviation: flatten_lut is indexed unchecked in the reference; here an index
// beyond n_lut means "no class" (the adapters pad the LUT to 256 x -1).
// -----------------------------------------------------------------------------
ORC_API void orc_map_dims(int h_img, int w_img, float res, int* rows, int* cols) {
  *rows = (int)(h_img / res);  // :121
  *cols = (int)(w_img / res);  // :122
}

ORC_API void orc_class_image_to_layers(const uint8_t* img, int h_img, int w_img, int stride,
                                       const int* lut, int n_lut, int C, float res, float* layers) {
  int rows, cols;
  orc_map_dims(h_img, w_img, res, &rows, &cols);
  const size_t L = (size_t)rows * cols;
  for (size_t i = 0; i < L * C; i++) layers[i] = 1.0f;   // :120-123
  for (size_t xi = 0; xi < (size_t)cols; xi++) {         // :135
    for (size_t yi = 0; yi < (size_t)rows; yi++) {       // :136
      // :137  map.size().height - yi*resolution - 1  : int - (size_t->float * float) - int, in float
      int src_row = std::max<int>(f2i_x86((float)h_img - (float)yi * res - 1), 0);
      int src_col = std::min<int>(f2i_x86((float)xi * res), w_img - 1);      // :138
      uint8_t v = img[(size_t)src_row * stride + src_col];
      int cls = (v < n_lut) ? lut[v] : -1;
      if (cls >= 0 && cls < C) layers[(size_t)cls * L + xi * rows + yi] = 0;  // :139-141
    }
  }
}

// -----------------------------------------------------------------------------
// a4  TopDownMap::computeDists   src/top_down_map.cpp:289-326
// In place over C column-major layers (rows x cols); mask out: rows x cols uint8.
// EDT: exact integer squared distance (two-pass lower envelope, integer
// arithmetic) then sqrtf((float)d2) — what cv::distanceTransform(DIST_L2,
// DIST_MASK_PRECISE) returns (verified against cv2 in tests/test_oracle_edt.py).
// -----------------------------------------------------------------------------
namespace {
const int64_t EDT_INF = (int64_t)1 << 40;

// 1D squared-distance transform of f (Felzenszwalb & Huttenlocher), integer-exact.
void dt1d(const int64_t* f, int n, int64_t* d, int* v, double* z) {
  int k = -1;
  for (int q = 0; q < n; q++) {
    if (f[q] >= EDT_INF) continue;
    while (true) {
      if (k < 0) { k = 0; v[0] = q; z[0] = -1e30; z[1] = 1e30; break; }
      int p = v[k];
      // intersection of parabolas rooted at p and q
      double s = ((double)(f[q] + (int64_t)q * q) - (double)(f[p] + (int64_t)p * p)) / (2.0 * (q - p));
      if (s <= z[k]) { k--; continue; }
      k++; v[k] = q; z[k] = s; z[k + 1] = 1e30; break;
    }
  }
  if (k < 0) { for (int q = 0; q < n; q++) d[q] = EDT_INF; return; }
  int j = 0;
  for (int q = 0; q < n; q++) {
    while (z[j + 1] < q) j++;
    // exactness guard: the envelope search uses doubles; re-check neighbours in integers
    int64_t best = (int64_t)(q - v[j]) * (q - v[j]) + f[v[j]];
    if (j > 0) best = std::min(best, (int64_t)(q - v[j - 1]) * (q - v[j - 1]) + f[v[j - 1]]);
    if (j < k) best = std::min(best, (int64_t)(q - v[j + 1]) * (q - v[j + 1]) + f[v[j + 1]]);
    d[q] = best;
  }
}
}  // namespace

// exact squared EDT of a binary image (seed where bin==0); out d2 (int64), col-major agnostic:
// treats the buffer as an n0 (fast) x n1 (slow) grid.
ORC_API void orc_edt_sq(const uint8_t* bin, int n0, int n1, int64_t* d2) {
  std::vector<int64_t> tmp((size_t)n0 * n1);
  // pass 1: along the fast axis
  {
    int n = std::max(n0, n1);
    std::vector<int64_t> f(n), d(n); std::vector<int> v(n + 1); std::vector<double> z(n + 2);
    for (int j = 0; j < n1; j++) {
      for (int i = 0; i < n0; i++) f[i] = bin[(size_t)j * n0 + i] == 0 ? 0 : EDT_INF;
      dt1d(f.data(), n0, d.data(), v.data(), z.data());
      for (int i = 0; i < n0; i++) tmp[(size_t)j * n0 + i] = d[i];
    }
    for (int i = 0; i < n0; i++) {
      for (int j = 0; j < n1; j++) f[j] = tmp[(size_t)j * n0 + i];
      dt1d(f.data(), n1, d.data(), v.data(), z.data());
      for (int j = 0; j < n1; j++) d2[(size_t)j * n0 + i] = d[j];
    }
  }
}

ORC_API void orc_compute_dists(float* layers, int rows, int cols, int C, float resolution, uint8_t* mask) {
  const size_t L = (size_t)rows * cols;
  // :294-299  mask = sum_c (uint8)layer_c ; threshold(mask, C-1, 255, BINARY)
  for (size_t i = 0; i < L; i++) {
    uint8_t m = 0;
    for (int c = 0; c < C; c++) m = (uint8_t)(m + (uint8_t)layers[(size_t)c * L + i]);  // Eigen cast<uint8_t>: truncation
    mask[i] = (m > C - 1) ? 255 : 0;
  }
  std::vector<uint8_t> bin(L);
  std::vector<int64_t> d2(L);
  for (int c = 0; c < C; c++) {
    float* lay = layers + (size_t)c * L;
    // :306 convertTo(CV_8UC1): saturate_cast<uchar>(cvRound(v)) — round half to even
    for (size_t i = 0; i < L; i++) {
      float v = lay[i];
      long r = lrintf(v);  // default rounding mode = nearest-even = cvRound
      bin[i] = (uint8_t)std::min<long>(std::max<long>(r, 0), 255);
    }
    orc_edt_sq(bin.data(), rows, cols, d2.data());   // :312
    for (size_t i = 0; i < L; i++) {
      float d = (d2[i] >= EDT_INF) ? 65536.0f : sqrtf((float)d2[i]);
      d = d * resolution;                            // :314
      d = (d > 50.0f) ? 50.0f : d;                   // :315 THRESH_TRUNC
      if (mask[i]) d = 0;                            // :317
      lay[i] = d;
    }
  }
  for (size_t i = 0; i < L; i++) mask[i] /= 255;     // :321
}

// a5  TopDownMap::getGeoRasterMap   src/top_down_map.cpp:410-427 (binary class layers in, 2 geo layers out)
ORC_API void orc_geo_raster(const float* class_layers, int rows, int cols, int C, float* geo) {
  const size_t L = (size_t)rows * cols;
  for (size_t i = 0; i < 2 * L; i++) geo[i] = 0;                       // :413-415
  for (int c = 3; c < C; c++)
    for (size_t i = 0; i < L; i++) geo[L + i] += 1 - class_layers[(size_t)c * L + i];  // :417-419
  for (int g = 0; g < 2; g++)
    for (size_t i = 0; i < L; i++) { float v = std::min(geo[g * L + i], 1.0f); geo[g * L + i] = 1 - v; }  // :422-425
  for (size_t i = 0; i < L; i++) geo[i] = 1 - geo[L + i];              // :426
}

// -----------------------------------------------------------------------------
// a6  polar offset table: TopDownMap::samplePts (:367-389, NDEBUG semantics) +
// TopDownMapPolar::samplePtsPolar (src/top_down_map_polar.cpp:7-19).
// tab: 2 x P (tab[2*p+0] = cos*rho -> row offset, tab[2*p+1] = sin*rho -> col offset).
// NOTE: the reference evaluates Eigen's packet cos/sin; libm cosf/sinf is used
// here.  The table is an INPUT to both the oracle and the device library, so
// this choice does not enter any parity comparison.
// -----------------------------------------------------------------------------
ORC_API void orc_polar_table(int n_theta, int n_r, float ang_res, float resolution, float* tab) {
  const int P = n_theta * n_r;
  // LinSpaced(rows, -res*(rows-1)/2., res*(rows-1)/2.) with res = 1, step == 1 exactly
  const float lo0 = (float)(-1.0f * (n_theta - 1) / 2.);
  const float lo1 = (float)(-1.0f * (n_r - 1) / 2.);
  const float inv_res = (float)(1. / resolution);          // row(1) *= 1./resolution (scalar cast to float)
  for (int p = 0; p < P; p++) {
    float a0 = lo0 + (float)(p % n_theta) * 1.0f;          // :376 (rot = 0: rotation is exact identity)
    float a1 = lo1 + (float)(p / n_theta) * 1.0f;          // :378
    a1 = a1 + (-lo1);                                      // polar :11 shift radius to start at 0
    a0 = a0 * ang_res;                                     // :13
    a1 = a1 * inv_res;                                     // :14
    tab[2 * p + 0] = cosf(a0) * a1;                        // :17
    tab[2 * p + 1] = sinf(a0) * a1;                        // :18
  }
}

// -----------------------------------------------------------------------------
// a7  TopDownMapPolar::getLocalMap   src/top_down_map_polar.cpp:21-53
// layers: C col-major rows x cols (post-computeDists), mask rows x cols.
// dists: C x P, mask_out: P.
// -----------------------------------------------------------------------------
ORC_API void orc_local_map_polar(const float* layers, const uint8_t* mask, int rows, int cols, int C,
                                 float resolution, const float* tab, int P, float cx, float cy,
                                 float scale, float res, float* dists, uint8_t* mask_out) {
  if (C < 1) return;
  const size_t L = (size_t)rows * cols;
  const float oy = cy / resolution, ox = cx / resolution;
  for (int p = 0; p < P; p++) {
    float q0 = tab[2 * p + 0] * scale * res;   // :28 (ang_sample_pts_*scale)*res
    float q1 = tab[2 * p + 1] * scale * res;
    q0 = q0 + oy;                              // :29 row(0) += center[1]/resolution
    q1 = q1 + ox;                              // :30
    int r = f2i_x86(std::round(q0));           // :31 round half away, cast<int>
    int c = f2i_x86(std::round(q1));
    bool in = r >= 0 && r < rows && c >= 0 && c < cols;
    for (int k = 0; k < C; k++) dists[(size_t)k * P + p] = in ? layers[(size_t)k * L + (size_t)c * rows + r] : 0.f;  // :33-42
    mask_out[p] = in ? mask[(size_t)c * rows + r] : 1;  // :44-52
  }
}

// a8  TopDownMap::getLocalMap (Cartesian)   src/top_down_map.cpp:429-459 with samplePts :367-389
ORC_API void orc_local_map_cart(const float* layers, const uint8_t* mask, int rows, int cols, int C,
                                float resolution, float cx, float cy, float rot, float res,
                                int out_rows, int out_cols, float* dists, uint8_t* mask_out) {
  if (C < 1) return;
  const size_t L = (size_t)rows * cols;
  const int P = out_rows * out_cols;
  const float sres = res / resolution;                       // :434
  const float c0 = cx / resolution, c1 = cy / resolution;
  // LinSpaced(rows, -res*(rows-1)/2., res*(rows-1)/2.) : float*int -> float, /2. -> double, -> float
  auto linsp = [](int size, float sres_, int i) -> float {
    float low = (float)(-sres_ * (size - 1) / 2.);
    float high = (float)(sres_ * (size - 1) / 2.);
    if (size == 1) return low;  // Eigen: m_size1 = 1, step = 0 -> i==m_size1 never for i=0 -> low + 0
    float step = (high - low) / (float)(size - 1);
    bool flip = std::abs(high) < std::abs(low);
    int size1 = size - 1;
    if (flip) return (i == 0) ? low : (high - (float)(size1 - i) * step);
    return (i == size1) ? high : (low + (float)i * step);
  };
  const float cr = cosf(rot), sr = sinf(rot);                // :383 cos(rot) float overload
  for (int p = 0; p < P; p++) {
    float x = linsp(out_rows, sres, p % out_rows);           // pts(0,p)
    float y = linsp(out_cols, sres, p / out_rows);           // pts(1,p)
    // rotm * pts : [c -s; s c]   (2x2 * 2xN product, no FMA in a baseline build)
    float xr = cr * x + (-sr) * y;
    float yr = sr * x + cr * y;
    xr = xr + c1;                                            // :387 x_vals += center[1]
    yr = yr + c0;                                            // :388 y_vals += center[0]
    int r = f2i_x86(std::round(xr));                         // :437
    int c = f2i_x86(std::round(yr));
    bool in = r >= 0 && r < rows && c >= 0 && c < cols;
    for (int k = 0; k < C; k++) dists[(size_t)k * P + p] = in ? layers[(size_t)k * L + (size_t)c * rows + r] : 0.f;
    mask_out[p] = in ? mask[(size_t)c * rows + r] : 1;
  }
}

// -----------------------------------------------------------------------------
// State / FilterParams mirrors   include/top_down_render/state_particle.h:9-38
// -----------------------------------------------------------------------------
struct OrcState {
  float init_x_px, init_y_px, dx_m, dy_m, theta, scale;
  uint8_t have_init; uint8_t pad[3];
};
static_assert(sizeof(OrcState) == 28, "State is 28 bytes");

struct OrcFilterParams {
  float regularization;
  int force_on_map;
  float fixed_scale, scale_log_min, scale_log_max;
  float map_width, map_height;     // StateParticle::width_/height_ (state_particle.cpp:46-47)
  int num_classes;
  float class_weights[16];
};

// rot -> row shift   state_particle.cpp:123-128
ORC_API int orc_rot_to_shift(float rot, int num_bins) {
  int rot_shift = d2i_x86(std::round((double)(rot * num_bins / 2) / M_PI));
  if (rot_shift == INT_MIN) return 0;  // NaN/inf rot: reference loops forever; not reachable with finite theta
  while (rot_shift >= num_bins) rot_shift -= num_bins;
  while (rot_shift < 0) rot_shift += num_bins;
  return rot_shift;
}

// theta-search candidate list   state_particle.cpp:197   for (float t=0; t<2*M_PI; t+=2*M_PI/40)
ORC_API int orc_search_list(int num_bins, float* thetas, int* shifts, int cap) {
  int n = 0;
  for (float t = 0; t < 2 * M_PI; t += 2 * M_PI / 40) {
    if (n < cap) { thetas[n] = t; shifts[n] = orc_rot_to_shift(t, num_bins); }
    n++;
  }
  return n;
}

// a10  StateParticle::getCostForRot   src/state_particle.cpp:112-155
// scan: C x (n_theta x n_r) col-major; classes likewise (the gathered local map);
// known = 1 - mask (float, n_theta x n_r).
ORC_API float orc_cost_for_shift(const float* scan, const float* classes, const float* known,
                                 int n_theta, int n_r, int C, const float* class_weights, int rot_shift) {
  const int P = n_theta * n_r;
  // :117  static_cast<float>(mask.sum())/mask.size() < 0.5   (integer-valued float sum: exact in any order)
  float ksum = eigen_linear_sum(known, P);
  if ((double)(ksum / (float)(long)P) < 0.5) return std::numeric_limits<float>::quiet_NaN();
  float cost = 0, normalization = 0;
  const int s = rot_shift, t = n_theta - rot_shift;
  for (int i = 0; i < C; i++) {
    const float* sc = scan + (size_t)i * P;
    const float* cl = classes + (size_t)i * P;
    // :136  scan.topRows(s) * classes.bottomRows(s)
    float S1 = eigen_slice_prod_sum(sc, cl + t, s, n_r, n_theta);
    cost = (float)((double)cost + (double)S1 * 0.01 * (double)class_weights[i]);
    // :138  scan.bottomRows(N-s) * classes.topRows(N-s)
    float S2 = eigen_slice_prod_sum(sc + s, cl, t, n_r, n_theta);
    cost = (float)((double)cost + (double)S2 * 0.01 * (double)class_weights[i]);
    normalization += eigen_slice_prod_sum(sc, known + t, s, n_r, n_theta);      // :141
    normalization += eigen_slice_prod_sum(sc + s, known, t, n_r, n_theta);      // :142
  }
  return cost / normalization;   // :154
}

// a9  StateParticle::computeWeight   src/state_particle.cpp:157-219
// Returns the weight; updates st->theta / st->have_init as the reference does.
// geo_layers may be null; if non-null the (unused) geo gather of :189 is performed ("literal" cost).
static float compute_weight_impl(OrcState* st, const OrcFilterParams* fp, const float* layers,
                                 const uint8_t* mask, const float* geo_layers, int rows, int cols,
                                 float resolution, const float* tab, int n_theta, int n_r,
                                 const float* scan, float res, const float* search_thetas,
                                 const int* search_shifts, int n_search) {
  const int C = fp->num_classes, P = n_theta * n_r;
  float cx = st->dx_m * st->scale + st->init_x_px;   // :161
  float cy = st->dy_m * st->scale + st->init_y_px;   // :162
  if (fp->force_on_map) {                            // :163-168
    if (cx < 0 || cy < 0 || cx > fp->map_width || cy > fp->map_height) return 0;
  }
  if (fp->fixed_scale < 0) {                         // :169-176
    if ((double)st->scale < std::pow(10, fp->scale_log_min) || (double)st->scale > std::pow(10, fp->scale_log_max)) return 0;
  }
  // :178-186 per-particle allocations, kept literally
  std::vector<float> classes((size_t)C * P);
  std::vector<uint8_t> m(P);
  orc_local_map_polar(layers, mask, rows, cols, C, resolution, tab, P, cx, cy, st->scale, res, classes.data(), m.data());  // :188
  if (geo_layers) {                                  // :189 result unused (:145-152 commented out)
    std::vector<float> geo((size_t)2 * P);
    std::vector<uint8_t> gm(P);
    orc_local_map_polar(geo_layers, mask, rows, cols, 2, resolution, tab, P, cx, cy, st->scale, res, geo.data(), gm.data());
  }
  std::vector<float> known(P);
  float best_cost = std::numeric_limits<float>::max();   // :193
  float best_theta = 0;
  if (!st->have_init) {                                  // :195-206
    for (int k = 0; k < n_search; k++) {
      for (int p = 0; p < P; p++) known[p] = 1 - (float)m[p];   // :199 temp rebuilt per call
      float cost = orc_cost_for_shift(scan, classes.data(), known.data(), n_theta, n_r, C, fp->class_weights, search_shifts[k]);
      if (cost < best_cost) { best_cost = cost; best_theta = search_thetas[k]; }
    }
    st->theta = best_theta;
    st->have_init = 1;
  } else {
    for (int p = 0; p < P; p++) known[p] = 1 - (float)m[p];
    best_cost = orc_cost_for_shift(scan, classes.data(), known.data(), n_theta, n_r, C, fp->class_weights,
                                   orc_rot_to_shift(st->theta, n_theta));   // :208-209
  }
  return (float)(1. / (double)(best_cost + fp->regularization));   // :212
}

ORC_API float orc_compute_weight(OrcState* st, const OrcFilterParams* fp, const float* layers,
                                 const uint8_t* mask, const float* geo_layers, int rows, int cols,
                                 float resolution, const float* tab, int n_theta, int n_r,
                                 const float* scan, float res, const float* search_thetas,
                                 const int* search_shifts, int n_search) {
  return compute_weight_impl(st, fp, layers, mask, geo_layers, rows, cols, resolution, tab, n_theta, n_r,
                             scan, res, search_thetas, search_shifts, n_search);
}

// ParticleFilter::update :104-105 — the for_each(par) region; threads stand in for TBB.
ORC_API void orc_score_all(OrcState* states, long n, const OrcFilterParams* fp, const float* layers,
                           const uint8_t* mask, const float* geo_layers, int rows, int cols, float resolution,
                           const float* tab, int n_theta, int n_r, const float* scan, float res,
                           const float* search_thetas, const int* search_shifts, int n_search,
                           float* weights, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int t) {
    long lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    for (long i = lo; i < hi; i++)
      weights[i] = compute_weight_impl(&states[i], fp, layers, mask, geo_layers, rows, cols, resolution, tab,
                                       n_theta, n_r, scan, res, search_thetas, search_shifts, n_search);
  };
  if (n_threads == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
  for (auto& x : th) x.join();
}

// Exhaustive grid (BASELINE cfg4): value of getCostForRot at every (centre, shift).
// costs: n_centres x n_shifts.  No counterpart function in the reference; composition of a7 + a10.
ORC_API void orc_cost_grid(const float* centers_xy, long n, float scale, const OrcFilterParams* fp,
                           const float* layers, const uint8_t* mask, int rows, int cols, float resolution,
                           const float* tab, int n_theta, int n_r, const float* scan, float res,
                           const int* shifts, int n_shifts, float* costs, int n_threads) {
  const int C = fp->num_classes, P = n_theta * n_r;
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int t) {
    std::vector<float> classes((size_t)C * P), known(P);
    std::vector<uint8_t> m(P);
    long lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    for (long i = lo; i < hi; i++) {
      orc_local_map_polar(layers, mask, rows, cols, C, resolution, tab, P, centers_xy[2 * i], centers_xy[2 * i + 1],
                          scale, res, classes.data(), m.data());
      for (int p = 0; p < P; p++) known[p] = 1 - (float)m[p];
      for (int k = 0; k < n_shifts; k++)
        costs[i * n_shifts + k] = orc_cost_for_shift(scan, classes.data(), known.data(), n_theta, n_r, C,
                                                     fp->class_weights, shifts[k]);
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
  for (auto& x : th) x.join();
}

// -----------------------------------------------------------------------------
// a11  ParticleFilter::update — weights   src/particle_filter.cpp:107-147
// The three `int i;` loop counters are uninitialised in the reference (UB); the
// intended i = 0 is used.  w: in raw weights, out normalised.  Returns arg-max.
// stats (optional, 6 floats): sum, num_valid, mean, bottom_stddev, num_under, fallback
// -----------------------------------------------------------------------------
ORC_API long orc_normalize(float* w, const float* last_dist, long n, float* stats) {
  float sum = 0; int num_valid = 0;
  for (long i = 0; i < n; i++) {                       // :110-116
    if (!std::isnan(w[i])) { sum += w[i]; ++num_valid; }
  }
  float mean = sum / num_valid;                        // :117
  float bottom_stddev = 0; int num_under = 0;
  for (long i = 0; i < n; i++) {                       // :120-125
    if (!std::isnan(w[i]) && w[i] < mean) {
      bottom_stddev += std::pow(w[i] - mean, 2);       // pow(float,int) -> double; float += double
      ++num_under;
    }
  }
  bottom_stddev = std::sqrt(bottom_stddev / num_under);  // :126 (float)
  int fallback = 0;
  if (sum == 0 || num_under < 1) {                     // :129-131
    for (long i = 0; i < n; i++) w[i] = 1;
    fallback = 1;
  } else {
    float rep = mean - bottom_stddev;                  // :133
    for (long i = 0; i < n; i++) if (std::isnan(w[i])) w[i] = rep;
  }
  float s1 = eigen_linear_sum(w, n);                   // :135
  for (long i = 0; i < n; i++) w[i] = w[i] / s1;
  for (long i = 0; i < n; i++) {                       // :138-141
    float d = std::min<float>(last_dist[i] * 5, 1);
    w[i] = d * w[i] + (1 - d) / (float)(size_t)n;
  }
  float s2 = eigen_linear_sum(w, n);                   // :142
  for (long i = 0; i < n; i++) w[i] = w[i] / s2;
  long arg = 0; float best = w[0];                     // :145-146 maxCoeff: first maximum
  for (long i = 1; i < n; i++) if (w[i] > best) { best = w[i]; arg = i; }
  if (stats) { stats[0] = sum; stats[1] = (float)num_valid; stats[2] = mean; stats[3] = bottom_stddev;
               stats[4] = (float)num_under; stats[5] = (float)fallback; }
  return arg;
}

// a12  systematic resample   src/particle_filter.cpp:172-185
// literal O(N*M) form.
ORC_API void orc_resample_literal(const float* w, long n, float shift, int M, int* idx) {
  for (int i = 0; i < M; i++) {
    float running_sum = 0;
    float sample = ((float)i + shift) / M;   // float / int -> float
    long j = 0;
    for (; j < n; j++) {
      running_sum += w[j];
      if (running_sum > sample || j == n - 1) break;
    }
    idx[i] = (int)j;
  }
}

// Same indices in one pass: the float prefix is the same sequence for every i (it restarts
// from 0 and adds the same terms in the same order), so "first j with prefix_j > sample" is
// a search over ONE sequence; a running maximum makes the search monotone even if some
// weights are negative (first j with prefix_j > x == first j with max_{k<=j} prefix_k > x).
ORC_API void orc_resample_fast(const float* w, long n, float shift, int M, int* idx, float* prefix_out) {
  std::vector<float> rmax(n);
  float run = 0, mx = -std::numeric_limits<float>::infinity();
  for (long j = 0; j < n; j++) {
    run += w[j];
    if (run > mx) mx = run;
    rmax[j] = mx;
    if (prefix_out) prefix_out[j] = run;
  }
  long j = 0;
  for (int i = 0; i < M; i++) {
    float sample = ((float)i + shift) / M;
    // samples are non-decreasing in i (RN is monotone), so j never moves back
    while (j < n - 1 && !(rmax[j] > sample)) j++;
    idx[i] = (int)j;
  }
}

// the single uniform draw of :172-173 from a seeded engine (libstdc++ generate_canonical<float,24>)
ORC_API float orc_uniform_draw(uint32_t seed) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<float> shift_dist(0., 1.);
  return shift_dist(gen);
}
// raw output number `discard` of std::mt19937(seed): where a shared engine stands after `discard` outputs were consumed
ORC_API uint32_t orc_engine_peek(uint32_t seed, uint64_t discard) {
  std::mt19937 gen(seed);
  gen.discard(discard);
  return (uint32_t)gen();
}
// the same draw from the shared engine after `discard` earlier outputs
ORC_API float orc_uniform_draw_from(uint32_t seed, uint64_t discard) {
  CountingMt gen(seed, discard);
  std::uniform_real_distribution<float> shift_dist(0., 1.);
  return shift_dist(gen);
}

// -----------------------------------------------------------------------------
// a13  pose   src/state_particle.cpp:98-102, src/particle_filter.cpp:191-236
// -----------------------------------------------------------------------------
static inline void ml_state(const OrcState& s, float out[4]) {
  out[0] = s.dx_m * s.scale + s.init_x_px;
  out[1] = s.dy_m * s.scale + s.init_y_px;
  out[2] = s.theta;
  out[3] = s.scale;
}

ORC_API void orc_mean_likelihood(const OrcState* st, long n, float mean[4]) {
  float acc[4] = {0, 0, 0, 0}; float cos_sum = 0, sin_sum = 0;
  for (long i = 0; i < n; i++) {                       // :195-200
    float s[4]; ml_state(st[i], s);
    for (int k = 0; k < 4; k++) acc[k] = acc[k] + s[k];
    cos_sum += cos(s[2]);
    sin_sum += sin(s[2]);
  }
  for (int k = 0; k < 4; k++) mean[k] = acc[k] / (float)(size_t)n;   // :201
  mean[2] = atan2(sin_sum / (float)(size_t)n, cos_sum / (float)(size_t)n);  // :202
}

static void cov_about(const OrcState* st, long n, const float ref[4], float cov[16]) {
  for (int k = 0; k < 16; k++) cov[k] = 0;
  for (long i = 0; i < n; i++) {
    float s[4]; ml_state(st[i], s);
    for (int k = 0; k < 4; k++) s[k] = s[k] - ref[k];
    while (s[2] > M_PI) s[2] -= 2 * M_PI;              // float compared/updated through double
    while (s[2] < -M_PI) s[2] += 2 * M_PI;
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) cov[c * 4 + r] = cov[c * 4 + r] + s[r] * s[c];
  }
  float den = (float)(size_t)(n - 1);
  for (int k = 0; k < 16; k++) cov[k] = cov[k] / den;  // :219
}

ORC_API void orc_mean_cov(const OrcState* st, long n, float mean[4], float cov[16]) {   // :205-220
  orc_mean_likelihood(st, n, mean);
  cov_about(st, n, mean, cov);
}
ORC_API void orc_ml_cov(const OrcState* st, long n, long argmax, float ml[4], float cov[16]) {  // :222-236
  ml_state(st[argmax], ml);
  cov_about(st, n, ml, cov);
}

// -----------------------------------------------------------------------------
// refine_map-style binning rule (BASELINE cfg5)   src/refine_map.cpp:76-94
// points (x,y) double-free restatement on floats: ind = floor(pt/res) + (int)(centre/res)
// counts saturate like the reference's uint8 "+= 1" (wraps mod 256) — kept as wrap.
// -----------------------------------------------------------------------------
ORC_API void orc_refine_bin(const float* xy, const int* cls, long n, float res, float cx, float cy,
                            int width, int height, int C, uint8_t* maps /* C x height x width row-major */) {
  std::memset(maps, 0, (size_t)C * width * height);
  for (long i = 0; i < n; i++) {
    int ix = (int)std::floor((double)xy[2 * i] / res) + (int)(cx / res);
    int iy = (int)std::floor((double)xy[2 * i + 1] / res) + (int)(cy / res);
    if (ix < 0 || ix >= width || iy < 0 || iy >= height) continue;
    if (cls[i] < 0 || cls[i] >= C) continue;
    maps[(size_t)cls[i] * width * height + (size_t)iy * width + ix] += 1;
  }
}

// -----------------------------------------------------------------------------
// SURVEY 8f rank 1  propagate   src/state_particle.cpp:57-78, src/particle_filter.cpp:86-92
// -----------------------------------------------------------------------------
// Literal restatement: one shared std::mt19937 walked in particle order, fresh normal_distribution<float> objects per
// particle exactly as the reference constructs them (theta_dist called once, disp_dist twice — the second call
// returns the polar method's saved value — scale_dist once unless frozen).  Eigen::Rotation2D<float>(theta) * trans
// is [c -s; s c] * trans with std::cos / std::sin on float; Vector2f::norm() is sqrt(x*x + y*y).
// z_out (may be NULL): the STANDARD normal variate behind every draw, 4 per particle (theta, dx, dy, scale; scale
// = 0 when frozen), i.e. (draw - mean) / stddev as libstdc++ computes it before `* stddev + mean` — recovered by
// running a second engine in lock step through N(0,1) objects, which consume the identical uniforms.
// orc_propagate_from: the engine is std::mt19937(seed) after `discard` outputs (what earlier calls on the shared engine
// consumed); draws_out (may be NULL) receives the outputs this call consumed.
ORC_API void orc_propagate_from(OrcState* st, float* last_dist, long n, float tx, float ty, float omega, int scale_freeze,
                                float pos_cov, float theta_cov, uint32_t seed, uint64_t discard, float* z_out, uint64_t* draws_out);
ORC_API void orc_propagate(OrcState* st, float* last_dist, long n, float tx, float ty, float omega, int scale_freeze,
                           float pos_cov, float theta_cov, uint32_t seed, float* z_out) {
  orc_propagate_from(st, last_dist, n, tx, ty, omega, scale_freeze, pos_cov, theta_cov, seed, 0, z_out, nullptr);
}
ORC_API void orc_propagate_from(OrcState* st, float* last_dist, long n, float tx, float ty, float omega, int scale_freeze,
                                float pos_cov, float theta_cov, uint32_t seed, uint64_t discard, float* z_out, uint64_t* draws_out) {
  CountingMt gen(seed, discard), gen_z(seed, discard);
  for (long i = 0; i < n; i++) {
    OrcState& s = st[i];
    const float c = std::cos(s.theta), sn = std::sin(s.theta);
    const float gx = c * tx - sn * ty, gy = sn * tx + c * ty;                  // :58
    const float lx = s.dx_m, ly = s.dy_m;                                      // :59
    s.dx_m += gx;                                                              // :60
    s.dy_m += gy;                                                              // :61
    const float dist = std::sqrt(gx * gx + gy * gy);                           // :63
    std::normal_distribution<float> disp_dist{0, pos_cov * dist};              // :64
    std::normal_distribution<float> theta_dist{0, theta_cov * dist};           // :65
    s.theta += theta_dist(gen) + omega;                                        // :67
    s.dx_m += disp_dist(gen);                                                  // :68
    s.dy_m += disp_dist(gen);                                                  // :69
    if (!scale_freeze) {                                                       // :71-74
      std::normal_distribution<float> scale_dist{1, static_cast<float>(std::min(2. / dist, 0.02))};
      s.scale *= scale_dist(gen);
    }
    const float mx = lx - s.dx_m, my = ly - s.dy_m;                            // :76
    last_dist[i] = std::sqrt(mx * mx + my * my);                               // :77
    // the standard variates, from the twin engine (same uniforms, unit distributions)
    std::normal_distribution<float> zt{0, 1}, zd{0, 1}, zs{0, 1};
    const float z0 = zt(gen_z), z1 = zd(gen_z), z2 = zd(gen_z), z3 = scale_freeze ? 0.f : zs(gen_z);
    if (z_out) { z_out[4 * i] = z0; z_out[4 * i + 1] = z1; z_out[4 * i + 2] = z2; z_out[4 * i + 3] = z3; }
  }
  if (draws_out) *draws_out = gen.n;
}

// -----------------------------------------------------------------------------
// SURVEY 8f rank 3  vector map -> binary class layers   src/top_down_map.cpp:391-408 (getRasterMap),
// :367-389 (samplePts), :328-365 (getClasses: even-odd rule over every polygon of a class, then "only one ground
// type per cell" over the exclusive classes)
// -----------------------------------------------------------------------------
// verts: x, y pairs of all polygons back to back (y already flipped as loadSvg does, :89); polygon k owns vertices
// [poly_start[k], poly_start[k+1]) and belongs to flattened class poly_class[k] (polygons of a class in the order
// given).  layers: C col-major rows x cols images, 0 inside a polygon of the class, 1 elsewhere, with
// rows = (int)(map_h / resolution), cols = (int)(map_w / resolution) (:399-400).
ORC_API void orc_raster_polygons(const float* verts, const int* poly_start, const int* poly_class, int n_poly, int map_w,
                                 int map_h, float rot, float resolution, int C, const int* exclusive, int n_excl,
                                 float* layers) {
  const int rows = (int)(map_h / resolution), cols = (int)(map_w / resolution);
  const size_t L = (size_t)rows * cols;
  auto linsp = [](int size, float sres_, int i) -> float {       // Eigen LinSpaced, as in orc_local_map_cart
    float low = (float)(-sres_ * (size - 1) / 2.);
    float high = (float)(sres_ * (size - 1) / 2.);
    if (size == 1) return low;
    float step = (high - low) / (float)(size - 1);
    bool flip = std::abs(high) < std::abs(low);
    int size1 = size - 1;
    if (flip) return (i == 0) ? low : (high - (float)(size1 - i) * step);
    return (i == size1) ? high : (low + (float)i * step);
  };
  const float cr = cosf(rot), sr = sinf(rot);
  const float c0 = (float)map_w / 2, c1 = (float)map_h / 2;       // map_size.cast<float>() / 2  (:405)
  std::vector<float> py(L), px(L);
  for (size_t p = 0; p < L; p++) {
    float a = linsp(rows, resolution, (int)(p % rows));           // pts(0, p)
    float b = linsp(cols, resolution, (int)(p / rows));           // pts(1, p)
    float ar = cr * a + (-sr) * b, br = sr * a + cr * b;          // rotm * pts
    py[p] = ar + c1;                                              // x_vals += center[1]
    px[p] = br + c0;                                              // y_vals += center[0]
  }
  for (int c = 0; c < C; c++) {
    float* out = layers + (size_t)c * L;
    for (size_t p = 0; p < L; p++) out[p] = -1.f;                 // class_fills = -1
    std::vector<float> buf(L);
    for (int k = 0; k < n_poly; k++) {
      if (poly_class[k] != c) continue;
      const float* v = verts + 2 * (size_t)poly_start[k];
      const int n = poly_start[k + 1] - poly_start[k];
      for (size_t p = 0; p < L; p++) buf[p] = -1.f;
      int j = n - 1;
      for (int i = 0; i < n; i++) {
        const float xi = v[2 * i], yi = v[2 * i + 1], xj = v[2 * j], yj = v[2 * j + 1];
        for (size_t p = 0; p < L; p++) {
          const bool a = (py[p] < yi) != (py[p] < yj);
          const bool b = px[p] < (xi + ((xj - xi) * (py[p] - yi) / (yj - yi)));
          buf[p] *= (float)(-2 * (int)(a && b)) + 1;              // -2*cond.cast<float>() + 1   (:343-346)
        }
        j = i;
      }
      for (size_t p = 0; p < L; p++) out[p] = std::max(out[p], buf[p]);
    }
    for (size_t p = 0; p < L; p++) { out[p] *= -1; out[p] += 1; out[p] /= 2; }       // :351-353
  }
  for (int a = 0; a < n_excl; a++) {                              // :357-364
    const int under = exclusive[a];
    for (int b = 0; b < n_excl; b++) {
      const int cls = exclusive[b];
      if (under < cls)
        for (size_t p = 0; p < L; p++) layers[(size_t)under * L + p] += 1 - layers[(size_t)cls * L + p];
    }
    for (size_t p = 0; p < L; p++) layers[(size_t)under * L + p] = std::min(layers[(size_t)under * L + p], 1.f);
  }
}

// -----------------------------------------------------------------------------
// SURVEY 8f rank 2  ActiveLocalizer::getBestRelPos   src/active_localizer.cpp:7-82 (+ getLocalMap :22-42,
// computeTotalDifference :7-20).  The geometric twin of the polar gather (top_down_map_polar.cpp:55-76) is
// orc_local_map_polar on the two geo layers with the mask ignored.
// -----------------------------------------------------------------------------
// preds: n x (x, y, theta).  local maps are n_theta x n_r (the reference hard-codes 100 x 25), gathered with the
// 4-argument getLocalMap(center, res = 2, ...) i.e. scale = 1.  rel: (dist, theta) of the best relative position.
ORC_API void orc_active_best_rel_pos(const float* layers, const uint8_t* mask, int rows, int cols, int C, float resolution,
                                     const float* tab, int n_theta, int n_r, const float* preds, int n, float rel[2],
                                     float* best_diff_out) {
  const int P = n_theta * n_r;
  std::vector<std::vector<float>> local((size_t)n, std::vector<float>((size_t)C * P));
  std::vector<float> orig((size_t)C * P);
  std::vector<uint8_t> m(P);
  float dist = 50;                                                          // :59
  float best_diff = 0;
  float best0 = 0, best1 = 0;
  while (best_diff < 6000 && dist < 150) {                                  // :62
    for (float theta = 0; theta < 2 * M_PI; theta += M_PI / 8) {            // :63 (float += double)
      for (int idx = 0; idx < n; idx++) {
        const float* pred = preds + 3 * idx;
        const float px = pred[0] + dist * std::cos(theta + pred[2]);       // :67 Vector2f(cos, sin) * dist, float
        const float py = pred[1] + dist * std::sin(theta + pred[2]);
        orc_local_map_polar(layers, mask, rows, cols, C, resolution, tab, P, px, py, 1.f, 2.f, orig.data(), m.data());   // :29
        const int num_bins = n_theta;
        int rot_shift = static_cast<int>(std::round(pred[2] * num_bins / 2 / M_PI));      // :32
        while (rot_shift >= num_bins) rot_shift -= num_bins;
        while (rot_shift < 0) rot_shift += num_bins;
        for (int c = 0; c < C; c++)                                         // :38-41 rows rotate down by rot_shift
          for (int col = 0; col < n_r; col++)
            for (int r = 0; r < num_bins; r++) {
              const int src = r < rot_shift ? r + num_bins - rot_shift : r - rot_shift;
              local[idx][(size_t)c * P + (size_t)col * num_bins + r] = orig[(size_t)c * P + (size_t)col * num_bins + src];
            }
      }
      float total = 0;                                                      // :7-20
      int cnt = 0;
      for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
          for (int c = 0; c < C; c++) {
            float sum = 0;
            for (int p = 0; p < P; p++) sum += std::abs(local[i][(size_t)c * P + p] - local[j][(size_t)c * P + p]);
            total += sum;
            cnt += 1;
          }
      const float diff = total / cnt;
      if (diff > best_diff) { best_diff = diff; best0 = dist; best1 = theta; }   // :74-77
    }
    dist += 25;                                                             // :80
  }
  rel[0] = best0; rel[1] = best1;
  if (best_diff_out) *best_diff_out = best_diff;
}

// -----------------------------------------------------------------------------
// SURVEY 8f rank 4  the sample matrix of ParticleFilter::computeGMM (src/particle_filter.cpp:262-272) and the adaptive
// particle count of update() (:151-158)
// -----------------------------------------------------------------------------
ORC_API void orc_gmm_samples(const OrcState* st, long n, int num_samples, double* samples /* num_samples x 4 */) {
  for (int i = 0; i < num_samples; i++) {
    const long idx = std::min<long>(n - 1, (long)i * n / num_samples);      // :265-266
    float s[4]; ml_state(st[idx], s);
    samples[4 * i + 0] = s[0];
    samples[4 * i + 1] = s[1];
    samples[4 * i + 2] = 50 * std::cos(s[2]);                               // int * float(cosf) -> float -> double
    samples[4 * i + 3] = 50 * std::sin(s[2]);
  }
}
// covs: n_cov 4x4 matrices (row-major here; only the top-left 2x2 block is read).  Eigen::eigenvalues() of a real 2x2
// matrix [[a, b], [c, d]]: (tr +- sqrt(tr^2 - 4 det)) / 2, complex when the discriminant is negative (then .real() = tr/2)
ORC_API int orc_adaptive_count(const float* covs, int n_cov, int last_num_particles, int max_num_particles) {
  int num = 0;
  for (int k = 0; k < n_cov; k++) {
    const float a = covs[16 * k + 0], b = covs[16 * k + 1], c = covs[16 * k + 4], d = covs[16 * k + 5];
    const float tr = a + d, det = a * d - b * c, disc = tr * tr - 4 * det;
    float e0, e1;
    if (disc >= 0) { const float sq = std::sqrt(disc); e0 = (tr - sq) / 2; e1 = (tr + sq) / 2; } else { e0 = e1 = tr / 2; }
    num += static_cast<int>(std::sqrt(e0) * std::sqrt(e1));                 // :155 area of the covariance ellipse
  }
  return std::min(std::max(num, 3 * last_num_particles / 4 + 10), max_num_particles);   // :157
}

// -----------------------------------------------------------------------------
// SURVEY 8f rank 4  particle initialisation   src/particle_filter.cpp:19-84 (initializeParticles),
// src/state_particle.cpp:3-49 (the StateParticle constructor's rejection sampling), src/top_down_map.cpp:159-170
// (getClassesAtPoint), and the small host-side pieces next to it: freezeScale (:343-357), the map-centre shift of
// ParticleFilter::updateMap (:325-333)
// -----------------------------------------------------------------------------
struct OrcInitParams {      // the FilterParams fields the constructors read (state_particle.h:19-38)
  float init_pos_px_x, init_pos_px_y, init_pos_px_cov;
  float init_pos_m_x, init_pos_m_y, init_pos_deg_theta, init_pos_deg_cov;
  float fixed_scale;
};

// TopDownMap::getClassesAtPoint(Vector2i) :159-170 on the DISTANCE layers (class present <=> layer < 1); returns a bit set
ORC_API unsigned orc_classes_at_point(const float* layers, int rows, int cols, int C, float resolution, int px, int py) {
  const int cx = f2i_x86((float)px / resolution), cy = f2i_x86((float)py / resolution);   // cast<float>() / res, cast<int>()
  unsigned bits = 0;
  for (int cls = 0; cls < C; cls++)
    if (cx < cols && cy < rows && cx >= 0 && cy >= 0 && layers[(size_t)cls * rows * cols + (size_t)cx * rows + cy] < 1) bits |= 1u << cls;
  return bits;
}

namespace {
// StateParticle::StateParticle(gen, map, params, init = true)   state_particle.cpp:3-49
OrcState construct_particle(CountingMt& gen, const float* layers, int rows, int cols, int C, float resolution, const OrcInitParams& p) {
  std::uniform_real_distribution<float> uniform_dist(0., 1.);                  // :7
  std::normal_distribution<float> normal_dist(0., 1.);                         // :8
  const float map_w = (float)cols * resolution, map_h = (float)rows * resolution;   // :11 size() = (cols, rows)
  OrcState s{};
  if (p.fixed_scale < 0) s.scale = (float)std::pow(10, (uniform_dist(gen) - 0.5) * 2);   // :14-15
  else s.scale = p.fixed_scale;
  while (true) {                                                               // :20-32
    if (p.init_pos_px_x > 0) {
      s.init_x_px = std::clamp<float>(normal_dist(gen) * p.init_pos_px_cov + p.init_pos_px_x, 0, map_w);
      s.init_y_px = std::clamp<float>(normal_dist(gen) * p.init_pos_px_cov + p.init_pos_px_y, 0, map_h);
    } else {
      s.init_x_px = uniform_dist(gen) * map_w;
      s.init_y_px = uniform_dist(gen) * map_h;
    }
    // Eigen::Vector2i(float, float): the coordinates are truncated
    if (orc_classes_at_point(layers, rows, cols, C, resolution, f2i_x86(s.init_x_px), f2i_x86(s.init_y_px)) & 2u) break;   // class 1 = road
  }
  if (p.init_pos_deg_theta != std::numeric_limits<float>::infinity()) {        // :34-42
    s.theta = normal_dist(gen) * p.init_pos_deg_cov + p.init_pos_deg_theta;
    s.theta *= M_PI / 180;                                                     // float *= double
    s.have_init = 1;
  } else {
    s.theta = 0;
    s.have_init = 0;
  }
  return s;
}
}  // namespace

// ParticleFilter::initializeParticles :19-84.  Returns the number of particles written to `out` (capacity >= max_n):
// 0 when the metric initial position lies off the map or has no road within 4 px (:31-53).  px_out (may be NULL): the
// init_pos_px_x / y the filter ends up with (overwritten from the metric position at :28-29).  Every loop pass
// constructs THREE particles from the shared engine — proto_part per outer pass, `particle` and the new_particles_
// twin per inner pass — and with a free scale the first of ten copies of proto_part's state survives.
ORC_API long orc_init_particles(uint32_t seed, const float* layers, int rows, int cols, int C, float resolution, int map_center_x,
                                int map_center_y, const OrcInitParams* params, int max_n, OrcState* out, int* scale_frozen_out,
                                float* px_out, uint64_t* draws_out) {
  CountingMt gen(seed);
  if (draws_out) *draws_out = 0;
  OrcInitParams p = *params;
  size_t num_at_scale = 1;
  bool scale_frozen = false;
  if (p.fixed_scale < 0) num_at_scale = 10; else scale_frozen = true;          // :20-25
  if (scale_frozen_out) *scale_frozen_out = scale_frozen;
  if (scale_frozen && p.init_pos_m_x != std::numeric_limits<float>::infinity()) {   // :27-54
    p.init_pos_px_x = (p.init_pos_m_x * p.fixed_scale) + map_center_x;
    p.init_pos_px_y = (p.init_pos_m_y * p.fixed_scale) + map_center_y;
    if (px_out) { px_out[0] = p.init_pos_px_x; px_out[1] = p.init_pos_px_y; }
    if (p.init_pos_px_x < 0 || p.init_pos_px_x >= cols || p.init_pos_px_y < 0 || p.init_pos_px_y >= rows) return 0;
    bool good_init = false;
    for (int dx = -4; dx <= 4; dx++)
      for (int dy = -4; dy <= 4; dy++)
        if (orc_classes_at_point(layers, rows, cols, C, resolution, f2i_x86(p.init_pos_px_x + dx), f2i_x86(p.init_pos_px_y + dy)) & 2u)
          good_init = true;
    if (!good_init) return 0;
  } else if (px_out) { px_out[0] = p.init_pos_px_x; px_out[1] = p.init_pos_px_y; }
  long n = 0;
  for (int i = 0; (size_t)i < (size_t)max_n / num_at_scale; i++) {             // :58 (int < size_t comparison)
    const OrcState proto = construct_particle(gen, layers, rows, cols, C, resolution, p);
    for (float scale = 0; scale < 1; scale += 1. / num_at_scale) {             // :60 float += double
      OrcState part = construct_particle(gen, layers, rows, cols, C, resolution, p);
      if (p.fixed_scale < 0) { part = proto; part.scale = std::pow(10., scale); }   // :62-65
      out[n++] = part;
      (void)construct_particle(gen, layers, rows, cols, C, resolution, p);     // :69 the new_particles_ twin
    }
  }
  if (draws_out) *draws_out = gen.n;                                           // engine outputs consumed (to resume it later)
  return n;
}

// ParticleFilter::freezeScale :343-357: geometric mean through a float accumulator and a double pow per particle
ORC_API float orc_freeze_scale(OrcState* st, long n) {
  float geo_mean = 1;
  for (long i = 0; i < n; i++) geo_mean *= std::pow(st[i].scale, 1. / (size_t)n);
  for (long i = 0; i < n; i++) st[i].scale = geo_mean;
  return geo_mean;
}

ORC_API int orc_abi_version() { return 1; }
