// opencv2/core/core.hpp — STAND-IN (test infrastructure only) for the slice of OpenCV the compiled reference units use.
// cv::Mat is a dense row-major matrix of uint8 / float / double, owning its data or wrapping the caller's (the reference
// wraps Eigen buffers, top_down_map.cpp:298-311).  The routines follow OpenCV's documented element-wise semantics:
//   convertTo to 8U: saturate_cast<uchar> = round half to even, clamped;  to 32F with a factor: float multiply;
//   threshold BINARY / TRUNC;  setTo under a mask;  Mat *= s.
// cv::distanceTransform(DIST_L2, DIST_MASK_PRECISE) is the THIRD-PARTY algorithm of row a4: here an exact squared
// Euclidean transform (column scan + exhaustive row minimum, integers) followed by sqrtf — the values OpenCV's own
// trueDistTrans produces, which tests/test_oracle.py establishes against the real cv2 (fixture + live).  With no zero
// pixel at all OpenCV returns a huge sentinel; so does this (the caller truncates at 50).
// imread / imwrite carry 8-bit gray PNG files through the codec of top_down_renderer_b200/host/png_gray.hpp (link -lz),
// itself checked against cv2 in tests/test_host_math.py.  Drawing calls are no-ops; cv::ml::EM is in opencv2/ml/ml.hpp.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "../../../../top_down_renderer_b200/host/png_gray.hpp"
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_64F 6
namespace cv {
template <class T> using Ptr = std::shared_ptr<T>;
struct Size { int width = 0, height = 0; Size() {} Size(double w, double h) : width((int)w), height((int)h) {} Size operator*(int s) const { return Size(width * s, height * s); } };
struct Point { int x = 0, y = 0; Point() {} Point(double x_, double y_) : x((int)x_), y((int)y_) {}
  Point operator+(const Point& o) const { return Point(x + o.x, y + o.y); } Point operator-(const Point& o) const { return Point(x - o.x, y - o.y); } };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} double operator[](int i) const { return v[i]; } };
inline size_t elem_size(int type) { return type == CV_8UC1 ? 1 : type == CV_32FC1 ? 4 : 8; }
class Mat {
 public:
  static size_t elem_size_of(int type) { return elem_size(type); }
  int rows = 0, cols = 0, type_ = CV_8UC1;
  uint8_t* data = nullptr;
  size_t step = 0;                                                       // bytes per row (always dense here)
  std::shared_ptr<std::vector<uint8_t>> own;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* external) : rows(r), cols(c), type_(type), data(static_cast<uint8_t*>(external)), step((size_t)c * elem_size_of(type)) {}
  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && type_ == type) return;       // like cv::Mat::create: keeps a fitting buffer
    rows = r; cols = c; type_ = type; step = (size_t)c * elem_size_of(type);
    own = std::make_shared<std::vector<uint8_t>>((size_t)r * c * elem_size(type), 0);
    data = own->data();
  }
  int type() const { return type_; }
  size_t total() const { return (size_t)rows * cols; }
  template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data)[(size_t)r * cols + c]; }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == nullptr || total() == 0; }
  double get(size_t i) const { return type_ == CV_8UC1 ? data[i] : type_ == CV_32FC1 ? reinterpret_cast<const float*>(data)[i] : reinterpret_cast<const double*>(data)[i]; }
  static uint8_t sat_u8(double v) { const double r = std::nearbyint(v); return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r); }
  void convertTo(Mat& dst, int type, double alpha = 1) const {
    Mat out; out.create(rows, cols, type);
    for (size_t i = 0; i < total(); i++) {
      if (type == CV_8UC1) out.data[i] = type_ == CV_32FC1 ? sat_u8((double)(reinterpret_cast<const float*>(data)[i] * (float)alpha)) : sat_u8(get(i) * alpha);
      else if (type == CV_32FC1) reinterpret_cast<float*>(out.data)[i] = (float)get(i) * (float)alpha;
      else reinterpret_cast<double*>(out.data)[i] = get(i) * alpha;
    }
    dst = out;
  }
  Mat& operator*=(double s) {
    if (type_ == CV_32FC1) for (size_t i = 0; i < total(); i++) reinterpret_cast<float*>(data)[i] = reinterpret_cast<float*>(data)[i] * (float)s;
    else if (type_ == CV_64F) for (size_t i = 0; i < total(); i++) reinterpret_cast<double*>(data)[i] *= s;
    else for (size_t i = 0; i < total(); i++) data[i] = sat_u8(data[i] * s);
    return *this;
  }
  void setTo(double v, const Mat& mask) {
    for (size_t i = 0; i < total(); i++) if (mask.data[i]) {
      if (type_ == CV_32FC1) reinterpret_cast<float*>(data)[i] = (float)v; else if (type_ == CV_64F) reinterpret_cast<double*>(data)[i] = v; else data[i] = sat_u8(v);
    }
  }
};
inline Scalar mean(const Mat& m) { double s = 0; for (size_t i = 0; i < m.total(); i++) s += m.get(i); return Scalar(m.total() ? s / m.total() : 0); }
enum { THRESH_BINARY = 0, THRESH_TRUNC = 2, DIST_L2 = 2, DIST_MASK_PRECISE = 0, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1, LINE_AA = 16 };
inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
  Mat out = (&src == &dst) ? dst : Mat(src.rows, src.cols, src.type());
  for (size_t i = 0; i < src.total(); i++) {
    if (src.type() == CV_8UC1) { const uint8_t v = src.data[i]; out.data[i] = type == THRESH_BINARY ? (v > (int)thresh ? Mat::sat_u8(maxval) : 0) : (v > (int)thresh ? Mat::sat_u8(thresh) : v); }
    else { const float v = reinterpret_cast<const float*>(src.data)[i]; reinterpret_cast<float*>(out.data)[i] = type == THRESH_BINARY ? (v > (float)thresh ? (float)maxval : 0.f) : (v > (float)thresh ? (float)thresh : v); }
  }
  dst = out;
  return thresh;
}
// exact Euclidean distance to the nearest zero pixel (see the header comment)
inline void distanceTransform(const Mat& src, Mat& dst, int, int) {
  const int R = src.rows, C = src.cols;
  const int64_t INF = (int64_t)1 << 40;
  std::vector<int64_t> g((size_t)R * C, INF);           // squared vertical distance to the nearest zero in the column
  for (int c = 0; c < C; c++) {
    int64_t last = -1;
    for (int r = 0; r < R; r++) { if (src.data[(size_t)r * C + c] == 0) last = r; if (last >= 0) g[(size_t)r * C + c] = (r - last) * (r - last); }
    last = -1;
    for (int r = R - 1; r >= 0; r--) { if (src.data[(size_t)r * C + c] == 0) last = r; if (last >= 0) g[(size_t)r * C + c] = std::min<int64_t>(g[(size_t)r * C + c], (last - r) * (last - r)); }
  }
  if (!(dst.data && dst.rows == R && dst.cols == C && dst.type() == CV_32FC1)) dst.create(R, C, CV_32FC1);
  float* out = reinterpret_cast<float*>(dst.data);
  for (int r = 0; r < R; r++)
    for (int c = 0; c < C; c++) {
      int64_t best = INF;
      for (int k = 0; k < C; k++) { const int64_t gk = g[(size_t)r * C + k]; if (gk < INF) best = std::min<int64_t>(best, gk + (int64_t)(c - k) * (c - k)); }
      out[(size_t)r * C + c] = best >= INF ? 3.0e38f : sqrtf((float)best);
    }
}
inline void flip(const Mat& src, Mat& dst, int /*0 = around the x axis*/) {
  Mat out(src.rows, src.cols, src.type());
  const size_t line = (size_t)src.cols * elem_size(src.type());
  for (int r = 0; r < src.rows; r++) std::memcpy(out.data + (size_t)(src.rows - 1 - r) * line, src.data + (size_t)r * line, line);
  dst = out;
}
inline bool imwrite(const std::string& path, const Mat& m) { return m.type() == CV_8UC1 && tdrhost::png::write_gray(path, m.data, m.cols, m.rows); }
inline Mat imread(const std::string& path, int flags = IMREAD_COLOR) {
  Mat m;
  if (flags != IMREAD_GRAYSCALE) return m;               // colour maps are not part of what this build drives
  std::vector<uint8_t> img; int w = 0, h = 0;
  if (!tdrhost::png::read_gray(path, img, w, h)) return m;
  m.create(h, w, CV_8UC1);
  std::memcpy(m.data, img.data(), img.size());
  return m;
}
// Eigen (column-major) <-> cv::Mat (row-major), same logical (row, col)
template <class M> void eigen2cv(const M& src, Mat& dst) {
  dst.create((int)src.rows(), (int)src.cols(), CV_32FC1);
  for (int r = 0; r < dst.rows; r++) for (int c = 0; c < dst.cols; c++) dst.at<float>(r, c) = src(r, c);
}
template <class M> void cv2eigen(const Mat& src, M& dst) {
  dst.resize(src.rows, src.cols);
  for (int r = 0; r < src.rows; r++) for (int c = 0; c < src.cols; c++) dst(r, c) = src.at<float>(r, c);
}
template <class... A> void circle(A&&...) {}
template <class... A> void arrowedLine(A&&...) {}
template <class... A> void ellipse(A&&...) {}
}  // namespace cv
