// opencv2/core/core.hpp — STAND-IN (test infrastructure only): the OpenCV types the compiled units name.  cv::Mat is a
// dense row-major double matrix (ParticleFilter::computeGMM fills its EM sample matrix with at<double>); the drawing
// calls of ParticleFilter::visualize are no-ops; cv::ml::EM is in opencv2/ml/ml.hpp.
#pragma once
#include <memory>
#include <vector>
#define CV_64F 6
#define CV_8UC1 0
#define CV_32FC1 5
namespace cv {
template <class T> using Ptr = std::shared_ptr<T>;
struct Size { int width = 0, height = 0; Size() {} Size(double w, double h) : width((int)w), height((int)h) {} Size operator*(int s) const { return Size(width * s, height * s); } };
struct Point { int x = 0, y = 0; Point() {} Point(double x_, double y_) : x((int)x_), y((int)y_) {}
  Point operator+(const Point& o) const { return Point(x + o.x, y + o.y); } Point operator-(const Point& o) const { return Point(x - o.x, y - o.y); } };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} double operator[](int i) const { return v[i]; } };
class Mat {
 public:
  int rows = 0, cols = 0;
  std::vector<double> d;
  Mat() {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), d((size_t)r * c, 0.0) {}
  template <class T> T& at(int r, int c) { static_assert(sizeof(T) == sizeof(double), "double matrices only"); return reinterpret_cast<T&>(d[(size_t)r * cols + c]); }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T&>(d[(size_t)r * cols + c]); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return d.empty(); }
};
inline Scalar mean(const Mat& m) { double s = 0; for (double v : m.d) s += v; return Scalar(m.d.empty() ? 0 : s / m.d.size()); }
enum { LINE_AA = 16 };
template <class... A> void circle(A&&...) {}
template <class... A> void arrowedLine(A&&...) {}
template <class... A> void ellipse(A&&...) {}
}  // namespace cv
