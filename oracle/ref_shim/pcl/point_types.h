// pcl/point_types.h — STAND-IN (test infrastructure only): pcl::PointXYZI's layout (32 bytes, intensity at byte 16) and
// the part of pcl::PointCloud the reference's renderers touch (points, width, height, at(column, row)).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <Eigen/Dense>
#define PCL_ADD_POINT4D union EIGEN_ALIGN16 { float data[4]; struct { float x; float y; float z; }; };
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)
namespace pcl {
struct EIGEN_ALIGN16 PointXYZI {
  PCL_ADD_POINT4D
  union { struct { float intensity; }; float data_c[4]; };
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");
template <class PointT> class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  const PointT& at(int column, int row) const { return points.at((size_t)row * width + column); }
  PointT& at(int column, int row) { return points.at((size_t)row * width + column); }
};
}  // namespace pcl
