// top_down_map_base.cpp — TEST INFRASTRUCTURE ONLY.  src/top_down_map.cpp of the reference cannot be compiled here: it
// is OpenCV (imread, distanceTransform, threshold, eigen2cv), nanosvg and Eigen expression templates (Map with
// strides, LinSpaced, replicate) throughout.  The units that ARE compiled from the reference (top_down_map_polar.cpp,
// state_particle.cpp, particle_filter.cpp, active_localizer.cpp) need a few TopDownMap members at link time; they are
// RESTATED here, each a few lines, with the reference lines they follow.  The map layers themselves (distance fields,
// mask, polar offset table) are installed by the harness from the oracle's outputs — inputs of this build, not results.
#include "top_down_render/top_down_map.h"

// src/top_down_map.cpp:9-16, the dynamic-map branch (map_path == "")
TopDownMap::TopDownMap(const TopDownMap::Params& params) {
  params_ = params;
  map_center_ = Eigen::Vector2i::Zero();
  have_map_ = false;
}
// :146-157 without the image: the harness installs the layers
void TopDownMap::updateMap(const cv::Mat&, const Eigen::Vector2i& map_center) { map_center_ = map_center; }
// :159-170
void TopDownMap::getClassesAtPoint(const Eigen::Vector2i& center_ind, std::vector<int>& classes) {
  Eigen::Vector2i center = (center_ind.cast<float>() / params_.resolution).cast<int>();
  classes.clear();
  for (int cls = 0; cls < params_.num_classes; cls++) {
    if (center[0] < class_maps_[cls].cols() && center[1] < class_maps_[cls].rows() && center[0] >= 0 && center[1] >= 0) {
      if (class_maps_[cls](center[1], center[0]) < 1) classes.push_back(cls);
    }
  }
}
// :172-175
void TopDownMap::getClassesAtPoint(const Eigen::Vector2f& center, std::vector<int>& classes) {
  Eigen::Vector2i center_ind = (center / params_.resolution).cast<int>();
  getClassesAtPoint(center_ind, classes);
}
int TopDownMap::numClasses() const { return params_.num_classes; }                                                  // :177-179
Eigen::Vector2i TopDownMap::size() const { return Eigen::Vector2i(class_maps_[0].cols(), class_maps_[0].rows()); }  // :181-183
Eigen::Vector2i TopDownMap::mapCenter() const { return map_center_; }                                               // :185-187
float TopDownMap::resolution() const { return params_.resolution; }                                                 // :189-191
bool TopDownMap::haveMap() const { return have_map_; }                                                              // :193-195
// :367-389 is Eigen expression-template code; TopDownMapPolar's constructor calls it for the offset table, which the
// harness then overwrites with the oracle's table (an INPUT of the device library too)
void TopDownMap::samplePts(Eigen::Vector2f, float, Eigen::Array2Xf& pts, int, int, float) { pts.setZero(); }
