// oracle/shim: pcl::PointXYZI and the two Eigen names the reference's data-type header mentions (see README.md).
#pragma once
#include <cmath>

#ifndef LMSF_SHIM_EIGEN_MATRIX4F
namespace Eigen {
// Sensor/lidar_data_type.h:74 names Eigen::Matrix4f::Identity() in a struct the extraction path never instantiates
struct Matrix4f {
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // row-major here; only (i, j) access is offered
  static Matrix4f Identity() { return Matrix4f(); }
  float& operator()(int i, int j) { return m[4 * i + j]; }
  const float& operator()(int i, int j) const { return m[4 * i + j]; }
};
}  // namespace Eigen
#endif

namespace pcl {

struct Vector3fMapShim {
  float x, y, z;
  // Eigen's norm() of a fixed-size float 3-vector: sqrt of the sequentially accumulated squares, in float
  float norm() const { return std::sqrt((x * x + y * y) + z * z); }
};

struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, pad_ = 1.0f;
  float intensity = 0;
  float pad2_[3] = {0, 0, 0};
  Vector3fMapShim getVector3fMap() const { return Vector3fMapShim{x, y, z}; }
};

}  // namespace pcl
