// reference_stubs.hpp — minimal stand-ins for the third-party and reference types the adapters touch,
// for building lmsf_b200_adapters.hpp where PCL / Eigen / the reference tree are absent (this repo's CI,
// tests/cpp).  In the reference tree define LMSF_WITH_REFERENCE and include the real headers instead:
// nothing here is used then.  Each declaration mirrors, member for member, only what the adapters use:
//   pcl::PointXYZI, pcl::PointCloud<T>                       (PCL)
//   Eigen::Isometry3d  (linear(), translation(), matrix())   (Eigen)
//   Slam3D::LidarData, CloudContainer, FeaturePointCloudContainer   include/Sensor/lidar_data_type.h:29-63
//   Algorithm::PointCloudProcessBase      include/Algorithm/PointClouds/processing/process_base.hpp:25-39
//   Algorithm::FilterBase                 include/Algorithm/PointClouds/processing/Filter/filter_base.hpp:24-54
//   Algorithm::RegistrationBase           include/Algorithm/PointClouds/registration/registration_base.hpp:24-34
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZI {
  float x, y, z, pad_;
  float intensity;
  float pad2_[3];
};
struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};
template <typename PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; }
  void resize(std::size_t n) { points.resize(n); width = (std::uint32_t)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
};
}  // namespace pcl

namespace Eigen {
// just enough of Isometry3d: a row-major 3x3 rotation and a translation
struct Matrix3dLite {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double& operator()(int r, int c) { return m[r * 3 + c]; }
  double operator()(int r, int c) const { return m[r * 3 + c]; }
};
struct Vector3dLite {
  double v[3] = {0, 0, 0};
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct Matrix4f {  // row/column access only: what AlignmentScore's relpose argument needs
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float& operator()(int r, int c) { return m[r * 4 + c]; }
  float operator()(int r, int c) const { return m[r * 4 + c]; }
  static Matrix4f Identity() { return Matrix4f(); }
};
class Isometry3d {
 public:
  static Isometry3d Identity() { return Isometry3d(); }
  Matrix3dLite& linear() { return R_; }
  const Matrix3dLite& linear() const { return R_; }
  Vector3dLite& translation() { return t_; }
  const Vector3dLite& translation() const { return t_; }
 private:
  Matrix3dLite R_;
  Vector3dLite t_;
};
}  // namespace Eigen

namespace Slam3D {
template <typename _PointT>
struct LidarData {
  pcl::PointCloud<_PointT> point_cloud;
};
template <typename _PointType>
using FeaturePointCloudContainer = std::unordered_map<std::string, typename pcl::PointCloud<_PointType>::ConstPtr>;
template <typename _FeatureT>
struct CloudContainer {
  double time_stamp_ = 0;
  FeaturePointCloudContainer<_FeatureT> pointcloud_data_;
};
}  // namespace Slam3D

namespace Algorithm {
using Slam3D::CloudContainer;
using Slam3D::FeaturePointCloudContainer;
using Slam3D::LidarData;

template <typename _InPointT, typename _OutPointT>
class PointCloudProcessBase {
 public:
  PointCloudProcessBase() {}
  virtual ~PointCloudProcessBase() {}
  virtual void Process(LidarData<_InPointT> const& data_in, CloudContainer<_OutPointT>& data_out) = 0;
};

template <typename _PointType>
using PointCloudPtr = typename pcl::PointCloud<_PointType>::Ptr;
template <typename _PointType>
using PointCloudConstPtr = typename pcl::PointCloud<_PointType>::ConstPtr;

template <typename _PointType>
class FilterBase {
 public:
  FilterBase() {}
  virtual ~FilterBase() {}
  virtual PointCloudPtr<_PointType> Filter(const PointCloudConstPtr<_PointType>& cloud_in) const {
    return PointCloudPtr<_PointType>(new pcl::PointCloud<_PointType>(*cloud_in));  // null filter: a copy (:39-40)
  }
};

template <typename _PointType>
class RegistrationBase {
 public:
  using PointCloudConstPtr = typename pcl::PointCloud<_PointType>::ConstPtr;
  using SourceInput = std::pair<std::string, PointCloudConstPtr>;
  virtual ~RegistrationBase() {}
  virtual void SetInputSource(SourceInput const& source_input) = 0;
  virtual void SetInputTarget(FeaturePointCloudContainer<_PointType> const& target_input) = 0;
  virtual void Solve(Eigen::Isometry3d& T) = 0;
};
}  // namespace Algorithm
