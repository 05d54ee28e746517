// oracle/shim: pcl::transformPointCloud(cloud, out, Matrix4f) — THIRD-PARTY ARITHMETIC, restated: PCL's dense scalar
// path, float, ((m0*x + m1*y) + m2*z) + m3 per coordinate, other fields copied (see README.md).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace pcl {

template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  if (&in != &out) out = in;
  for (std::size_t i = 0; i < in.points.size(); ++i) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3);
    out.points[i].y = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3);
    out.points[i].z = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3);
  }
}

#ifdef LMSF_SHIM_TRANSFORM_MATRIX4D
// pcl::transformPointCloud(cloud, out, Matrix4d) (LidarTrackerLocalMap.hpp:217 passes Isometry3d::matrix()): PCL computes
// in the matrix' scalar type and stores float — static_cast<float>(((m0*x + m1*y) + m2*z) + m3) with x, y, z widened to
// double (pcl/common/impl/transforms.hpp, detail::Transformer<Scalar>::se3).  Other fields copied.
template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4d& T) {
  if (&in != &out) out = in;
  for (std::size_t i = 0; i < in.points.size(); ++i) {
    const double x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = static_cast<float>(T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3));
    out.points[i].y = static_cast<float>(T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3));
    out.points[i].z = static_cast<float>(T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3));
  }
}
#endif

}  // namespace pcl
