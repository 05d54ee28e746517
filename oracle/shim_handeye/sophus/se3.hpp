// oracle/shim_handeye: TEST INFRASTRUCTURE.  Common/pose.hpp includes Sophus for one member (Pose::se3(), the se(3)
// logarithm) that the hand-eye calibration never calls; Sophus is absent here, so the type exists and the member traps.
#pragma once
#include <eigen3/Eigen/Dense>
namespace Sophus {
struct SE3d {
  SE3d(const Eigen::Quaterniond&, const Eigen::Vector3d&) {}
  Eigen::Matrix<double, 6, 1> log() const { __builtin_trap(); }
};
}  // namespace Sophus
