// oracle/ref_lm.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified factory-default
// registration (Algorithm/PointClouds/registration/ceres_edgeSurfFeatureRegistration.hpp:28-244) with its own matchers,
// cost functions and SE3 parameterization, compiled where it lies under /root/reference against oracle/shim/ +
// oracle/shim_fixed/.  ceres::Solve is answered by the oracle's restated trust-region loop (shim_fixed/ceres/ceres.h)
// running over the reference's own residual blocks.  Pins row a5.3: the stateful outer budget, the per-outer problem
// construction (block order, Huber 0.1, 4 inner iterations, DENSE_QR), re-matching, pointAssociateToMap, pose in / out.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51)

namespace common { const std::string RED, YELLOW, GREEN, RESET; }

#define LMSF_SHIM_EIGEN_MATRIX4F
#include <Eigen/Dense>
#include "Algorithm/PointClouds/registration/ceres_edgeSurfFeatureRegistration.hpp"

namespace {
using Point = pcl::PointXYZI;
using Reg = Algorithm::CeresEdgeSurfFeatureRegistration<Point>;
pcl::PointCloud<Point>::Ptr load(const float* xyzi, int n) {
  auto pc = std::make_shared<pcl::PointCloud<Point>>();
  pc->points.resize(n);
  for (int i = 0; i < n; ++i) {
    pc->points[i].x = xyzi[4 * i];
    pc->points[i].y = xyzi[4 * i + 1];
    pc->points[i].z = xyzi[4 * i + 2];
    pc->points[i].intensity = xyzi[4 * i + 3];
  }
  return pc;
}
struct Quiet {
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(nullptr)) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// one registration object = one LiDAR's solver: its outer budget is state that lives across Solve() calls (:46,100-101)
void* ref_lm_create() { return new Reg("loam_edge", "loam_surf"); }
void ref_lm_destroy(void* h) { delete static_cast<Reg*>(h); }

int ref_lm_set_map(void* h, int kind, const float* xyzi, int n) {
  static_cast<Reg*>(h)->SetInputSource(
      std::make_pair(std::string(kind ? "loam_surf" : "loam_edge"), pcl::PointCloud<Point>::ConstPtr(load(xyzi, n))));
  return 0;
}

// SetInputTarget + Solve from the prior {R row-major, t}; result in place; q_of_prior = Quaterniond(T.rotation())
int ref_lm_solve(void* h, const float* edge, int ne, const float* surf, int ns, double R[9], double t[3],
                 double q_of_prior[4]) {
  Quiet quiet;
  Reg* reg = static_cast<Reg*>(h);
  Slam3D::FeaturePointCloudContainer<Point> scan;
  scan["loam_edge"] = load(edge, ne);
  scan["loam_surf"] = load(surf, ns);
  reg->SetInputTarget(scan);
  Eigen::Isometry3d T;
  for (int i = 0; i < 9; ++i) T.linear().d[i] = R[i];
  for (int i = 0; i < 3; ++i) T.translation().d[i] = t[i];
  Eigen::Quaterniond q0(T.rotation());
  q_of_prior[0] = q0.x();
  q_of_prior[1] = q0.y();
  q_of_prior[2] = q0.z();
  q_of_prior[3] = q0.w();
  reg->Solve(T);
  for (int i = 0; i < 9; ++i) R[i] = T.linear().d[i];
  for (int i = 0; i < 3; ++i) t[i] = T.translation().d[i];
  return 0;
}

}  // extern "C"
