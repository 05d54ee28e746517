// oracle/ref_gn.cpp — TEST INFRASTRUCTURE.  C entry point around the reference's OWN, unmodified Gauss-Newton registration
// (Algorithm/PointClouds/registration/edgeSurfFeatureRegistration.hpp:54-350: SetInputSource, SetInputTarget, Solve,
// addSurfCostFactor, addEdgeCostFactor, GNOptimization, pointAssociateToMap) with its own matchers, compiled where it
// lies under /root/reference against oracle/shim/ + oracle/shim_fixed/ (PCL as a container, KdTreeFLANN answered by the
// vendored nanoflann, Eigen coefficient by coefficient with its 3x3 / 6x6 eigen solvers, pivoted QR, inverse and
// quaternion algebra answered by oracle_math.h, products as sequential sums).  Pins rows a4.3 / a5.1 / a5.2: the loop,
// the row order, the float truncations, the degeneracy quirk, the half-angle update, the convergence test.
#include <cmath>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51)

namespace common { const std::string RED, YELLOW, GREEN, RESET; }  // Common/color.hpp, included earlier in the node

#define LMSF_SHIM_EIGEN_MATRIX4F
#include <Eigen/Dense>
#include "Algorithm/PointClouds/registration/edgeSurfFeatureRegistration.hpp"

namespace {
using Point = pcl::PointXYZI;
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

// Solve() from the prior {R (row-major 3x3), t}; the result is written back in place.  q_of_prior receives
// Quaterniond(T.rotation()) — the quaternion the reference starts from (:121) — so that the caller can hand the very same
// prior to an interface that takes quaternions.
extern "C" int ref_gn_solve(const float* map_edge, int n_me, const float* map_surf, int n_ms, const float* edge, int ne,
                            const float* surf, int ns, int max_iters, double R[9], double t[3], double q_of_prior[4]) {
  Quiet quiet;
  Algorithm::EdgeSurfFeatureRegistration<Point> reg("loam_edge", "loam_surf");
  reg.SetInputSource(std::make_pair(std::string("loam_edge"), pcl::PointCloud<Point>::ConstPtr(load(map_edge, n_me))));
  reg.SetInputSource(std::make_pair(std::string("loam_surf"), pcl::PointCloud<Point>::ConstPtr(load(map_surf, n_ms))));
  Slam3D::FeaturePointCloudContainer<Point> scan;
  scan["loam_edge"] = load(edge, ne);
  scan["loam_surf"] = load(surf, ns);
  reg.SetInputTarget(scan);
  reg.SetMaxIteration((uint16_t)max_iters);
  Eigen::Isometry3d T;
  for (int i = 0; i < 9; ++i) T.linear().d[i] = R[i];
  for (int i = 0; i < 3; ++i) T.translation().d[i] = t[i];
  Eigen::Quaterniond q0(T.rotation());
  q_of_prior[0] = q0.x();
  q_of_prior[1] = q0.y();
  q_of_prior[2] = q0.z();
  q_of_prior[3] = q0.w();
  reg.Solve(T);
  for (int i = 0; i < 9; ++i) R[i] = T.linear().d[i];
  for (int i = 0; i < 3; ++i) t[i] = T.translation().d[i];
  return 0;
}
