// oracle/ref_tracker.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified tracker
// (LidarTracker/LidarTrackerLocalMap.hpp:30-262: Solve, RegistrationLocalMap, updateLocalMap, needUpdataLocalMap)
// instantiated as the factory does (multiLidar_loamFeatureEstimator_factory.hpp:80-95: LidarTrackerLocalMap<F,
// RegistrationBase<F>>, SetLocalMap({"loam_edge", "loam_surf"}, "sliding_Localmap"), SetRegistration(new
// CeresEdgeSurfFeatureRegistration("loam_edge", "loam_surf"))), compiled where it lies under /root/reference.
//
// What is the reference's: the prediction (constant motion / caller's delta), the call into its own registration, the
// motion increment, the keyframe test (translation norm, 2 acos(q.w) of the normalised quaternion, the time rule), the
// transform of the features with T.matrix(), the order in which the two local maps are updated and handed to the
// registration.  What is NOT: factory/Map/LocalMap_factory.hpp is absent from the reference tree — the sliding window is
// the inferred stand-in of oracle/shim_inferred (row a6'); ceres::Solve is the oracle's restated loop (as in ref_lm.cpp);
// PCL / Eigen are the container-only stand-ins of oracle/shim + oracle/shim_fixed (Transform algebra restated there).
//
// One member is overridden: RegistrationLocalMap (:168-177) is declared bool and falls off its end, which this GCC turns
// into a trap at every -O level.  FixedTracker repeats its two calls (SetInputTarget, Registration) and returns; to reach
// the private registration_ptr_ the header is included with `private` read as `protected`.  Everything else — Solve,
// updateLocalMap, needUpdataLocalMap — runs as written.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51)

#include "Common/color.hpp"  // the node includes it before the registration headers

#define LMSF_SHIM_EIGEN_MATRIX4F
#define LMSF_SHIM_TRANSFORM_MATRIX4D
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/common/transforms.h>
#include "Algorithm/PointClouds/registration/ceres_edgeSurfFeatureRegistration.hpp"
#define private protected
#include "LidarTracker/LidarTrackerLocalMap.hpp"
#undef private

namespace Slam3D { int g_inferred_window = 10; }

namespace {
using Point = pcl::PointXYZI;
using TrackerBase = Slam3D::LidarTrackerLocalMap<Point, Algorithm::RegistrationBase<Point>>;
struct Tracker : TrackerBase {
  bool RegistrationLocalMap(Slam3D::FeaturePointCloudContainer<Point> const& data, Eigen::Isometry3d& predict_pose) override {
    registration_ptr_->SetInputTarget(data);       // :174
    registration_ptr_->Registration(predict_pose);  // :175
    return true;                                    // the reference returns nothing
  }
};
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

void* ref_tracker_create(int window) {
  Quiet quiet;
  Slam3D::g_inferred_window = window;
  Tracker* t = new Tracker();
  t->SetLocalMap({"loam_edge", "loam_surf"}, "sliding_Localmap");
  std::unique_ptr<Algorithm::RegistrationBase<Point>> reg(
      new Algorithm::CeresEdgeSurfFeatureRegistration<Point>("loam_edge", "loam_surf"));
  t->SetRegistration(std::move(reg));
  return t;
}
void ref_tracker_destroy(void* h) { delete static_cast<Tracker*>(h); }

// Solve(features, stamp, deltaT).  delta_R / delta_t in: the caller's prediction (identity = constant-motion model),
// out: the motion increment.  pose_R / pose_t out: GetCurrPoseInLocalFrame().  n_map out: sizes of the two local maps.
int ref_tracker_solve(void* h, const float* edge, int ne, const float* surf, int ns, double stamp, double delta_R[9],
                      double delta_t[3], double pose_R[9], double pose_t[3], int n_map[2]) {
  Quiet quiet;
  Tracker* t = static_cast<Tracker*>(h);
  Slam3D::FeaturePointCloudContainer<Point> data;
  data["loam_edge"] = load(edge, ne);
  data["loam_surf"] = load(surf, ns);
  Eigen::Isometry3d d;
  for (int i = 0; i < 9; ++i) d.linear().d[i] = delta_R[i];
  for (int i = 0; i < 3; ++i) d.translation().d[i] = delta_t[i];
  t->Solve(data, stamp, d);
  for (int i = 0; i < 9; ++i) delta_R[i] = d.linear().d[i];
  for (int i = 0; i < 3; ++i) delta_t[i] = d.translation().d[i];
  const Eigen::Isometry3d& p = t->GetCurrPoseInLocalFrame();
  for (int i = 0; i < 9; ++i) pose_R[i] = p.linear().d[i];
  for (int i = 0; i < 3; ++i) pose_t[i] = p.translation().d[i];
  auto maps = t->GetLocalMap();
  n_map[0] = maps.count("loam_edge") ? (int)maps["loam_edge"]->size() : 0;
  n_map[1] = maps.count("loam_surf") ? (int)maps["loam_surf"]->size() : 0;
  return 0;
}

// the local map of one feature (0 edge, 1 surf) as XYZI floats; returns the number of points (out may be null)
int ref_tracker_map(void* h, int kind, float* out, int cap) {
  Tracker* t = static_cast<Tracker*>(h);
  auto maps = t->GetLocalMap();
  const std::string name = kind ? "loam_surf" : "loam_edge";
  if (!maps.count(name)) return 0;
  const auto& pc = *maps[name];
  const int n = (int)pc.size();
  if (out)
    for (int i = 0; i < n && i < cap; ++i) {
      out[4 * i] = pc.points[i].x;
      out[4 * i + 1] = pc.points[i].y;
      out[4 * i + 2] = pc.points[i].z;
      out[4 * i + 3] = pc.points[i].intensity;
    }
  return n;
}

}  // extern "C"
