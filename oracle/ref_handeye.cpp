// oracle/ref_handeye.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified hand-eye
// calibration class (Algorithm/calibration/handeye_calibration_base.hpp:36-246) and pose class (Common/pose.hpp),
// compiled where they lie under /root/reference against oracle/shim_fixed/ (Eigen coefficient by coefficient; the
// quaternion algebra answered by oracle_math.h, AngleAxisd(q) and JacobiSVD by restatements of their published
// definitions) and oracle/shim_handeye/ (an empty Sophus).  Pins row f3's algebra and bookkeeping: the screw-motion
// gate (EPSILON_R / EPSILON_T), the accumulation of sub-threshold motions, the 300-pair store with its replacement
// rule (smallest rotation first), the (L(q_primary) - R(q_sub)) blocks, the sign convention of the null vector, the
// rot_cov[2] > 0.25 gate, the translation system.  NOT pinned: Eigen's own SVD arithmetic (absent).
#include <cmath>
#include <cstdint>
#include <iostream>
#include <queue>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51): the header calls abs(double)
                      // and make_pair unqualified

#include "Algorithm/calibration/handeye_calibration_base.hpp"

namespace {
struct Quiet {  // the class logs every call to std::cout
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
Slam3D::Pose pose_of(const double p[7]) {  // {qx, qy, qz, qw, tx, ty, tz}; the constructor normalises q
  return Slam3D::Pose(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6]));
}
}  // namespace

extern "C" void* ref_handeye_create() { return new Algorithm::HandEyeCalibrationBase(); }
extern "C" void ref_handeye_destroy(void* h) { delete static_cast<Algorithm::HandEyeCalibrationBase*>(h); }

// AddPose (:73-108): 1 = the motion pair was stored and at least three pairs are held
extern "C" int ref_handeye_add_pose(void* h, const double delta_primary[7], const double delta_sub[7]) {
  Quiet q;
  return static_cast<Algorithm::HandEyeCalibrationBase*>(h)->AddPose(pose_of(delta_primary), pose_of(delta_sub)) ? 1 : 0;
}

// CalibExRotation (:115-153) && CalibExTranslation (:155-190), the order MultiLidarSystem::process() calls them in
// (System/ML_System.hpp:263-270); R9 row-major rotation, t3 translation of GetCalibResult (:193-202)
extern "C" int ref_handeye_calibrate(void* h, double R9[9], double t3[3]) {
  Quiet q;
  auto* c = static_cast<Algorithm::HandEyeCalibrationBase*>(h);
  if (!c->CalibExRotation()) return 0;
  if (!c->CalibExTranslation()) return 0;
  Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
  if (!c->GetCalibResult(T)) return 0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R9[i * 3 + j] = T.linear()(i, j);
    t3[i] = T.translation()[i];
  }
  return 1;
}
