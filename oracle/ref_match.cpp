// oracle/ref_match.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified feature matchers
// (Algorithm/PointClouds/registration/FeatureMatch/{FeatureMatchBase.hpp:21-45, EdgeFeatureMatch.hpp:17-88,
// surfFeatureMatch.hpp:17-89}), compiled where they lie under /root/reference against oracle/shim/ (PCL as a container,
// KdTreeFLANN answered by the reference's vendored nanoflann) and oracle/shim_fixed/ (Eigen's fixed-size matrices
// coefficient by coefficient; its eigen solver and QR answered by the oracle's restatements, oracle_math.h).
// Pins rows a4.1 / a4.2: control flow, types (the `float distance`, the fp32 search threshold) and expression order.
#include <cmath>
#include <memory>
#include <string>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51)

#define LMSF_SHIM_EIGEN_MATRIX4F  // Matrix4f comes from shim_fixed/Eigen/Dense in this translation unit
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/kdtree/kdtree_flann.h>
#include "Algorithm/PointClouds/registration/FeatureMatch/EdgeFeatureMatch.hpp"
#include "Algorithm/PointClouds/registration/FeatureMatch/surfFeatureMatch.hpp"

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
}  // namespace

// out10 per query, the layout of lmsf_match: edge {n(3), residual, a(3), b(3)}; surf {n(3), residual, D, 0...}
extern "C" int ref_match(int kind, const float* map_xyzi, int n_map, const float* q_xyz, int nq, unsigned char* ok,
                         double* out10) {
  auto map = load(map_xyzi, n_map);
  Algorithm::EdgeFeatureMatch<Point> em;
  Algorithm::SurfFeatureMatch<Point> sm;
  if (kind == 0)
    em.SetSearchTarget(map);
  else
    sm.SetSearchTarget(map);
  for (int i = 0; i < nq; ++i) {
    Point p;
    p.x = q_xyz[3 * i];
    p.y = q_xyz[3 * i + 1];
    p.z = q_xyz[3 * i + 2];
    double* o = out10 + 10 * i;
    for (int k = 0; k < 10; ++k) o[k] = 0.0;
    if (kind == 0) {
      Algorithm::EdgeFeatureMatch<Point>::EdgeCostFactorInfo r;
      ok[i] = em.Match(p, r) ? 1 : 0;
      if (ok[i]) {
        for (int k = 0; k < 3; ++k) {
          o[k] = r.norm_[k];
          o[4 + k] = r.points_set_[0][k];
          o[7 + k] = r.points_set_[1][k];
        }
        o[3] = r.residuals_;
      }
    } else {
      Algorithm::SurfFeatureMatch<Point>::SurfCostFactorInfo r;
      ok[i] = sm.Match(p, r) ? 1 : 0;
      if (ok[i]) {
        for (int k = 0; k < 3; ++k) o[k] = r.norm_[k];
        o[3] = r.residuals_;
        o[4] = r.D_;
      }
    }
  }
  return 0;
}
