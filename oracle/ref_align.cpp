// oracle/ref_align.cpp — TEST INFRASTRUCTURE.  C entry point around the reference's OWN, unmodified
// PointCloudAlignmentEvaluate (Algorithm/PointClouds/registration/alignEvaluate.hpp:42-87), compiled where it lies under
// /root/reference against oracle/shim/: pcl::KdTreeFLANN answered by the reference's vendored nanoflann (exact, fp32
// L2_Simple), pcl::transformPointCloud restated (PCL's scalar float path).  Pins the control flow of row f2 — inlier
// test, accumulation order, overlap ratio, the two return branches — against the reference's real code.
#include <cassert>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51)

#include "Sensor/lidar_data_type.h"
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/common/transforms.h>
#include "Algorithm/PointClouds/registration/alignEvaluate.hpp"

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

// SetTargetPoints(target) + AlignmentScore(cloud, relpose (row-major 4x4), inlier_thresh, inlier_ratio_thresh)
extern "C" int ref_align_score(const float* target, int n_target, const float* cloud, int n, const float T[16],
                               double inlier_thresh, double inlier_ratio_thresh, double* score, double* overlap) {
  Slam3D::PointCloudAlignmentEvaluate<Point> ev;
  ev.SetTargetPoints(load(target, n_target));
  Eigen::Matrix4f M;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) M(i, j) = T[4 * i + j];
  std::pair<double, double> r = ev.AlignmentScore(load(cloud, n), M, inlier_thresh, inlier_ratio_thresh);
  *score = r.first;
  *overlap = r.second;
  return 0;
}
