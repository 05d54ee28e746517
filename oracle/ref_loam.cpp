// oracle/ref_loam.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified
// LOAMFeatureProcessorBase (Algorithm/PointClouds/processing/FeatureExtract/LOAMFeatureProcessor_base.hpp:59-343),
// PointCloudCommonProcess (Algorithm/PointClouds/processing/common_processing.hpp:87-111) and DistanceFilter
// (Filter/distance_filter.hpp:24-44), compiled where they lie under /root/reference (oracle/Makefile, target `ref`)
// against the container-only PCL stand-ins of oracle/shim/.  Used by tests/test_oracle.py to pin the oracle's
// restatement of rows a1.* and f4 against the reference's real code.  Never linked into the product.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <iterator>
#include <memory>
#include <string>
#include <utility>
#include <vector>

using namespace std;  // the reference's headers use unqualified `string` / `make_pair` (an earlier using-directive)

#include "Algorithm/PointClouds/processing/FeatureExtract/LOAMFeatureProcessor_base.hpp"
#include "Algorithm/PointClouds/processing/Preprocess/RotaryLidar_preprocessing.hpp"
#include <pcl/filters/filter.h>  // removeNaNFromPointCloud (the stand-in of oracle/shim: x, y, z finite, order kept)

namespace {

using Point = pcl::PointXYZI;

void load(const float* xyzi, int n, pcl::PointCloud<Point>& pc) {
  pc.points.resize(n);
  for (int i = 0; i < n; ++i) {
    pc.points[i].x = xyzi[4 * i + 0];
    pc.points[i].y = xyzi[4 * i + 1];
    pc.points[i].z = xyzi[4 * i + 2];
    pc.points[i].intensity = xyzi[4 * i + 3];
  }
  pc.width = n;
  pc.height = 1;
}

int store(const pcl::PointCloud<Point>& pc, float* out, int cap) {
  int n = (int)pc.points.size();
  if (n > cap) return -1;
  for (int i = 0; i < n; ++i) {
    out[4 * i + 0] = pc.points[i].x;
    out[4 * i + 1] = pc.points[i].y;
    out[4 * i + 2] = pc.points[i].z;
    out[4 * i + 3] = pc.points[i].intensity;
  }
  return n;
}

struct Quiet {  // the reference's constructor prints its parameters
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(nullptr)) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

// LOAMFeatureProcessorBase::Process.  edge_out / surf_out: caller-owned, capacity `cap` points each.
// Returns 0, or -1 when an output does not fit.
int ref_loam_extract(const float* xyzi, int n, int n_scans, float min_range, float max_range, float edge_thresh,
                     int remove_bad_points, int cap, float* edge_out, int* n_edge, float* surf_out, int* n_surf) {
  Quiet q;
  Algorithm::LOAMFeatureProcessorBase<Point, Point> proc((uint16_t)n_scans, min_range, max_range, edge_thresh, 0.1f,
                                                         remove_bad_points != 0);
  Slam3D::LidarData<Point> in;
  load(xyzi, n, in.point_cloud);
  Slam3D::CloudContainer<Point> out;
  proc.Process(in, out);
  int ne = store(*out.pointcloud_data_.at("loam_edge"), edge_out, cap);
  int ns = store(*out.pointcloud_data_.at("loam_surf"), surf_out, cap);
  if (ne < 0 || ns < 0) return -1;
  *n_edge = ne;
  *n_surf = ns;
  return 0;
}

// PointCloudCommonProcess::Process with removeNaN and the distance filter (no VoxelGrid / outlier filter set: PCL's
// arithmetic is not available; FilterBase::Filter then returns a copy, filter_base.hpp:39-40).
int ref_common_process(const float* xyzi, int n, int remove_nan, float dist_near, float dist_far, int cap, float* out,
                       int* n_out) {
  Algorithm::PointCloudCommonProcess<Point> proc("processed", remove_nan != 0);
  proc.SetDistanceFilter(dist_near, dist_far);
  Slam3D::LidarData<Point> in;
  load(xyzi, n, in.point_cloud);
  in.point_cloud.is_dense = false;  // a driver cloud that may hold NaN returns (MultiLidarSLAM_node.cpp:126-133)
  Slam3D::CloudContainer<Point> res;
  proc.Process(in, res);
  int m = store(*res.pointcloud_data_.at("processed"), out, cap);
  if (m < 0) return -1;
  *n_out = m;
  return 0;
}


// What the node does to a raw sweep before the estimator sees it: getCloudFromMsg's pcl::removeNaNFromPointCloud
// (src/apps/src/MultiLidarSLAM_node.cpp:126-133), then RotaryLidarPreProcess<PointXYZI>::Process
// (Preprocess/RotaryLidar_preprocessing.hpp:31-104).  Returns the number of points written (cap: capacity of out).
int ref_rotary_preprocess(const float* xyzi, int n, double scan_period, int cap, float* out) {
  Slam3D::LidarData<Point> data;
  load(xyzi, n, data.point_cloud);
  data.point_cloud.is_dense = false;  // as a driver publishes a cloud that may hold NaN returns (dense clouds are copied)
  std::vector<int> indices;
  pcl::removeNaNFromPointCloud(data.point_cloud, data.point_cloud, indices);
  if (data.point_cloud.points.empty()) return 0;  // the reference would index points[0] of an empty cloud
  Algorithm::RotaryLidarPreProcess<Point> pre(scan_period);
  pre.Process(data);
  return store(data.point_cloud, out, cap);
}
}  // extern "C"
