// oracle/ref_sc.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified ScanContext class
// (Algorithm/PointClouds/processing/GlobalDescriptor/scanContext/Scancontext.hpp:59-318), compiled where it lies under
// /root/reference (oracle/Makefile, target `ref`) against oracle/shim/ (PCL as a container; Eigen's dynamic matrix as a
// container whose three reductions — mean, norm, dot — are sequential sums, see shim/Eigen/Dense).  Used by
// tests/test_oracle_sc.py to pin the oracle's restatement of row f1 against the reference's real control flow:
// polar binning, ring key, sector-key alignment, shift refinement, argmin — and, through the reference's own
// SceneRecognitionScanContext (LoopDetection/SceneRecognitionScanContext.hpp), the keyframe database: tree-rebuild
// schedule, ring-key top-10, selection and threshold of descFindSimilar.  Never linked into the product.
#include <algorithm>
#include <cmath>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include <deque>
#include <fstream>
#include <memory>
#include <unordered_map>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51 precedes every header)

// names SceneRecognitionScanContext.hpp expects from headers included before it in the node (Save / Load only)
namespace boost { namespace filesystem {
inline bool is_directory(const std::string&) { return true; }
inline bool create_directory(const std::string&) { return true; }
} }
namespace common { const std::string GREEN, RESET; }

#include "Algorithm/PointClouds/processing/GlobalDescriptor/scanContext/Scancontext.hpp"
#include "LoopDetection/SceneRecognitionScanContext.hpp"

namespace {

using Point = pcl::PointXYZI;
const int NR = 20, NS = 60;

struct Quiet {  // MakeScanContext prints its run time (TicToc::toc(string))
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(nullptr)) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

Eigen::MatrixXd to_mat(const float* d) {
  Eigen::MatrixXd m(NR, NS);
  for (int r = 0; r < NR; ++r)
    for (int s = 0; s < NS; ++s) m(r, s) = (double)d[r * NS + s];
  return m;
}

}  // namespace

extern "C" {

// MakeScanContext + MakeRingkeyFromScanContext.  desc1200: [ring][sector] (the matrix holds float heights widened to
// double, so the narrowing is exact); key20: the ring key narrowed to float as eig2stdvec does
// (SceneRecognitionScanContext.hpp:61-94).
int ref_sc_make(const float* xyzi, int n, float* desc1200, float* key20) {
  Quiet q;
  pcl::PointCloud<Point> pc;
  pc.points.resize(n);
  for (int i = 0; i < n; ++i) {
    pc.points[i].x = xyzi[4 * i];
    pc.points[i].y = xyzi[4 * i + 1];
    pc.points[i].z = xyzi[4 * i + 2];
    pc.points[i].intensity = xyzi[4 * i + 3];
  }
  Algorithm::ScanContext<Point> sc;
  Eigen::MatrixXd desc = sc.MakeScanContext(pc);
  Eigen::MatrixXd key = sc.MakeRingkeyFromScanContext(desc);
  for (int r = 0; r < NR; ++r) {
    for (int s = 0; s < NS; ++s) desc1200[r * NS + s] = (float)desc(r, s);
    key20[r] = (float)key(r, 0);
  }
  return 0;
}

// DistanceBtnScanContext
int ref_sc_distance(const float* desc_a, const float* desc_b, double* dist, int* shift) {
  Algorithm::ScanContext<Point> sc;
  std::pair<double, int> r = sc.DistanceBtnScanContext(to_mat(desc_a), to_mat(desc_b));
  *dist = r.first;
  *shift = r.second;
  return 0;
}

// SceneRecognitionScanContext (LoopDetection/SceneRecognitionScanContext.hpp:61-94, 112-124, 260-333): the keyframe
// database with its ring-key tree rebuilt every tenth keyframe over [0, size - 50).
void* ref_scdb_create() { return new Slam3D::SceneRecognitionScanContext<Point>(); }
void ref_scdb_destroy(void* h) { delete static_cast<Slam3D::SceneRecognitionScanContext<Point>*>(h); }

// AddKeyFramePoints of one cloud (container with a single entry: every entry is selected, :176-181)
int ref_scdb_add(void* h, const float* xyzi, int n) {
  Quiet q;
  auto pc = std::make_shared<pcl::PointCloud<Point>>();
  pc->points.resize(n);
  for (int i = 0; i < n; ++i) {
    pc->points[i].x = xyzi[4 * i];
    pc->points[i].y = xyzi[4 * i + 1];
    pc->points[i].z = xyzi[4 * i + 2];
    pc->points[i].intensity = xyzi[4 * i + 3];
  }
  std::unordered_map<std::string, pcl::PointCloud<Point>::ConstPtr> in;
  in["cloud"] = pc;
  static_cast<Slam3D::SceneRecognitionScanContext<Point>*>(h)->AddKeyFramePoints(in);
  return 0;
}

// LoopDetect(id): loop id or -1, and the yaw the reference puts into the relative pose (float deg2rad(shift * 6))
int ref_scdb_loop_detect(void* h, unsigned id, long long* loop_id, double* yaw_rad) {
  Quiet q;
  uint32_t i = id;
  auto r = static_cast<Slam3D::SceneRecognitionScanContext<Point>*>(h)->LoopDetect(i);
  *loop_id = r.first;
  *yaw_rad = r.second.linear().yaw;
  return 0;
}

}  // extern "C"
