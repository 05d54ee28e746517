// lmsf_b200_adapters.hpp — header-only C++ adapters that put the C ABI (lmsf_b200.h) behind the
// reference's own three abstract interfaces, so its factories can instantiate the B200 path unchanged:
//
//   lmsf::CudaLoamFeatureProcessor<In,Out> : Algorithm::PointCloudProcessBase<In,Out>
//        replaces  LOAMFeatureProcessorBase   (FeatureExtract/LOAMFeatureProcessor_base.hpp:26, built at
//                                              factory/System/ML_SystemFactory.hpp:197)
//   lmsf::CudaVoxelGridFilter<P>           : Algorithm::FilterBase<P>
//        replaces  VoxelGridFilter            (Filter/voxel_grid.hpp:19, make_voxelGrid filter_factory.hpp:36-41)
//   lmsf::CudaEdgeSurfRegistration<P>      : Algorithm::RegistrationBase<P>
//        replaces  CeresEdgeSurfFeatureRegistration / EdgeSurfFeatureRegistration
//                                             (registration/ceres_edgeSurfFeatureRegistration.hpp:26,
//                                              edgeSurfFeatureRegistration.hpp:27, built at ML_SystemFactory.hpp:189-190)
//   lmsf::CudaPointCloudCommonProcess<P>   : Algorithm::PointCloudProcessBase<P,P>
//        replaces  PointCloudCommonProcess    (processing/common_processing.hpp:28, removeNaN + VoxelGrid + DistanceFilter)
//   lmsf::CudaMultiLidarExtrinsics             the calibration branch of MultiLidarSystem::process()
//                                             (System/ML_System.hpp:239-323) + HandEyeCalibrationBase
//   lmsf::CudaPointCloudAlignmentEvaluate<P>   (no abstract base in the reference: same public methods)
//        replaces  Slam3D::PointCloudAlignmentEvaluate   (registration/alignEvaluate.hpp:23-98)
//   lmsf::CudaSceneRecognitionScanContext<P>   (no abstract base in the reference: same public methods)
//        replaces  Slam3D::SceneRecognitionScanContext   (LoopDetection/SceneRecognitionScanContext.hpp:21-345)
//
// Error convention: the reference's seams return void and print on failure; the adapters print
// lmsf_strerror to stderr and leave the output as the reference would (empty clouds / pose untouched).
// In the reference tree: #define LMSF_WITH_REFERENCE before including this header and include the
// reference's process_base.hpp / filter_base.hpp / registration_base.hpp first.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "lmsf_b200.h"
#ifndef LMSF_WITH_REFERENCE
#include "lmsf_compat/reference_stubs.hpp"
#endif

namespace lmsf {

// shared, reference-counted context: one per LiDAR (the reference builds one processor + one tracker per LiDAR)
class Context {
 public:
  explicit Context(int device = 0, const lmsf_params* params = nullptr) {
    lmsf_params p;
    if (params) p = *params; else lmsf_params_default(&p);
    int rc = lmsf_ctx_create(device, &p, &ctx_);
    if (rc != LMSF_OK) throw std::runtime_error(std::string("lmsf_ctx_create: ") + lmsf_strerror(rc));
  }
  ~Context() { lmsf_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  lmsf_ctx* get() const { return ctx_; }
 private:
  lmsf_ctx* ctx_ = nullptr;
};
using ContextPtr = std::shared_ptr<Context>;

inline lmsf_params DefaultParams(int n_scans, float min_range, float max_range) {
  lmsf_params p;
  lmsf_params_default(&p);
  p.n_scans = n_scans;
  p.min_range = min_range;
  p.max_range = max_range;
  return p;
}

namespace detail {
template <typename PointT>
inline void pack(const pcl::PointCloud<PointT>& c, std::vector<float>& out) {
  out.resize(c.points.size() * 4);
  for (std::size_t i = 0; i < c.points.size(); ++i) {
    out[4 * i + 0] = c.points[i].x;
    out[4 * i + 1] = c.points[i].y;
    out[4 * i + 2] = c.points[i].z;
    out[4 * i + 3] = c.points[i].intensity;
  }
}
template <typename PointT>
inline typename pcl::PointCloud<PointT>::Ptr unpack(const float* xyzi, int n) {
  typename pcl::PointCloud<PointT>::Ptr c(new pcl::PointCloud<PointT>());
  c->points.resize(n);
  for (int i = 0; i < n; ++i) {
    PointT p{};
    p.x = xyzi[4 * i + 0];
    p.y = xyzi[4 * i + 1];
    p.z = xyzi[4 * i + 2];
    p.intensity = xyzi[4 * i + 3];
    c->points[i] = p;
  }
  c->width = (std::uint32_t)n;
  c->height = 1;
  c->is_dense = true;
  return c;
}
inline bool check(int rc, const char* what) {
  if (rc == LMSF_OK) return true;
  std::fprintf(stderr, "[lmsf_b200] %s failed: %s\n", what, lmsf_strerror(rc));
  return false;
}
// Eigen::Quaterniond(T.rotation()) and q.toRotationMatrix(), as the reference does around Solve()
template <typename Iso>
inline void iso_to_pose(const Iso& T, double p[7]) {
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = T.linear()(r, c);
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0);
    p[3] = 0.5 * s;
    s = 0.5 / s;
    p[0] = (R[7] - R[5]) * s;
    p[1] = (R[2] - R[6]) * s;
    p[2] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * s;
    s = 0.5 / s;
    p[3] = (R[k * 3 + j] - R[j * 3 + k]) * s;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
    p[0] = v[0];
    p[1] = v[1];
    p[2] = v[2];
  }
  for (int i = 0; i < 3; ++i) p[4 + i] = T.translation()(i);
}
template <typename Iso>
inline void pose_to_iso(const double p[7], Iso& T) {
  const double x = p[0], y = p[1], z = p[2], w = p[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  T.linear()(0, 0) = 1 - (tyy + tzz);
  T.linear()(0, 1) = txy - twz;
  T.linear()(0, 2) = txz + twy;
  T.linear()(1, 0) = txy + twz;
  T.linear()(1, 1) = 1 - (txx + tzz);
  T.linear()(1, 2) = tyz - twx;
  T.linear()(2, 0) = txz - twy;
  T.linear()(2, 1) = tyz + twx;
  T.linear()(2, 2) = 1 - (txx + tyy);
  for (int i = 0; i < 3; ++i) T.translation()(i) = p[4 + i];
}
}  // namespace detail

// ---------------------------------------------------------------- seam 1
template <typename _InputPointT, typename _OutputFeatureT>
class CudaLoamFeatureProcessor : public Algorithm::PointCloudProcessBase<_InputPointT, _OutputFeatureT> {
 public:
  // same leading arguments as LOAMFeatureProcessorBase(N_SCANS_, min_distance, max_distance, edge_thresh, ...)
  explicit CudaLoamFeatureProcessor(ContextPtr ctx) : ctx_(std::move(ctx)) {}
  void Process(Slam3D::LidarData<_InputPointT> const& data_in, Slam3D::CloudContainer<_OutputFeatureT>& data_out) override {
    detail::pack(data_in.point_cloud, in_);
    const int n = (int)data_in.point_cloud.points.size();
    edge_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    surf_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    int ne = 0, ns = 0;
    if (!detail::check(lmsf_extract_features(ctx_->get(), in_.data(), n, nullptr, edge_.data(), &ne, surf_.data(), &ns),
                       "lmsf_extract_features")) {
      ne = ns = 0;
    }
    data_out.pointcloud_data_.insert(std::make_pair(std::string("loam_edge"), detail::unpack<_OutputFeatureT>(edge_.data(), ne)));
    data_out.pointcloud_data_.insert(std::make_pair(std::string("loam_surf"), detail::unpack<_OutputFeatureT>(surf_.data(), ns)));
  }
 private:
  ContextPtr ctx_;
  std::vector<float> in_, edge_, surf_;
};

// ---------------------------------------------------------------- direct-method front end (row f4)
// Same surface as Algorithm::PointCloudCommonProcess<P> (processing/common_processing.hpp:28-121): removeNaN,
// VoxelGrid, DistanceFilter; the result is inserted under output_name.  SetOutlierRemoval is not offered.
template <typename _PointType>
class CudaPointCloudCommonProcess : public Algorithm::PointCloudProcessBase<_PointType, _PointType> {
 public:
  CudaPointCloudCommonProcess(ContextPtr ctx, std::string output_name, bool removal_nan = true)
      : ctx_(std::move(ctx)), output_name_(std::move(output_name)), removal_nan_(removal_nan) {}
  void SetVoxelGrid(std::string const& name, float leaf) {
    if (name == "VoxelGrid") leaf_ = leaf;
  }
  void SetDistanceFilter(float const& distance_near_thresh, float const& distance_far_thresh) {
    near_ = distance_near_thresh;
    far_ = distance_far_thresh;
  }
  void Process(Slam3D::LidarData<_PointType> const& data_in, Slam3D::CloudContainer<_PointType>& data_out) override {
    detail::pack(data_in.point_cloud, in_);
    const int n = (int)data_in.point_cloud.points.size();
    out_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    int m = 0;
    if (!detail::check(lmsf_common_process(ctx_->get(), in_.data(), n, removal_nan_ ? 1 : 0, leaf_, near_, far_, out_.data(), &m),
                       "lmsf_common_process"))
      m = 0;
    data_out.pointcloud_data_.insert(std::make_pair(output_name_, detail::unpack<_PointType>(out_.data(), m)));
  }
 private:
  ContextPtr ctx_;
  std::string output_name_;
  bool removal_nan_;
  float leaf_ = 0.f, near_ = 0.f, far_ = 0.f;
  std::vector<float> in_, out_;
};

// ---------------------------------------------------------------- sweep preprocessing (row f4)
// Same surface as Algorithm::RotaryLidarPreProcess<P> (Preprocess/RotaryLidar_preprocessing.hpp:22-104): Process(LidarData&)
// rewrites the cloud in place — the node's removeNaN (MultiLidarSLAM_node.cpp:126-133) folded in, intensity := relative time
// of the point in the sweep.  A tracker context created with lmsf_params.rotary_scan_period > 0 does the same inside its
// extraction and needs no separate call.
template <typename _PointT>
class CudaRotaryLidarPreProcess {
 public:
  explicit CudaRotaryLidarPreProcess(ContextPtr ctx, double SCAN_PERIOD = 0.1) : ctx_(std::move(ctx)), period_((float)SCAN_PERIOD) {}
  virtual ~CudaRotaryLidarPreProcess() {}
  virtual void Process(Slam3D::LidarData<_PointT>& lidar_data) {
    detail::pack(lidar_data.point_cloud, in_);
    const int n = (int)lidar_data.point_cloud.points.size();
    out_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    int m = 0;
    if (!detail::check(lmsf_rotary_preprocess(ctx_->get(), in_.data(), n, period_, out_.data(), &m), "lmsf_rotary_preprocess"))
      return;  // the cloud is left as it came
    lidar_data.point_cloud.points = detail::unpack<_PointT>(out_.data(), m)->points;
    lidar_data.point_cloud.width = (std::uint32_t)m;
    lidar_data.point_cloud.height = 1;
  }
 private:
  ContextPtr ctx_;
  float period_;
  std::vector<float> in_, out_;
};

// ---------------------------------------------------------------- seam 2
template <typename _PointType>
class CudaVoxelGridFilter : public Algorithm::FilterBase<_PointType> {
 public:
  CudaVoxelGridFilter(ContextPtr ctx, float leaf) : ctx_(std::move(ctx)), leaf_(leaf) {}
  // VoxelGridFilter::Reset("VoxelGrid", leaf) (Filter/voxel_grid.hpp:25-29)
  void Reset(std::string const& name, float leaf) {
    if (name == "VoxelGrid") leaf_ = leaf;
  }
  Algorithm::PointCloudPtr<_PointType> Filter(const Algorithm::PointCloudConstPtr<_PointType>& cloud_in) const override {
    std::vector<float> in, out;
    detail::pack(*cloud_in, in);
    const int n = (int)cloud_in->points.size();
    out.resize((std::size_t)(n > 0 ? n : 1) * 4);
    int nv = 0;
    if (!(leaf_ > 0.f) ||
        !detail::check(lmsf_voxel_downsample(ctx_->get(), in.data(), n, leaf_, out.data(), &nv, nullptr), "lmsf_voxel_downsample")) {
      return Algorithm::PointCloudPtr<_PointType>(new pcl::PointCloud<_PointType>(*cloud_in));  // filter_base.hpp:39-40
    }
    auto res = detail::unpack<_PointType>(out.data(), nv);
    res->header = cloud_in->header;  // filter_base.hpp:43
    return res;
  }
 private:
  ContextPtr ctx_;
  float leaf_;
};

// ---------------------------------------------------------------- seam 3
template <typename _PointType>
class CudaEdgeSurfRegistration : public Algorithm::RegistrationBase<_PointType> {
  using Base = Algorithm::RegistrationBase<_PointType>;
 public:
  CudaEdgeSurfRegistration(ContextPtr ctx, std::string const& edge_name, std::string const& surf_name,
                           int solver = LMSF_SOLVER_HUBER_LM)
      : ctx_(std::move(ctx)), edge_name_(edge_name), surf_name_(surf_name), solver_(solver) {}

  // "Source" is the local map, selected by name (ceres_edgeSurfFeatureRegistration.hpp:56-71)
  void SetInputSource(typename Base::SourceInput const& source_input) override {
    if (!source_input.second || source_input.second->empty()) return;
    int kind = -1;
    if (source_input.first == edge_name_) kind = LMSF_KIND_EDGE;
    else if (source_input.first == surf_name_) kind = LMSF_KIND_SURF;
    if (kind < 0) return;
    std::vector<float> buf;
    detail::pack(*source_input.second, buf);
    detail::check(lmsf_map_set(ctx_->get(), kind, buf.data(), (int)source_input.second->points.size()), "lmsf_map_set");
  }
  // "Target" is the current scan's feature container (:73-84); the clouds are kept by shared_ptr
  void SetInputTarget(Slam3D::FeaturePointCloudContainer<_PointType> const& target_input) override {
    auto e = target_input.find(edge_name_);
    if (e != target_input.end()) edge_in_ = e->second;
    auto s = target_input.find(surf_name_);
    if (s != target_input.end()) surf_in_ = s->second;
  }
  void SetMaxIteration(std::uint16_t const& n) { lmsf_set_lm_outer(ctx_->get(), (int)n); }
  // predicted pose in, registered pose out (:96-130); on failure T is left as predicted
  void Solve(Eigen::Isometry3d& T) override {
    std::vector<float> e, s;
    int ne = 0, ns = 0;
    if (edge_in_) { detail::pack(*edge_in_, e); ne = (int)edge_in_->points.size(); }
    if (surf_in_) { detail::pack(*surf_in_, s); ns = (int)surf_in_->points.size(); }
    double pose[7];
    detail::iso_to_pose(T, pose);
    lmsf_reg_stats st;
    if (detail::check(lmsf_register(ctx_->get(), e.data(), ne, s.data(), ns, solver_, pose, &st), "lmsf_register")) {
      detail::pose_to_iso(pose, T);
      last_ = st;
    }
  }
  const lmsf_reg_stats& LastStats() const { return last_; }
 private:
  ContextPtr ctx_;
  std::string edge_name_, surf_name_;
  int solver_;
  typename pcl::PointCloud<_PointType>::ConstPtr edge_in_, surf_in_;
  lmsf_reg_stats last_{};
};

// ---------------------------------------------------------------- alignment score (row f2)
// Same public surface as Slam3D::PointCloudAlignmentEvaluate<_PointT> (registration/alignEvaluate.hpp:23-98): the
// target cloud becomes a device index (map slot `kind` of the context: use a context of its own, or a slot the
// tracker does not need), the score is computed on the device.
template <typename _PointT>
class CudaPointCloudAlignmentEvaluate {
 public:
  explicit CudaPointCloudAlignmentEvaluate(ContextPtr ctx, std::string name = "processed", int kind = LMSF_KIND_SURF)
      : ctx_(std::move(ctx)), name_(std::move(name)), kind_(kind) {}
  std::string GetTargetName() const { return name_; }
  void SetTargetPoints(typename pcl::PointCloud<_PointT>::ConstPtr const& cloud) {
    std::vector<float> buf;
    detail::pack(*cloud, buf);
    set_target_ = detail::check(lmsf_map_set(ctx_->get(), kind_, buf.data(), (int)cloud->points.size()), "lmsf_map_set");
  }
  // (mean squared distance of the inliers, overlap ratio); (DBL_MAX, ratio) when the overlap is too small (:82-85)
  std::pair<double, double> AlignmentScore(typename pcl::PointCloud<_PointT>::ConstPtr const& cloud,
                                           Eigen::Matrix4f const& relpose, double const& inlier_thresh,
                                           double const& inlier_ratio_thresh) {
    const double none = 1.7976931348623157e308;
    if (!set_target_ || cloud->empty()) return std::make_pair(none, 0.0);
    std::vector<float> buf;
    detail::pack(*cloud, buf);
    float T[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) T[r * 4 + c] = relpose(r, c);
    double score = none, overlap = 0;
    if (!detail::check(lmsf_align_score(ctx_->get(), kind_, buf.data(), (int)cloud->points.size(), T, inlier_thresh,
                                        inlier_ratio_thresh, &score, &overlap, nullptr), "lmsf_align_score"))
      return std::make_pair(none, 0.0);
    return std::make_pair(score, overlap);
  }
 private:
  ContextPtr ctx_;
  std::string name_;
  int kind_;
  bool set_target_ = false;
};

// ---------------------------------------------------------------- multi-LiDAR extrinsic calibration loop (row f3)
// The calibration branch of MultiLidarSystem::process() (System/ML_System.hpp:239-323) for a primary and one
// auxiliary LiDAR, on two device contexts: status 0 = both trackers + hand-eye initialisation
// (HandEyeCalibrationBase, Algorithm/calibration/handeye_calibration_base.hpp), status 1 = primary tracker +
// registration of the auxiliary sweep against the primary's local map.  Poses are {qx,qy,qz,qw,tx,ty,tz}.
class CudaMultiLidarExtrinsics {
 public:
  CudaMultiLidarExtrinsics(ContextPtr primary, ContextPtr auxiliary) : ctx_{std::move(primary), std::move(auxiliary)} {
    if (lmsf_handeye_create(&he_) != LMSF_OK) throw std::runtime_error("lmsf_handeye_create");
  }
  ~CudaMultiLidarExtrinsics() { lmsf_handeye_destroy(he_); }
  CudaMultiLidarExtrinsics(const CudaMultiLidarExtrinsics&) = delete;
  CudaMultiLidarExtrinsics& operator=(const CudaMultiLidarExtrinsics&) = delete;

  int Status() const { return status_; }                 // EXTRINSIC_CALIB_STATUS_: 0 initialising, 1 refining
  const double* Extrinsic() const { return ext_; }       // lidar_lidar_estrinsic_: auxiliary LiDAR in the primary's frame
  const double* PrimaryPose() const { return pose0_; }

  // one synchronised pair of sweeps (float32 XYZI, n0 / n1 points)
  bool Process(const float* sweep0, int n0, const float* sweep1, int n1, double stamp) {
    double d0[7] = {0, 0, 0, 1, 0, 0, 0}, d1[7] = {0, 0, 0, 1, 0, 0, 0}, p1[7];
    lmsf_track_stats st0, st1;
    if (status_ == 0) {
      // (the reference runs the two trackers under `omp parallel for`; the contexts are independent, so may the caller)
      if (!detail::check(lmsf_tracker_step(ctx_[0]->get(), sweep0, n0, stamp, d0, pose0_, &st0), "lmsf_tracker_step") ||
          !detail::check(lmsf_tracker_step(ctx_[1]->get(), sweep1, n1, stamp, d1, p1, &st1), "lmsf_tracker_step"))
        return false;
      if (st0.first) return true;
      int enough = 0, ok = 0;
      lmsf_handeye_add_pose(he_, d0, d1, &enough);
      if (enough) {
        double e[7];
        lmsf_handeye_calibrate(he_, e, nullptr, &ok);
        if (ok) {
          std::memcpy(ext_, e, sizeof e);
          status_ = 1;
        }
      }
      return true;
    }
    if (!detail::check(lmsf_tracker_step(ctx_[0]->get(), sweep0, n0, stamp, d0, pose0_, &st0), "lmsf_tracker_step"))
      return false;
    double sub[7];
    mul(pose0_, ext_, sub);  // sub_lidar_pose = primary_lidar_pose * lidar_lidar_estrinsic_ (:301)
    lmsf_reg_stats rs;
    if (!detail::check(lmsf_tracker_register_aux(ctx_[0]->get(), sweep1, n1, sub, &rs), "lmsf_tracker_register_aux"))
      return false;
    double inv0[7];
    inv(pose0_, inv0);
    mul(inv0, sub, ext_);    // lidar_lidar_estrinsic_ = primary_lidar_pose.inverse() * sub_lidar_pose (:306)
    return true;
  }

 private:
  static void rotv(const double* q, const double* v, double* o) {
    const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
    double cx = 2 * (uy * v[2] - uz * v[1]), cy = 2 * (uz * v[0] - ux * v[2]), cz = 2 * (ux * v[1] - uy * v[0]);
    o[0] = v[0] + w * cx + (uy * cz - uz * cy);
    o[1] = v[1] + w * cy + (uz * cx - ux * cz);
    o[2] = v[2] + w * cz + (ux * cy - uy * cx);
  }
  static void mul(const double* a, const double* b, double* o) {
    double q[4] = {a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
                   a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]};
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), t[3];
    rotv(a, b + 4, t);
    for (int i = 0; i < 4; ++i) o[i] = q[i] / n;
    for (int i = 0; i < 3; ++i) o[4 + i] = t[i] + a[4 + i];
  }
  static void inv(const double* a, double* o) {
    double qi[4] = {-a[0], -a[1], -a[2], a[3]}, t[3];
    rotv(qi, a + 4, t);
    for (int i = 0; i < 4; ++i) o[i] = qi[i];
    for (int i = 0; i < 3; ++i) o[4 + i] = -t[i];
  }
  ContextPtr ctx_[2];
  lmsf_handeye* he_ = nullptr;
  int status_ = 0;
  double ext_[7] = {0, 0, 0, 1, 0, 0, 0};
  double pose0_[7] = {0, 0, 0, 1, 0, 0, 0};
};

// ---------------------------------------------------------------- loop-closure place recognition (row f1)
// Same public surface as Slam3D::SceneRecognitionScanContext<_PointCloudT> (AddKeyFramePoints :61-94,
// FindSimilarPointCloud :102-111, LoopDetect :119-134); descriptors, ring keys and the search live on the device.
template <typename _PointCloudT>
class CudaSceneRecognitionScanContext {
 public:
  using PointCloudContainer = std::unordered_map<std::string, typename pcl::PointCloud<_PointCloudT>::ConstPtr>;
  explicit CudaSceneRecognitionScanContext(ContextPtr ctx, std::vector<std::string> target_names = {})
      : ctx_(std::move(ctx)), target_names_(std::move(target_names)) {}

  void AddKeyFramePoints(PointCloudContainer const& pcl_in) {
    std::vector<float> buf;
    const int n = select(pcl_in, buf);
    int id = -1;
    if (!detail::check(lmsf_scdb_add_cloud(ctx_->get(), buf.data(), n, &id), "lmsf_scdb_add_cloud")) return;
    size_ = id + 1;
    limit_ = lmsf_sc_tree_limit(size_);  // the tree is rebuilt every 10th keyframe over [0, size - 50) (:74-92)
  }

  std::pair<std::int64_t, Eigen::Isometry3d> FindSimilarPointCloud(PointCloudContainer const& scan_in) {
    if (limit_ <= 0) return std::make_pair((std::int64_t)-1, Eigen::Isometry3d::Identity());  // no tree yet (:104)
    std::vector<float> buf, desc(LMSF_SC_CELLS), key(LMSF_SC_RINGS);
    const int n = select(scan_in, buf);
    if (!detail::check(lmsf_sc_make(ctx_->get(), buf.data(), n, desc.data(), key.data()), "lmsf_sc_make"))
      return std::make_pair((std::int64_t)-1, Eigen::Isometry3d::Identity());
    return find(key.data(), desc.data());
  }

  std::pair<std::int64_t, Eigen::Isometry3d> LoopDetect(std::uint32_t const& id) {
    if (size_ < 50 + 1 || limit_ <= 0) return std::make_pair((std::int64_t)-1, Eigen::Isometry3d::Identity());  // :129
    std::vector<float> desc(LMSF_SC_CELLS), key(LMSF_SC_RINGS);
    if (!detail::check(lmsf_scdb_get(ctx_->get(), (int)id, desc.data(), key.data()), "lmsf_scdb_get"))
      return std::make_pair((std::int64_t)-1, Eigen::Isometry3d::Identity());
    return find(key.data(), desc.data());
  }

  int Size() const { return size_; }
  double LastDistance() const { return last_dist_; }
  int LastShift() const { return last_shift_; }

 private:
  // extractInterestPointClouds (:236-252): every cloud of the container, or the named ones, concatenated
  int select(PointCloudContainer const& in, std::vector<float>& out) const {
    out.clear();
    auto append = [&out](const pcl::PointCloud<_PointCloudT>& c) {
      for (const auto& p : c.points) {
        out.push_back(p.x);
        out.push_back(p.y);
        out.push_back(p.z);
        out.push_back(p.intensity);
      }
    };
    if (target_names_.empty()) {
      for (auto it = in.begin(); it != in.end(); ++it)
        if (it->second) append(*it->second);
    } else {
      for (const std::string& name : target_names_) {
        auto it = in.find(name);
        if (it != in.end() && it->second) append(*it->second);
      }
    }
    return (int)(out.size() / 4);
  }
  // descFindSimilar (:260-333): ring-key top-10, best SC distance, threshold 0.2, yaw = shift * 6 deg about z
  std::pair<std::int64_t, Eigen::Isometry3d> find(const float* key, const float* desc) {
    std::int32_t id = -1, shift = 0;
    double dist = 0;
    Eigen::Isometry3d rel = Eigen::Isometry3d::Identity();
    if (!detail::check(lmsf_scdb_search(ctx_->get(), key, desc, 1, limit_, 0.2, &id, &dist, &shift), "lmsf_scdb_search"))
      return std::make_pair((std::int64_t)-1, rel);
    last_dist_ = dist;
    last_shift_ = shift;
    if (id < 0) return std::make_pair((std::int64_t)-1, rel);
    const float yaw = (float)((double)(shift * (360.0 / LMSF_SC_SECTORS)) * 3.14159265358979323846 / 180.0);  // deg2rad (:335)
    const double c = std::cos((double)yaw), s = std::sin((double)yaw);
    rel.linear()(0, 0) = c;
    rel.linear()(0, 1) = -s;
    rel.linear()(1, 0) = s;
    rel.linear()(1, 1) = c;
    return std::make_pair((std::int64_t)id, rel);
  }
  ContextPtr ctx_;
  std::vector<std::string> target_names_;
  int size_ = 0, limit_ = 0;
  double last_dist_ = 0;
  int last_shift_ = 0;
};

}  // namespace lmsf
