// adapter_smoke.cpp — drives the three reference seams through the C++ adapters exactly as the reference's
// factories would (ML_SystemFactory.hpp:179-198): processor -> registration with "loam_edge"/"loam_surf".
// Input: a binary file of float32 XYZI points (two sweeps back to back: n0, n1 given on the command line).
// Prints the registered pose; tests/test_gpu_parity.py compares it with the ctypes path.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lmsf_b200_adapters.hpp"

using Pt = pcl::PointXYZI;

static Slam3D::LidarData<Pt> load(const float* p, int n) {
  Slam3D::LidarData<Pt> d;
  d.point_cloud.points.resize(n);
  for (int i = 0; i < n; ++i) {
    Pt q{};
    q.x = p[4 * i];
    q.y = p[4 * i + 1];
    q.z = p[4 * i + 2];
    q.intensity = p[4 * i + 3];
    d.point_cloud.points[i] = q;
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: adapter_smoke <file> <n0> <n1> <n_scans>\n");
    return 2;
  }
  const int n0 = std::atoi(argv[2]), n1 = std::atoi(argv[3]), n_scans = std::atoi(argv[4]);
  std::vector<float> buf((std::size_t)(n0 + n1) * 4);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f || std::fread(buf.data(), sizeof(float), buf.size(), f) != buf.size()) return 3;
  std::fclose(f);
  lmsf_params prm = lmsf::DefaultParams(n_scans, 2.f, 80.f);
  lmsf::ContextPtr ctx;
  try {
    ctx = std::make_shared<lmsf::Context>(0, &prm);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 4;
  }
  std::unique_ptr<Algorithm::PointCloudProcessBase<Pt, Pt>> proc(new lmsf::CudaLoamFeatureProcessor<Pt, Pt>(ctx));
  std::unique_ptr<Algorithm::FilterBase<Pt>> vox(new lmsf::CudaVoxelGridFilter<Pt>(ctx, 0.4f));
  std::unique_ptr<Algorithm::RegistrationBase<Pt>> reg(new lmsf::CudaEdgeSurfRegistration<Pt>(ctx, "loam_edge", "loam_surf"));

  Slam3D::CloudContainer<Pt> f0, f1;
  proc->Process(load(buf.data(), n0), f0);
  proc->Process(load(buf.data() + 4 * (std::size_t)n0, n1), f1);
  auto surf_map = vox->Filter(f0.pointcloud_data_["loam_surf"]);
  reg->SetInputSource({"loam_edge", f0.pointcloud_data_["loam_edge"]});
  reg->SetInputSource({"loam_surf", surf_map});
  reg->SetInputTarget(f1.pointcloud_data_);
  Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
  reg->Solve(T);
  double p[7];
  lmsf::detail::iso_to_pose(T, p);
  std::printf("features %zu %zu %zu %zu voxels %zu\n", f0.pointcloud_data_["loam_edge"]->size(),
              f0.pointcloud_data_["loam_surf"]->size(), f1.pointcloud_data_["loam_edge"]->size(),
              f1.pointcloud_data_["loam_surf"]->size(), surf_map->size());
  std::printf("pose %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", p[0], p[1], p[2], p[3], p[4], p[5], p[6]);

  // alignment score as loopDetection.hpp:176-178 drives it (its own context: SetTargetPoints replaces a map index)
  {
    lmsf::ContextPtr ctx2 = std::make_shared<lmsf::Context>(0, &prm);
    lmsf::CudaPointCloudAlignmentEvaluate<Pt> eval(ctx2, "loam_surf");
    eval.SetTargetPoints(f0.pointcloud_data_["loam_surf"]);
    Eigen::Matrix4f Tf = Eigen::Matrix4f::Identity();
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Tf(r, c) = (float)T.linear()(r, c);
      Tf(r, 3) = (float)T.translation()(r);
    }
    auto sc = eval.AlignmentScore(f1.pointcloud_data_["loam_surf"], Tf, 1.0, 0.3);
    auto self = eval.AlignmentScore(f0.pointcloud_data_["loam_surf"], Eigen::Matrix4f::Identity(), 0.1, 0.6);
    std::printf("align %.17g %.17g %.17g %.17g\n", sc.first, sc.second, self.first, self.second);
  }

  // direct-method front end: removeNaN + VoxelGrid + DistanceFilter behind PointCloudProcessBase
  {
    std::unique_ptr<Algorithm::PointCloudProcessBase<Pt, Pt>> pre;
    auto* cp = new lmsf::CudaPointCloudCommonProcess<Pt>(ctx, "processed", true);
    cp->SetVoxelGrid("VoxelGrid", 0.5f);
    cp->SetDistanceFilter(3.0f, 40.0f);
    pre.reset(cp);
    Slam3D::CloudContainer<Pt> pc;
    pre->Process(load(buf.data(), n0), pc);
    std::printf("common %zu\n", pc.pointcloud_data_["processed"]->size());
  }

  // sweep preprocessing behind RotaryLidarPreProcess's surface: the relative times end up in the intensity channel
  {
    lmsf::CudaRotaryLidarPreProcess<Pt> rot(ctx, 0.1);
    Slam3D::LidarData<Pt> d = load(buf.data(), n0);
    rot.Process(d);
    float lo = 1e9f, hi = -1e9f;
    for (auto const& q : d.point_cloud.points) {
      lo = q.intensity < lo ? q.intensity : lo;
      hi = q.intensity > hi ? q.intensity : hi;
    }
    std::printf("rotary %zu %.6f %.6f\n", d.point_cloud.points.size(), lo, hi);
  }

  // multi-LiDAR extrinsics: compile-and-run check of the calibration loop class (two contexts, same sweeps: the
  // increments are identical, so the pairs are accepted; identical LiDARs cannot excite the rotation -> status stays 0)
  {
    lmsf::ContextPtr ca = std::make_shared<lmsf::Context>(0, &prm), cb = std::make_shared<lmsf::Context>(0, &prm);
    lmsf::CudaMultiLidarExtrinsics rig(ca, cb);
    bool ok0 = rig.Process(buf.data(), n0, buf.data(), n0, 0.0);
    bool ok1 = rig.Process(buf.data() + 4 * (std::size_t)n0, n1, buf.data() + 4 * (std::size_t)n0, n1, 0.1);
    std::printf("rig %d %d %d\n", (int)ok0, (int)ok1, rig.Status());
  }

  // place recognition as loopDetection.hpp drives SceneRecognitionScanContext: 60 keyframes (sweep 0, then sweep 1
  // repeated), then the revisit query (sweep 0 again) and LoopDetect on a stored keyframe
  lmsf::CudaSceneRecognitionScanContext<Pt> scene(ctx, {"loam_edge", "loam_surf"});
  auto before = scene.FindSimilarPointCloud(f0.pointcloud_data_);
  scene.AddKeyFramePoints(f0.pointcloud_data_);
  for (int k = 1; k < 61; ++k) scene.AddKeyFramePoints(f1.pointcloud_data_);
  auto hit = scene.FindSimilarPointCloud(f0.pointcloud_data_);
  std::printf("loop %lld %lld %d %.17g %d\n", (long long)before.first, (long long)hit.first, scene.Size(),
              scene.LastDistance(), scene.LastShift());
  auto self = scene.LoopDetect(5);
  std::printf("loopdetect %lld %.17g %d\n", (long long)self.first, scene.LastDistance(), scene.LastShift());
  return 0;
}
