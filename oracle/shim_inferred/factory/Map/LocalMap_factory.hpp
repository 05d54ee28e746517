// oracle/shim_inferred: factory/Map/LocalMap_factory.hpp — TEST INFRASTRUCTURE, and an INFERENCE, not the reference's code.
//
// LidarTracker/LidarTrackerLocalMap.hpp:15 includes this header and uses PointCloudLocalMapBase<_PointType> and
// make_localMap<_PointType>(type, name, params...) from it (:39-40, :81-82), but the file is absent from the reference
// tree (SURVEY.md §8 row a6').  What the tracker itself requires of it — and nothing more — is restated here so that the
// reference's own tracker can be compiled and run (oracle/ref_tracker.cpp):
//   AddFrameForMotion(cloud)  :221  a new keyframe enters the window; the oldest leaves once `window` frames are held
//   AddFrameForTime(cloud)    :226  the newest frame is refreshed without growing the window
//   GetLocalMap()             :229  (name, concatenation of the frames, oldest first) — handed to SetInputSource
//   is_full()                 :147
// The tracker reuses ONE transformed cloud object for every feature (:208, :217), so a frame is copied when it enters.
// The window length is the factory's default when the caller passes none (multiLidarEstimator_factory.hpp:194).
#pragma once
#include <deque>
#include <memory>
#include <string>
#include <utility>

#include <pcl/point_cloud.h>

namespace Slam3D {

template <typename _PointType>
class PointCloudLocalMapBase {
 public:
  using PointCloudPtr = typename pcl::PointCloud<_PointType>::Ptr;
  using PointCloudConstPtr = typename pcl::PointCloud<_PointType>::ConstPtr;
  virtual ~PointCloudLocalMapBase() {}
  virtual void AddFrameForMotion(PointCloudConstPtr const& frame) = 0;
  virtual void AddFrameForTime(PointCloudConstPtr const& frame) = 0;
  virtual std::pair<std::string, PointCloudConstPtr> GetLocalMap() const = 0;
  virtual bool is_full() const = 0;
};

template <typename _PointType>
class SlidingWindowLocalMap : public PointCloudLocalMapBase<_PointType> {
  using Base = PointCloudLocalMapBase<_PointType>;

 public:
  SlidingWindowLocalMap(std::string const& name, int window) : name_(name), window_(window) {}
  void AddFrameForMotion(typename Base::PointCloudConstPtr const& frame) override {
    frames_.push_back(*frame);
    if ((int)frames_.size() > window_) frames_.pop_front();
    rebuild();
  }
  void AddFrameForTime(typename Base::PointCloudConstPtr const& frame) override {
    if (!frames_.empty()) frames_.pop_back();
    frames_.push_back(*frame);
    rebuild();
  }
  std::pair<std::string, typename Base::PointCloudConstPtr> GetLocalMap() const override {
    return std::make_pair(name_, typename Base::PointCloudConstPtr(map_));
  }
  bool is_full() const override { return (int)frames_.size() >= window_; }

 private:
  void rebuild() {
    auto m = std::make_shared<pcl::PointCloud<_PointType>>();
    for (auto const& f : frames_) m->points.insert(m->points.end(), f.points.begin(), f.points.end());
    map_ = m;
  }
  std::string name_;
  int window_;
  std::deque<pcl::PointCloud<_PointType>> frames_;
  typename Base::PointCloudPtr map_ = std::make_shared<pcl::PointCloud<_PointType>>();
};

extern int g_inferred_window;  // defined by the translation unit that instantiates the tracker (default 10)

template <typename _PointType>
std::unique_ptr<PointCloudLocalMapBase<_PointType>> make_localMap(std::string const& /*type*/, std::string const& name) {
  return std::unique_ptr<PointCloudLocalMapBase<_PointType>>(new SlidingWindowLocalMap<_PointType>(name, g_inferred_window));
}
template <typename _PointType>
std::unique_ptr<PointCloudLocalMapBase<_PointType>> make_localMap(std::string const& /*type*/, std::string const& name,
                                                                  int window) {
  return std::unique_ptr<PointCloudLocalMapBase<_PointType>>(new SlidingWindowLocalMap<_PointType>(name, window));
}

}  // namespace Slam3D
