// oracle/shim_inferred: <pcl/registration/registration.h> — the interface registration_adapter.hpp names in its
// pcl::Registration specialization (:86-178).  Never instantiated by the tests (the feature path uses RegistrationBase);
// present so that the reference's header parses.  Test infrastructure only.
#pragma once
#include <pcl/point_cloud.h>

namespace pcl {
template <typename S, typename T>
class Registration {
 public:
  virtual ~Registration() {}
  virtual void setInputTarget(typename PointCloud<T>::ConstPtr const&) {}
  virtual void setInputSource(typename PointCloud<S>::ConstPtr const&) {}
  virtual void align(PointCloud<S>&, Eigen::Matrix<float, 4, 4> const&) {}
  virtual bool hasConverged() { return false; }
  virtual Eigen::Matrix4f getFinalTransformation() { return Eigen::Matrix4f(); }
  virtual float getFitnessScore() { return 0.f; }
};
}  // namespace pcl
