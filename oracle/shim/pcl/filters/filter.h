// oracle/shim: pcl::Filter interface + removeNaNFromPointCloud (see README.md).
#pragma once
#include <cmath>
#include <memory>
#include <vector>

#include <pcl/point_cloud.h>

namespace pcl {

template <typename PointT>
class Filter {
 public:
  using Ptr = std::shared_ptr<Filter<PointT>>;
  using ConstPtr = std::shared_ptr<const Filter<PointT>>;
  virtual ~Filter() {}
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& cloud) { input_ = cloud; }
  void filter(PointCloud<PointT>& out) { applyFilter(out); }

 protected:
  virtual void applyFilter(PointCloud<PointT>& out) = 0;
  typename PointCloud<PointT>::ConstPtr input_;
};

template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
  if (&in != &out) {
    out.header = in.header;
    out.points.resize(in.points.size());
  }
  index.resize(in.points.size());
  std::size_t j = 0;
  for (std::size_t i = 0; i < in.points.size(); ++i) {
    const PointT p = in.points[i];
    if (!in.is_dense && !(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
    out.points[j] = p;
    index[j] = static_cast<int>(i);
    ++j;
  }
  out.points.resize(j);
  index.resize(j);
  out.height = 1;
  out.width = static_cast<std::uint32_t>(j);
  out.is_dense = true;
}

}  // namespace pcl
