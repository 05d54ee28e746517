// oracle/shim: pcl::KdTreeFLANN's two calls used on the path (setInputCloud, nearestKSearch), answered by the
// REFERENCE'S OWN vendored nanoflann 1.3.2 (exact search, L2_Simple: the same sequential fp32 accumulation as FLANN's
// L2_Simple behind pcl::KdTreeFLANN, leaf 15 as PCL sets it).  FLANN itself is not in this image.  See README.md.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

#include <pcl/point_cloud.h>

#include "nanoflann.hpp"

namespace pcl {

template <typename PointT>
class KdTreeFLANN {
 public:
  using Ptr = std::shared_ptr<KdTreeFLANN<PointT>>;
  using ConstPtr = std::shared_ptr<const KdTreeFLANN<PointT>>;

  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& cloud) {
    cloud_ = cloud;
    view_.pts = &cloud_->points;
    tree_.reset(new Tree(3, view_, nanoflann::KDTreeSingleIndexAdaptorParams(15)));
    tree_->buildIndex();
  }

  // k nearest, ascending; returns the number found (k = min(k, cloud size), as PCL)
  int nearestKSearch(const PointT& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    const std::size_t n = cloud_ ? cloud_->points.size() : 0;
    if ((std::size_t)k > n) k = (int)n;
    k_indices.resize(k);
    k_sqr_distances.resize(k);
    if (k == 0) return 0;
    std::vector<std::size_t> idx(k);
    nanoflann::KNNResultSet<float> rs(k);
    rs.init(idx.data(), k_sqr_distances.data());
    const float q[3] = {p.x, p.y, p.z};
    tree_->findNeighbors(rs, q, nanoflann::SearchParams(10));
    for (int i = 0; i < k; ++i) k_indices[i] = (int)idx[i];
    return (int)rs.size();
  }

 private:
  struct View {
    const std::vector<PointT>* pts = nullptr;
    std::size_t kdtree_get_point_count() const { return pts->size(); }
    float kdtree_get_pt(const std::size_t i, const std::size_t dim) const {
      const PointT& p = (*pts)[i];
      return dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
    }
    template <class BBOX>
    bool kdtree_get_bbox(BBOX&) const { return false; }
  };
  using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, View>, View, 3>;
  typename PointCloud<PointT>::ConstPtr cloud_;
  View view_;
  std::unique_ptr<Tree> tree_;
};

}  // namespace pcl
