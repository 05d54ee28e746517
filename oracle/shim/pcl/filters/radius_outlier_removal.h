// oracle/shim: pcl::RadiusOutlierRemoval — declaration only; PCL's arithmetic is not available here (see README.md).
#pragma once
#include <cstdio>
#include <cstdlib>

#include <pcl/filters/filter.h>

namespace pcl {

template <typename PointT>
class RadiusOutlierRemoval : public Filter<PointT> {
 public:
  void setLeafSize(float, float, float) {}
  void setRadiusSearch(double) {}
  void setMinNeighborsInRadius(int) {}
  void setMeanK(int) {}
  void setStddevMulThresh(double) {}

 protected:
  void applyFilter(PointCloud<PointT>&) override {
    std::fprintf(stderr, "oracle/shim: pcl::RadiusOutlierRemoval is not available in this image\n");
    std::abort();
  }
};

}  // namespace pcl
