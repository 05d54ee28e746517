// oracle/shim: pcl::VoxelGrid — declaration only; PCL's arithmetic is not available here (see README.md).
#pragma once
#include <cstdio>
#include <cstdlib>

#include <pcl/filters/filter.h>

namespace pcl {

template <typename PointT>
class VoxelGrid : public Filter<PointT> {
 public:
  void setLeafSize(float, float, float) {}
  void setRadiusSearch(double) {}
  void setMinNeighborsInRadius(int) {}
  void setMeanK(int) {}
  void setStddevMulThresh(double) {}

 protected:
  void applyFilter(PointCloud<PointT>&) override {
    std::fprintf(stderr, "oracle/shim: pcl::VoxelGrid is not available in this image\n");
    std::abort();
  }
};

}  // namespace pcl
