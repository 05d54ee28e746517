// reference_prelude.hpp — what a translation unit of the reference has in scope before it reaches the three seams: used
// to compile lmsf_b200_adapters.hpp with LMSF_WITH_REFERENCE against the reference's OWN headers (Sensor/lidar_data_type.h,
// processing/process_base.hpp:25-39, processing/Filter/filter_base.hpp:24-54, registration/registration_base.hpp:24-34),
// compiled where they lie under /root/reference; PCL / Eigen are the container-only stand-ins of oracle/shim +
// oracle/shim_fixed.  Test infrastructure (tests/cpp/Makefile: adapter_smoke_ref).
#include <cmath>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
using namespace std;
#define LMSF_SHIM_EIGEN_MATRIX4F
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "Sensor/lidar_data_type.h"
#include "Algorithm/PointClouds/processing/process_base.hpp"
#include "Algorithm/PointClouds/processing/Filter/filter_base.hpp"
#include "Algorithm/PointClouds/registration/registration_base.hpp"
