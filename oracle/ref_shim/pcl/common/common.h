// oracle/ref_shim/pcl/common/common.h -- TEST INFRASTRUCTURE: empty stand-in so that the reference's Algorithms.hpp (which includes it
// but uses nothing from it on the functions we call) compiles in this PCL-less container.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
// compile-only (Planner's constructor in tests/CameraPathGen.cpp, never called)
template <typename PointT> void getMinMax3D(const PointCloud<PointT>&, PointT&, PointT&) {}
}
