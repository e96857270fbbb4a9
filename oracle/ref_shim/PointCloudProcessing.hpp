// oracle/ref_shim/PointCloudProcessing.hpp -- TEST INFRASTRUCTURE.  Shadows the reference's PCL filter helpers so that
// tests/CameraPathGen.cpp compiles; downsample (a pcl::VoxelGrid filter there) is only reached from Planner methods that
// nothing here calls, and aborts if it ever is.
#pragma once
#include <cstdlib>
#include <pcl/point_cloud.h>
namespace PointCloudProcessing {
template <typename PointT> void downsample(typename pcl::PointCloud<PointT>::Ptr, typename pcl::PointCloud<PointT>::Ptr, double) { std::abort(); }
}
