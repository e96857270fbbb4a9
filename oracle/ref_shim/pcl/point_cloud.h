#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT> struct PointCloud { typedef std::shared_ptr<PointCloud<PointT>> Ptr; std::vector<PointT> points; };
}
// compile-only stand-in for tests/CameraPathGen.cpp's Planner::run_* methods (never executed)
namespace pcl {
template <typename PointT> struct KdTreeFLANN {
    void setInputCloud(typename PointCloud<PointT>::Ptr) {}
    int radiusSearch(const PointT&, double, std::vector<int>&, std::vector<float>&) { return 0; }
    int nearestKSearch(const PointT&, int, std::vector<int>&, std::vector<float>&) { return 0; }
};
}
