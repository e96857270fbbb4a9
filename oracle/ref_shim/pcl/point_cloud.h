#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT> struct PointCloud { typedef std::shared_ptr<PointCloud<PointT>> Ptr; std::vector<PointT> points; };
}
