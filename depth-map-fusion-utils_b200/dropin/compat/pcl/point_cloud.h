// compat/pcl/point_cloud.h -- minimal pcl::PointCloud (points vector + shared Ptr).
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    std::vector<PointT> points;
    std::size_t size() const { return points.size(); }
};
}  // namespace pcl
