// oracle/ref_shim/VisualizationUtilities.hpp -- TEST INFRASTRUCTURE.  Shadows the reference's PCL/VTK viewer wrapper so that
// its driver tests/CameraPathGen.cpp compiles (for willCollide, repositionCamerasSampled and setCover, which do not touch
// the viewer).  Every method is a no-op; the driver's Planner methods that draw are never called.
#pragma once
#include <string>
#include <vector>
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <Camera.hpp>
namespace VisualizationUtilities {
struct PCLVisualizerWrapper {
    PCLVisualizerWrapper() {}
    PCLVisualizerWrapper(int, int, int) {}
    template <typename PointT> void addPointCloud(typename pcl::PointCloud<PointT>::Ptr, std::string = "") {}
    void addCamera(Camera&, Eigen::Affine3f, std::string, int = 0) {}
    void addLine(std::vector<double>, std::vector<double>, std::string) {}
    void addCoordinateSystem() {}
    void spinViewer() {}
};
}
