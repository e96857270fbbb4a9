// oracle/ref_shim/VisualizationUtilities.hpp -- TEST INFRASTRUCTURE.  Shadows the reference's PCL/VTK viewer wrapper
// (include/VisualizationUtilities.hpp: PCLVisualizerWrapper :70-110, VizThread :432-471) so that its drivers compile headless:
// tests/CameraPathGen.cpp (built into oracle/_ref for willCollide, repositionCamerasSampled and setCover, which do not touch the
// viewer) and, for the relink proof of tests/test_relink_cpu.py, tests/SetCover.cpp, Raytracing.cpp, CameraMotionPlanner.cpp,
// CameraPlacement.cpp against the drop-in headers.  Same member names and call shapes; every method is a no-op.
#pragma once
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Camera.hpp>
namespace pcl { namespace visualization {
enum RenderingProperties { PCL_VISUALIZER_POINT_SIZE, PCL_VISUALIZER_OPACITY, PCL_VISUALIZER_LINE_WIDTH, PCL_VISUALIZER_COLOR };
// whatever the drivers call on viz.viewer_ (removeAllShapes, removeCoordinateSystem, setPointCloudRenderingProperties, addSphere, ...)
struct PCLVisualizer {
    typedef PCLVisualizer* Ptr;
    template <class... A> bool removeAllShapes(A&&...) { return true; }
    template <class... A> bool removeAllPointClouds(A&&...) { return true; }
    template <class... A> bool removeCoordinateSystem(A&&...) { return true; }
    template <class... A> bool removeAllCoordinateSystems(A&&...) { return true; }
    template <class... A> bool removeShape(A&&...) { return true; }
    template <class... A> bool removePointCloud(A&&...) { return true; }
    template <class... A> bool setPointCloudRenderingProperties(A&&...) { return true; }
    template <class... A> bool setShapeRenderingProperties(A&&...) { return true; }
    template <class... A> bool addSphere(A&&...) { return true; }
    template <class... A> bool addLine(A&&...) { return true; }
    template <class... A> bool addText3D(A&&...) { return true; }
    template <class... A> bool addCoordinateSystem(A&&...) { return true; }
    template <class... A> void setBackgroundColor(A&&...) {}
    template <class... A> void spinOnce(A&&...) {}
    bool wasStopped() const { return true; }
};
} }
namespace VisualizationUtilities {
struct PCLVisualizerWrapper {
    pcl::visualization::PCLVisualizer viewer_storage_;
    pcl::visualization::PCLVisualizer::Ptr viewer_ = &viewer_storage_;
    PCLVisualizerWrapper() {}
    PCLVisualizerWrapper(double, double, double) {}
    template <typename PointT> void addPointCloud(typename pcl::PointCloud<PointT>::Ptr, std::string = "") {}
    template <typename PointT> void updatePointCloud(typename pcl::PointCloud<PointT>::Ptr, std::string = "") {}
    template <typename PointT> void addPointCloudNormals(typename pcl::PointCloud<PointT>::Ptr, pcl::PointCloud<pcl::Normal>::Ptr) {}
    template <class Volume> void addVolume(Volume&, int = 0) {}
    template <class Volume> void addVolumeWithVoxelsClassified(Volume&) {}
    template <class Volume> void addPointCloudInVolume(Volume&) {}
    template <class Volume> void addPointCloudInVolumeRayTraced(Volume&) {}
    void addCamera(Camera&, Eigen::Affine3f, std::string, int = 0) {}
    void addLine(std::vector<double>, std::vector<double>, std::string) {}
    template <class... A> void addSphere(A&&...) {}
    void addCoordinateSystem() {}
    void spinViewer() {}
    void spinViewerOnce() {}
    void clear() {}
    void execute() {}
    bool viewerGood() { return false; }
};
}
class VizThread {
    std::mutex mtx_;
    bool changed_ = false;
    virtual void input() {}
    virtual void process(VisualizationUtilities::PCLVisualizerWrapper&) {}
  public:
    virtual ~VizThread() {}
    bool updateViewer() { mtx_.lock(); changed_ = true; mtx_.unlock(); return false; }
    void makeThreads() { input(); }
    void spin() {}
};
