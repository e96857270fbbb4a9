// Camera.hpp -- drop-in for the reference's include/Camera.hpp (class Camera, :17-86).
//
// Same public interface (constructor defaults, method names, argument and return types), so drivers that include
// <Camera.hpp> compile unchanged.  These host helpers are only used by drivers directly; the ray marches themselves run
// on the GPU (RayTracingEngine.hpp -> libdmf_b200.so) and read the intrinsics through intrinsics().
#pragma once
#include <cmath>
#include <tuple>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Core>

using namespace std;

class Camera
{
    vector<float> K_;            // row-major 3x3: fx 0 cx / 0 fy cy / 0 0 1
    int height_ = 480, width_ = 640;

    struct Pinhole { double fx, cx, fy, cy; };
    Pinhole pinhole() const { return Pinhole{K_[0], K_[2], K_[4], K_[5]}; }
    static double pixel_spacing(double ax, double ay, double bx, double by) { return sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by)); }

    public:
    Camera() {}
    Camera(vector<float>& K, int height = 480, int width = 640) : K_(K), height_(height), width_(width) {}

    // pixel (r,c) at depth_mm -> camera-frame point, evaluated in double and narrowed by the float tuple
    tuple<float, float, float> projectPoint(int r, int c, int depth_mm)
    {
        const Pinhole k = pinhole();
        const double z = depth_mm * 0.001;
        return make_tuple(z * ((double)c - k.cx) / k.fx, z * ((double)r - k.cy) / k.fy, z);
    }
    // camera-frame point -> (row, col), rounding half away from zero
    tuple<int, int> deProjectPoint(double x, double y, double z)
    {
        const Pinhole k = pinhole();
        const int col = int(round((x * k.fx) / z + k.cx));
        const int row = int(round((y * k.fy) / z + k.cy));
        return make_tuple(row, col);
    }
    tuple<float, float, float> transformPoints(double x, double y, double z, Eigen::Affine3f& transformation)
    {
        Eigen::Vector3f p((float)x, (float)y, (float)z);
        Eigen::Vector3f q = transformation * p;
        return make_tuple(q(0), q(1), q(2));
    }
    tuple<float, float, float> getPoint(int r, int c, int depth_mm) { return projectPoint(r, c, depth_mm); }
    tuple<int, int> getPixel(double x, double y, double z, Eigen::Affine3f transformation = Eigen::Affine3f::Identity())
    {
        std::tie(x, y, z) = transformPoints(x, y, z, transformation);
        return deProjectPoint(x, y, z);
    }
    int getHeight() { return height_; }
    int getWidth() { return width_; }
    bool validPixel(int r, int c) { return r >= 0 && c >= 0 && r < height_ && c < width_; }
    float getAreaCovered(int depth_mm)
    {
        double x1, y1, x2, y2, x3, y3, z;
        std::tie(x1, y1, z) = getPoint(0, 0, depth_mm);
        std::tie(x2, y2, z) = getPoint(0, height_, depth_mm);
        std::tie(x3, y3, z) = getPoint(width_, 0, depth_mm);
        return pixel_spacing(x1, y1, x2, y2) * pixel_spacing(x1, y1, x3, y3);
    }
    float getDistance(int depth_mm)
    {
        double x1, y1, x2, y2, z;
        std::tie(x1, y1, z) = getPoint(100, 100, depth_mm);
        std::tie(x2, y2, z) = getPoint(101, 101, depth_mm);
        return pixel_spacing(x1, y1, x2, y2);
    }

    // ---- addition (not in the reference): read access for the GPU engine ----
    const vector<float>& intrinsics() const { return K_; }
};
