// oracle/ref_shim/pcl/point_types.h -- TEST INFRASTRUCTURE: the PCL point types the reference's Volume.hpp stores and Algorithms.hpp places cameras from.
#pragma once
#include <cstdint>
namespace pcl {
struct PointXYZRGB { float x = 0, y = 0, z = 0; std::uint8_t r = 0, g = 0, b = 0; };
struct Normal { float normal[3] = {0, 0, 0}; float curvature = 0; };
struct PointXYZ { float x = 0, y = 0, z = 0; };                       // tests/CameraPathGen.cpp (kd-tree queries, never executed)
struct PointXYZRGBNormal { float x = 0, y = 0, z = 0; std::uint8_t r = 0, g = 0, b = 0; float normal[3] = {0, 0, 0}; float curvature = 0; };   // Algorithms.hpp
}
