// compat/pcl/point_types.h -- minimal stand-ins for the PCL point types Volume.hpp stores (no PCL in this container).
#pragma once
#include <cstdint>
namespace pcl {
struct PointXYZRGB { float x = 0, y = 0, z = 0; std::uint8_t r = 0, g = 0, b = 0; };
struct Normal {
    union { float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
    float curvature = 0;
    Normal() : normal{0, 0, 0} {}
    Normal(float nx, float ny, float nz) : normal{nx, ny, nz} {}
};
struct PointXYZRGBNormal { float x = 0, y = 0, z = 0; std::uint8_t r = 0, g = 0, b = 0; float normal[3] = {0, 0, 0}; };
}  // namespace pcl
