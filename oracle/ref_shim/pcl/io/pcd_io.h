// oracle/ref_shim/pcl/io/pcd_io.h -- TEST INFRASTRUCTURE: empty stand-in so that the reference's Algorithms.hpp (which includes it
// but uses nothing from it on the functions we call) compiles in this PCL-less container.
#pragma once
