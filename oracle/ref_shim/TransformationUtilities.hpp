// oracle/ref_shim/TransformationUtilities.hpp -- TEST INFRASTRUCTURE.  Shadows the reference's TransformationUtilities.hpp
// (dynamic Eigen matrices, AngleAxis ... far outside the minimal Eigen look-alike) so that the reference's Algorithms.hpp
// compiles unmodified.  Algorithms.hpp uses exactly one name from it, in positionCamerasOnHemisphere (Algorithms.hpp:340),
// which nothing here calls; the stand-in aborts if it ever is.
#pragma once
#include <cstdlib>
#include <vector>
#include <Eigen/Dense>
namespace TransformationUtilities {
inline Eigen::Affine3f vectorToAffineMatrix(std::vector<double>) { std::abort(); }
}
