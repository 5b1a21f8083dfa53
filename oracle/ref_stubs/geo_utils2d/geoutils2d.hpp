// TEST INFRASTRUCTURE ONLY: stands in for geo_utils2d/geoutils2d.hpp (polytope vertex enumeration with SDLP / quickhull),
// which the optimizer header references from an inline helper (extractVs) that the optimisation path never calls.
#pragma once
#include <Eigen/Eigen>
namespace geoutils {
inline bool enumerateVs(const Eigen::MatrixXd&, Eigen::MatrixXd&) { return false; }
}  // namespace geoutils
