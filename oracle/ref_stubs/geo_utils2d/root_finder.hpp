// TEST INFRASTRUCTURE ONLY: stands in for geo_utils2d/root_finder.hpp (included, not used, by plan_utils/poly_traj_utils.hpp)
#pragma once
