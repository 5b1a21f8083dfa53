// TEST INFRASTRUCTURE ONLY: stands in for common/basics/semantics.h (nothing of it is used by plan_utils/poly_traj_utils.hpp)
#pragma once
