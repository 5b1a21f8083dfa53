// TEST INFRASTRUCTURE ONLY: the two types of the reference's `common` package that plan_utils/poly_traj_utils.hpp names
// (fields / accessors it touches only), so that the header compiles into oracle/_ref without the ROS workspace.
#pragma once
#include <Eigen/Eigen>
namespace common {
struct State {
  Eigen::Vector2d vec_position;
  double angle = 0.0, curvature = 0.0, velocity = 0.0, acceleration = 0.0, steer = 0.0, time_stamp = 0.0;
};
class VehicleParam {
 public:
  double width() const { return width_; }
  double length() const { return length_; }
  double wheel_base() const { return wheel_base_; }
  double d_cr() const { return d_cr_; }
  void set_width(double v) { width_ = v; }
  void set_length(double v) { length_ = v; }
  void set_wheel_base(double v) { wheel_base_ = v; }
  void set_d_cr(double v) { d_cr_ = v; }
  double width_ = 1.9, length_ = 4.88, wheel_base_ = 2.85, d_cr_ = 1.015;
};
}  // namespace common
