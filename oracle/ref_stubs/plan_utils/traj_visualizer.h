// TEST INFRASTRUCTURE ONLY: stands in for plan_utils/traj_visualizer.h (RViz drawing, DecompUtil corridor messages).  The
// optimizer source only needs the few corridor-debug types its displayBugPoly() names; nothing here does anything.
#pragma once
#include <Eigen/Eigen>
#include <ros/ros.h>
#include <visualization_msgs/Marker.h>
#include <visualization_msgs/MarkerArray.h>
#include <vector>
#include "plan_utils/traj_container.hpp"
template <class T> using vec_E = std::vector<T>;
struct Hyperplane2D {
  Hyperplane2D() {}
  template <class A, class B> Hyperplane2D(const A&, const B&) {}
};
struct Polyhedron2D {
  void add(const Hyperplane2D&) {}
};
namespace decomp_ros_msgs {
struct PolyhedronArray { std_msgs::Header header; };
}  // namespace decomp_ros_msgs
namespace DecompROS {
inline decomp_ros_msgs::PolyhedronArray polyhedron_array_to_ros(const vec_E<Polyhedron2D>&) { return decomp_ros_msgs::PolyhedronArray(); }
}  // namespace DecompROS
