// TEST INFRASTRUCTURE ONLY: stands in for <ros/ros.h> when reference sources are compiled into oracle/_ref.  The
// optimizer only logs, reads the clock and publishes debug markers through ROS; all of that is inert here.
#pragma once
#include <chrono>
#include <cstdio>
#include <string>
#include <vector>
namespace ros {
struct Duration {
  double s = 0.0;
  double toSec() const { return s; }
};
struct Time {
  double s = 0.0;
  static Time now() {
    Time t;
    t.s = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return t;
  }
  double toSec() const { return s; }
  Duration operator-(const Time& o) const { return Duration{s - o.s}; }
};
struct Publisher {
  template <class M> void publish(const M&) const {}
};
struct NodeHandle {
  template <class M> Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
};
inline void shutdown() {}
}  // namespace ros
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
