// STAND-IN (test infrastructure): ros::Time::now().toSec() is all the hot-path headers use (timing printouts).
#ifndef VXREF_ROS_H
#define VXREF_ROS_H
#include <chrono>
namespace ros {
struct Time {
  double t = 0;
  Time() {}
  explicit Time(double s) : t(s) {}
  static Time now() { return Time(std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count()); }
  double toSec() const { return t; }
  Time& fromSec(double s) { t = s; return *this; }
};
}  // namespace ros
#endif
