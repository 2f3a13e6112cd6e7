// STAND-IN (test infrastructure): the fields of sensor_msgs::Imu that IMU_PRE::push_imu reads (preintegration.hpp:49-75).
#ifndef VXREF_SENSOR_MSGS_IMU
#define VXREF_SENSOR_MSGS_IMU
#include <memory>
#include "../ros/ros.h"
namespace sensor_msgs {
struct Vec3 { double x = 0, y = 0, z = 0; };
struct Header { ros::Time stamp; };
struct Imu { Header header; Vec3 angular_velocity, linear_acceleration; };
typedef std::shared_ptr<Imu> ImuPtr;
typedef std::shared_ptr<const Imu> ImuConstPtr;
}  // namespace sensor_msgs
#endif
