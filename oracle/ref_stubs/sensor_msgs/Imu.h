// stand-in for <sensor_msgs/Imu.h> (TEST INFRASTRUCTURE)
#ifndef MSFL_SENSOR_MSGS_IMU_STANDIN_H
#define MSFL_SENSOR_MSGS_IMU_STANDIN_H
#include "../ros/ros.h"
namespace geometry_msgs {
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct PoseWithCovariance {
  double position[3] = {0, 0, 0}, orientation[4] = {0, 0, 0, 1};
};
}  // namespace geometry_msgs
namespace sensor_msgs {
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Vector3 linear_acceleration, angular_velocity;
};
typedef std::shared_ptr<const Imu> ImuConstPtr;
}  // namespace sensor_msgs
#endif
