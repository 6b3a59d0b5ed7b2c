// stand-in for <nav_msgs/Odometry.h> (TEST INFRASTRUCTURE)
#ifndef MSFL_NAV_MSGS_ODOMETRY_STANDIN_H
#define MSFL_NAV_MSGS_ODOMETRY_STANDIN_H
#include "../sensor_msgs/Imu.h"
namespace nav_msgs {
struct Odometry {
  std_msgs::Header header;
  geometry_msgs::PoseWithCovariance pose;
};
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
}  // namespace nav_msgs
#endif
