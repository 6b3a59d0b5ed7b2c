// Replaces the reference's slam/msg_conversion.h (ROS <-> internal conversions; callers' side, out of scope) with the
// few overloads msf_loam_node.cc names (TEST INFRASTRUCTURE); included first, claims the real header's include guard.
#ifndef MSF_LOAM_VELODYNE_MSG_CONVERSION_H
#define MSF_LOAM_VELODYNE_MSG_CONVERSION_H
#include "common/rigid_transform.h"
#include "common/time.h"
#include "nav_msgs/Odometry.h"
#include "proto/config.pb.h"
inline Time FromROS(const ros::Time &) { return Time(); }
inline Vector3d FromROS(const geometry_msgs::Vector3 &v) { return Vector3d(v.x, v.y, v.z); }
inline Rigid3d FromROS(const geometry_msgs::PoseWithCovariance &) { return Rigid3d(); }
inline Rigid3d FromProto(const proto::Rigid3d &) { return Rigid3d(); }
#endif
