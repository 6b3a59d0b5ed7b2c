// Replaces the reference's slam/local/laser_odometry.h (the odometry / mapping driver: the CALLER of the hot path, out of
// scope) with a recorder: the scan RealHandleLaserCloudMessage registers is kept for the harness (TEST INFRASTRUCTURE).
// msf_loam_node.cc includes the real header by a path relative to its own directory, which no -I order can shadow, so
// this file is included first and claims the real header's include guard.
#ifndef MSF_LOAM_VELODYNE_LASER_ODOMETRY_H
#define MSF_LOAM_VELODYNE_LASER_ODOMETRY_H
#include "common/timestamped_pointcloud.h"
#include "proto/config.pb.h"
#include "slam/imu_fusion/types.h"
class LaserOdometry {
 public:
  explicit LaserOdometry(bool, proto::MsfLoamConfig) {}
  void AddLaserScan(const TimestampedPointCloud<PointTypeOriginal> &scan) {
    last_scan = scan;
    ++n_scans;
  }
  void AddImu(const ImuData &) {}
  void AddOdom(const OdometryData &) {}
  TimestampedPointCloud<PointTypeOriginal> last_scan;
  int n_scans = 0;
};
#endif
