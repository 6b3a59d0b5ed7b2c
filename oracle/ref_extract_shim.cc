// ref_extract_shim.cc -- TEST INFRASTRUCTURE.  The reference's scan registration, src/msf_loam_node.cc, compiled WHOLE and
// UNMODIFIED: it is #included below from the reference checkout (its extraction parameters g_min_range /
// g_lidar2imu_transfrom live in an anonymous namespace, so the harness has to share the translation unit), with its
// main() renamed.  ROS, rosbag, protobuf, gflags, PCL and the generated proto/config.pb.h are absent from the image and
// are stood in by oracle/ref_stubs/ (empty shells: main() is never run); the reference's LaserOdometry driver -- the
// caller of the hot path -- is shadowed by a recorder that keeps the registered scan.
// What runs is the reference's own RealHandleLaserCloudMessage (msf_loam_node.cc:160-378): RemoveInvalidPointsFromCloud
// (:86-111), ComputeRelaTimeForEachPoint (:128-156), ring concatenation and margins (:188-195), curvature (:213-240),
// sector sort + greedy pick (:251-351), VoxelGridWrapper (:113-126) and the extrinsic (:367-371): rows a-1 .. a-4.
// Third-party pieces underneath: std::sort (this toolchain's libstdc++, like any build of the reference), atan2 (the
// float overload, see oracle/ref_harness/atan2_overload.cc -- <math.h> is in this include graph too), pcl::VoxelGrid
// (stand-in; the reference discards its result: quirk Q1).
#include "msfl_standin_laser_odometry.h"
#include "msfl_standin_msg_conversion.h"
#define main msf_loam_node_main
#include <math.h>  // the ROS / tf headers of the real include graph pull in <math.h>: unqualified atan2(float, float) is the float overload
#include "msf_loam_node.cc"
#undef main

namespace {
void copy_out(const PointCloudOriginal &c, float *xyzi, uint16_t *ring, int cap, int *n) {
  *n = (int)c.size();
  for (int i = 0; i < (int)c.size() && i < cap; ++i) {
    xyzi[4 * i] = c.points[i].x, xyzi[4 * i + 1] = c.points[i].y, xyzi[4 * i + 2] = c.points[i].z;
    xyzi[4 * i + 3] = c.points[i].intensity;
    if (ring) ring[i] = c.points[i].ring;
  }
}
}  // namespace

extern "C" {
// One RealHandleLaserCloudMessage call.  in: n raw points (x y z intensity, ring); T_ext = g_lidar2imu_transfrom
// [t xyz, q xyzw]; min_range = g_min_range.  out (capacity n points each): the registered full cloud (ring-major,
// intensity = relative time, extrinsic applied) with rings, and the four feature clouds in the reference's push order.
// Returns the number of scans handed to LaserOdometry::AddLaserScan (1).
int msflref_extract_features(const float *xyzi, const uint16_t *ring, int n, const double T_ext[7], double min_range,
                             float *full, uint16_t *full_ring, int *n_full, float *sharp, int *n_sharp, float *less_sharp,
                             int *n_less_sharp, float *flat, int *n_flat, float *less_flat, int *n_less_flat) {
  std::shared_ptr<PointCloudOriginal> raw(new PointCloudOriginal);
  raw->points.resize(n);
  for (int i = 0; i < n; ++i) {
    PointTypeOriginal &p = raw->points[i];
    p.x = xyzi[4 * i], p.y = xyzi[4 * i + 1], p.z = xyzi[4 * i + 2], p.intensity = xyzi[4 * i + 3];
    p.ring = ring[i];
    p.time = 0.f;
  }
  raw->width = n;
  std::shared_ptr<sensor_msgs::PointCloud2> msg(new sensor_msgs::PointCloud2);
  msg->decoded = raw;
  g_min_range = min_range;
  g_lidar2imu_transfrom = Rigid3d(Eigen::Vector3d(T_ext[0], T_ext[1], T_ext[2]), Eigen::Quaterniond(T_ext[6], T_ext[3], T_ext[4], T_ext[5]));
  std::shared_ptr<LaserOdometry> sink = std::make_shared<LaserOdometry>(true, proto::MsfLoamConfig());
  RealHandleLaserCloudMessage(msg, sink);
  const TimestampedPointCloud<PointTypeOriginal> &s = sink->last_scan;
  copy_out(*s.cloud_full_res, full, full_ring, n, n_full);
  copy_out(*s.cloud_corner_sharp, sharp, nullptr, n, n_sharp);
  copy_out(*s.cloud_corner_less_sharp, less_sharp, nullptr, n, n_less_sharp);
  copy_out(*s.cloud_surf_flat, flat, nullptr, n, n_flat);
  copy_out(*s.cloud_surf_less_flat, less_flat, nullptr, n, n_less_flat);
  return sink->n_scans;
}
}
