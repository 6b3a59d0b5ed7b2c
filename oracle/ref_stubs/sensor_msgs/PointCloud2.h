// Stand-in for <sensor_msgs/PointCloud2.h> (TEST INFRASTRUCTURE): the message carries an already-decoded
// PointXYZIRT cloud -- decoding the wire format is row f-4 (pcl::fromROSMsg, third-party), not part of this harness.
#ifndef MSFL_SENSOR_MSGS_POINTCLOUD2_STANDIN_H
#define MSFL_SENSOR_MSGS_POINTCLOUD2_STANDIN_H
#include "../ros/ros.h"
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs::Header header;
  std::shared_ptr<const void> decoded;  // pcl::PointCloud<PointXYZIRT>, see pcl_conversions stand-in
};
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}  // namespace sensor_msgs
#endif
