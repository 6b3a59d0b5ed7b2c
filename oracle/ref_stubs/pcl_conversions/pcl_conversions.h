// Stand-in for <pcl_conversions/pcl_conversions.h> (TEST INFRASTRUCTURE): pcl::fromROSMsg hands back the decoded cloud
// the stand-in message carries.
#ifndef MSFL_PCL_CONVERSIONS_STANDIN_H
#define MSFL_PCL_CONVERSIONS_STANDIN_H
#include "../pcl/point_cloud.h"
#include "../sensor_msgs/PointCloud2.h"
namespace pcl {
template <typename PointT>
void fromROSMsg(const sensor_msgs::PointCloud2 &msg, PointCloud<PointT> &cloud) {
  cloud = *static_cast<const PointCloud<PointT> *>(msg.decoded.get());
}
}  // namespace pcl
#endif
