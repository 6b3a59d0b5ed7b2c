// Stand-in for <pcl/common/copy_point.h> (TEST INFRASTRUCTURE): pcl::copyPoint copies the fields the two point types
// share -- everything for one type, x y z intensity between PointXYZIRT and PointXYZI.
#ifndef MSFL_PCL_COPY_POINT_STANDIN_H
#define MSFL_PCL_COPY_POINT_STANDIN_H
#include "../point_cloud.h"
namespace pcl {
template <typename PointT>
inline void copyPoint(const PointT &in, PointT &out) { out = in; }
template <typename InT, typename OutT>
inline void copyPoint(const InT &in, OutT &out) {
  out.x = in.x, out.y = in.y, out.z = in.z;
  out.intensity = in.intensity;
}
template <typename InT, typename OutT>
inline void copyPointCloud(const PointCloud<InT> &in, PointCloud<OutT> &out) {
  out.points.resize(in.points.size());
  out.width = in.width, out.height = in.height, out.is_dense = in.is_dense;
  for (size_t i = 0; i < in.points.size(); ++i) copyPoint(in.points[i], out.points[i]);
}
}  // namespace pcl
#endif
