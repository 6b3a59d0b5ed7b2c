// Stand-in for <pcl/filters/filter.h> (TEST INFRASTRUCTURE): the interface HybridGrid::InsertScan filters a cell through.
#ifndef MSFL_PCL_FILTER_STANDIN_H
#define MSFL_PCL_FILTER_STANDIN_H
#include <memory>

#include "../point_cloud.h"
namespace pcl {
template <typename PointT>
class Filter {
 public:
  virtual ~Filter() {}
  virtual void setInputCloud(const std::shared_ptr<const PointCloud<PointT>> &cloud) = 0;
  virtual void filter(PointCloud<PointT> &out) = 0;
};
}  // namespace pcl
#endif
