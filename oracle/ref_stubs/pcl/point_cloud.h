// Stand-in for <pcl/point_cloud.h> (TEST INFRASTRUCTURE): a vector of points with PCL's member names.
#ifndef MSFL_PCL_POINT_CLOUD_STANDIN_H
#define MSFL_PCL_POINT_CLOUD_STANDIN_H
#include <memory>
#include <string>
#include <vector>

#include "point_types.h"
namespace pcl {
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  struct Header {
    std::uint64_t stamp = 0;
    std::string frame_id;
  } header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  void resize(size_t n) {
    points.resize(n);
    width = (std::uint32_t)n, height = 1;
  }
  const PointT &front() const { return points.front(); }
  PointCloud &operator+=(const PointCloud &o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (std::uint32_t)points.size(), height = 1;
    return *this;
  }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void reserve(size_t n) { points.reserve(n); }
  void push_back(const PointT &p) {
    points.push_back(p);
    width = (std::uint32_t)points.size();
  }
  PointT &operator[](size_t i) { return points[i]; }
  const PointT &operator[](size_t i) const { return points[i]; }
  auto begin() { return points.begin(); }
  auto end() { return points.end(); }
  auto begin() const { return points.begin(); }
  auto end() const { return points.end(); }
};
}  // namespace pcl
#endif
