// ref_map_shim.cc -- TEST INFRASTRUCTURE.  C entry points over the reference's own sub-map store,
// src/slam/map/hybrid_grid.cc compiled UNMODIFIED from the reference checkout (oracle/Makefile, target `ref`): the
// Cartographer-derived nested grid, HybridGrid::InsertScan (:503-521: append to the 3 m cells the scan touches, then
// re-filter each touched cell) and HybridGrid::GetSurroundedCloud (:470-501: the cells hit by the scan points within
// 60 m, +-1 m in every axis, at the given pose).  Stand-ins (oracle/ref_stubs/): Eigen's arrays, boost::unordered_set
// (std's), pcl::Filter / pcl::VoxelGrid (the oracle's restatement of the centroid filter), glog.  The reference
// concatenates the selected cells in hash order of their shared pointers (heap addresses), so callers compare the
// result as a multiset of points.
#include <pcl/filters/voxel_grid.h>

#include "slam/map/hybrid_grid.h"

namespace {
struct RefMap {
  HybridGrid grid;
  pcl::VoxelGrid<PointType> filter;
  RefMap(float resolution, float leaf) : grid(resolution) { filter.setLeafSize(leaf, leaf, leaf); }
};
PointCloudPtr cloud_of(const float *xyzi, int n) {
  PointCloudPtr c(new PointCloud);
  c->points.resize(n);
  for (int i = 0; i < n; ++i) {
    PointType &p = c->points[i];
    p.x = xyzi[4 * i], p.y = xyzi[4 * i + 1], p.z = xyzi[4 * i + 2], p.intensity = xyzi[4 * i + 3];
  }
  c->width = n;
  return c;
}
}  // namespace

extern "C" {
void *msflref_map_create(float resolution, float leaf) { return new RefMap(resolution, leaf); }
void msflref_map_free(void *m) { delete static_cast<RefMap *>(m); }
// laser_mapping.cc:330-338: the scan is already in the world frame
void msflref_map_insert(void *m, const float *scan_world_xyzi, int n) {
  RefMap *r = static_cast<RefMap *>(m);
  r->grid.InsertScan(cloud_of(scan_world_xyzi, n), r->filter);
}
// returns the number of points of the surrounding cloud; writes at most cap of them
int msflref_map_surround(void *m, const float *scan_xyzi, int n, const double pose[7], float *out_xyzi, int cap) {
  RefMap *r = static_cast<RefMap *>(m);
  const Rigid3d T(Eigen::Vector3d(pose[0], pose[1], pose[2]), Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]));
  PointCloudPtr c = r->grid.GetSurroundedCloud(cloud_of(scan_xyzi, n), T);
  const int k = (int)c->size();
  for (int i = 0; i < k && i < cap; ++i) {
    out_xyzi[4 * i] = c->points[i].x, out_xyzi[4 * i + 1] = c->points[i].y, out_xyzi[4 * i + 2] = c->points[i].z;
    out_xyzi[4 * i + 3] = c->points[i].intensity;
  }
  return k;
}
}
