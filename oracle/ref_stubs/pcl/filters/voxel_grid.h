// Stand-in for <pcl/filters/voxel_grid.h> (TEST INFRASTRUCTURE).  filter() delegates to the oracle's restatement of
// pcl::VoxelGrid<PointXYZI> (msflo_voxel_grid); getIndices() is PCLBase::getIndices(): with no indices set by the caller,
// initCompute() fills indices_ with 0..n-1 -- ALL input points, not the voxel representatives -- which is what makes
// VoxelGridWrapper (msf_loam_node.cc:113-126) an identity copy.
#ifndef MSFL_PCL_VOXEL_GRID_STANDIN_H
#define MSFL_PCL_VOXEL_GRID_STANDIN_H
#include <memory>
#include <vector>

#include "../common/copy_point.h"
#include "filter.h"
extern "C" int msflo_voxel_grid(const float *xyzi, int n, float leaf, float *out_xyzi);
namespace pcl {
typedef std::shared_ptr<std::vector<int>> IndicesPtr;
template <typename PointT>
class VoxelGrid : public Filter<PointT> {
 public:
  void setInputCloud(const std::shared_ptr<const PointCloud<PointT>> &cloud) override { input_ = cloud; }
  void setLeafSize(float lx, float, float) { leaf_ = lx; }
  void filter(PointCloud<PointT> &out) override {  // safe in place: the input is read completely first
    const int n = (int)input_->points.size();
    indices_.reset(new std::vector<int>(n));
    for (int i = 0; i < n; ++i) (*indices_)[i] = i;
    std::vector<float> in(4 * (size_t)n + 4), res(4 * (size_t)n + 4);
    for (int i = 0; i < n; ++i) {
      const PointT &p = input_->points[i];
      in[4 * i] = p.x, in[4 * i + 1] = p.y, in[4 * i + 2] = p.z, in[4 * i + 3] = p.intensity;
    }
    const int m = n ? msflo_voxel_grid(in.data(), n, leaf_, res.data()) : 0;
    out.points.resize(m);
    for (int i = 0; i < m; ++i) {
      PointT &p = out.points[i];
      p.x = res[4 * i], p.y = res[4 * i + 1], p.z = res[4 * i + 2], p.intensity = res[4 * i + 3];
    }
    out.width = m, out.height = 1;
  }
  IndicesPtr getIndices() const { return indices_; }

 private:
  std::shared_ptr<const PointCloud<PointT>> input_;
  IndicesPtr indices_;
  float leaf_ = 0.f;
};
template <typename PointT>
inline void copyPointCloud(const PointCloud<PointT> &in, const std::vector<int> &indices, PointCloud<PointT> &out) {
  out.points.resize(indices.size());
  for (size_t i = 0; i < indices.size(); ++i) out.points[i] = in.points[indices[i]];
  out.width = (std::uint32_t)indices.size(), out.height = 1, out.is_dense = in.is_dense;
}
}  // namespace pcl
#endif
