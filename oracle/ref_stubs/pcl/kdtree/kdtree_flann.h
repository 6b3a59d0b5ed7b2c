// Stand-in for <pcl/kdtree/kdtree_flann.h> (TEST INFRASTRUCTURE).  pcl::KdTreeFLANN::nearestKSearch is an EXACT k-NN
// (FLANN KDTreeSingleIndex, L2_Simple<float>, checks = all leaves), so any exact search with the same distance
// arithmetic returns the same neighbours: this class answers with the oracle's kd-tree (msflo_kdtree_*, itself checked
// against OpenCV's bundled FLANN KDTreeSingleIndex in tests/test_oracle.py).  Every query and its result is appended to a
// process-wide log so the harness can hand the reference's sequence of searches to the tests.
#ifndef MSFL_PCL_KDTREE_STANDIN_H
#define MSFL_PCL_KDTREE_STANDIN_H
#include <memory>
#include <vector>

#include "../point_cloud.h"

extern "C" {
struct msflo_kdtree;
msflo_kdtree *msflo_kdtree_build(const float *xyzi, int n);
void msflo_kdtree_free(msflo_kdtree *t);
int msflo_kdtree_knn(const msflo_kdtree *t, const float q[3], int k, int *idx, float *d2);
}

namespace msfl_ref {
struct KnnLog {
  bool enabled = false;
  std::vector<int> idx;    // k indices per search, -1 padded
  std::vector<float> d2;   // k squared distances per search
  std::vector<int> k_of;   // k of every search
};
KnnLog &knn_log();  // defined in oracle/ref_shim.cc
}  // namespace msfl_ref

namespace pcl {
template <typename PointT>
class KdTreeFLANN {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> PointCloudConstPtr;
  KdTreeFLANN() = default;
  KdTreeFLANN(const KdTreeFLANN &) = delete;
  ~KdTreeFLANN() {
    if (tree_) msflo_kdtree_free(tree_);
  }
  void setInputCloud(const PointCloudConstPtr &cloud) {
    if (tree_) msflo_kdtree_free(tree_);
    xyzi_.resize(cloud->points.size() * 4);
    for (size_t i = 0; i < cloud->points.size(); ++i) {
      const PointT &p = cloud->points[i];
      xyzi_[4 * i] = p.x, xyzi_[4 * i + 1] = p.y, xyzi_[4 * i + 2] = p.z, xyzi_[4 * i + 3] = p.intensity;
    }
    tree_ = msflo_kdtree_build(xyzi_.data(), (int)cloud->points.size());
  }
  int nearestKSearch(const PointT &point, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const {
    k_indices.assign(k, 0);            // PCL resizes both outputs to k, FLANN fills what it finds
    k_sqr_distances.assign(k, 0.f);
    const float q[3] = {point.x, point.y, point.z};
    const int found = msflo_kdtree_knn(tree_, q, k, k_indices.data(), k_sqr_distances.data());
    msfl_ref::KnnLog &log = msfl_ref::knn_log();
    if (log.enabled) {
      log.k_of.push_back(k);
      for (int j = 0; j < k; ++j) {
        log.idx.push_back(j < found ? k_indices[j] : -1);
        log.d2.push_back(j < found ? k_sqr_distances[j] : 0.f);
      }
    }
    return found;
  }

 private:
  std::vector<float> xyzi_;
  msflo_kdtree *tree_ = nullptr;
};
}  // namespace pcl
#endif
