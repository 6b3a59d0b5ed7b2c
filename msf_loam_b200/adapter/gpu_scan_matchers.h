// gpu_scan_matchers.h -- reference-side adapters: drop-in subclasses of MSF_LOAM's matchers that
// forward to libmsfl.so through the C ABI (include/msfl.h).  This file is meant to be compiled
// INSIDE the reference tree (it includes its headers, PCL, Eigen and Ceres).  In this repository, where those
// libraries do not exist, it is compiled against the minimal declarations of tests/adapter_stubs/ and its
// marshalling is run against libmsfl.so by tests/test_adapter.py.  See INTEGRATION.md.
//
//   OdometryScanMatcher::MatchScan2Scan   odometry_scan_matcher.h:10-12   -> msfl_scan2scan
//   MappingScanMatcher::MatchScan2Map     mapping_scan_matcher.h:14-21    -> msfl_set_submap + msfl_scan2map
//   LaserMapping's frame (MatchScan2Map + InsertScan2Map, laser_mapping.cc:258-340) -> msfl_mapping_frame (GpuMappingFrame)
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "msfl.h"
#include "slam/imu_fusion/imu_factor.h"
#include "slam/imu_fusion/pose_local_parameterization.h"
#include "slam/local/scan_matching/mapping_scan_matcher.h"
#include "slam/local/scan_matching/odometry_scan_matcher.h"

namespace msfl_adapter {

template <typename PointT>
inline msfl_cloud View(const pcl::PointCloud<PointT> &c, bool with_ring);

template <>
inline msfl_cloud View<PointType>(const pcl::PointCloud<PointType> &c, bool) {  // pcl::PointXYZI
  return msfl_cloud{c.empty() ? nullptr : &c.points[0], c.size(), sizeof(PointType), offsetof(PointType, x),
                    offsetof(PointType, intensity), MSFL_NO_FIELD};
}
template <>
inline msfl_cloud View<PointTypeOriginal>(const pcl::PointCloud<PointTypeOriginal> &c, bool with_ring) {  // PointXYZIRT
  return msfl_cloud{c.empty() ? nullptr : &c.points[0], c.size(), sizeof(PointTypeOriginal),
                    offsetof(PointTypeOriginal, x), offsetof(PointTypeOriginal, intensity),
                    with_ring ? offsetof(PointTypeOriginal, ring) : MSFL_NO_FIELD};
}

inline void ToArray(const Rigid3d &T, double out[7]) {  // Rigid3d::ToVector7 order (rigid_transform.h:59-64)
  out[0] = T.translation().x(); out[1] = T.translation().y(); out[2] = T.translation().z();
  out[3] = T.rotation().x(); out[4] = T.rotation().y(); out[5] = T.rotation().z(); out[6] = T.rotation().w();
}
inline Rigid3d FromArray(const double in[7]) {
  return Rigid3d(Eigen::Vector3d(in[0], in[1], in[2]), Eigen::Quaterniond(in[6], in[3], in[4], in[5]));
}

class Engine {
 public:
  explicit Engine(int device = 0) {
    msfl_params p;
    msfl_default_params(&p);
    // the ROS node matches one scan at a time: a thread-block cluster of 16 CTAs (8 where a GPC cannot host 16) owns
    // the scan for the whole call -- association into shared memory + solve in ONE launch (scan2map_fused.cu);
    // batches keep the default of one CTA per scan
    p.lm_cluster = 16;
    if (msfl_create(&p, device, &e_) != MSFL_OK) throw std::runtime_error(std::string("msfl_create: ") + msfl_last_error());
  }
  ~Engine() { msfl_destroy(e_); }
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;
  msfl_engine *get() const { return e_; }

 private:
  msfl_engine *e_ = nullptr;
};

}  // namespace msfl_adapter

// Replaces `scan_matcher_(std::make_unique<OdometryScanMatcher>())` at laser_odometry.cc:56.
class GpuOdometryScanMatcher : public OdometryScanMatcher {
 public:
  bool MatchScan2Scan(const TimestampedPointCloud<PointTypeOriginal> &scan_last,
                      const TimestampedPointCloud<PointTypeOriginal> &scan_curr,
                      Rigid3d *pose_estimate_curr2last) override {
    using msfl_adapter::View;
    const msfl_cloud lc = View(*scan_last.cloud_corner_less_sharp, true), ls = View(*scan_last.cloud_surf_less_flat, true);
    const msfl_cloud cs = View(*scan_curr.cloud_corner_sharp, false), cf = View(*scan_curr.cloud_surf_flat, false);
    double pose[7];
    msfl_adapter::ToArray(*pose_estimate_curr2last, pose);
    const int rc = msfl_scan2scan(engine_.get(), &lc, &ls, &cs, &cf, pose, nullptr);
    CHECK_GE(rc, 0) << msfl_last_error();             // fatal errors abort, like the reference's glog CHECKs
    *pose_estimate_curr2last = msfl_adapter::FromArray(pose);
    return rc == MSFL_OK;                              // false <=> fewer than 10 correspondences (:262-267)
  }

 private:
  msfl_adapter::Engine engine_;
};

// Replaces `scan_matcher_(std::make_unique<MappingScanMatcher>())` at laser_mapping.cc:42.
class GpuMappingScanMatcher : public MappingScanMatcher {
 public:
  bool MatchScan2Map(const TimestampedPointCloud<PointType> &cloud_map, const TimestampedPointCloud<PointType> &scan_curr,
                     const bool is_initialized, const std::shared_ptr<IntegrationBase> &preintegration,
                     const Vector3d &gravity_vector, const RobotState &prev_state, Rigid3d *pose_estimate_map_scan2world,
                     Vector3d *velocity) override {
    using msfl_adapter::View;
    if (is_initialized) {
      // the IMU side-car stays on the host (one 15-residual block, not data-parallel)
      PredictWithImu(prev_state, pose_estimate_map_scan2world, velocity);
      const msfl_cloud mc = View(*cloud_map.cloud_corner_less_sharp, false), ms = View(*cloud_map.cloud_surf_less_flat, false);
      const msfl_cloud sc = View(*scan_curr.cloud_corner_less_sharp, false), ss = View(*scan_curr.cloud_surf_less_flat, false);
      CHECK_EQ(msfl_set_submap(engine_.get(), &mc, &ms), MSFL_OK) << msfl_last_error();
      // IntegrationBase buffers -> msfl_deskew (GetDeltaQP inputs, scan_undistortion.cc:22-42)
      const auto &pi = *preintegration;
      std::vector<double> dq(4 * pi.delta_q_buf_.size()), dp(3 * pi.delta_p_buf_.size());
      for (size_t i = 0; i < pi.delta_q_buf_.size(); ++i) {
        const auto &q = pi.delta_q_buf_[i];
        dq[4 * i] = q.x(); dq[4 * i + 1] = q.y(); dq[4 * i + 2] = q.z(); dq[4 * i + 3] = q.w();
        for (int k = 0; k < 3; ++k) dp[3 * i + k] = pi.delta_p_buf_[i][k];
      }
      msfl_deskew dk{pi.sum_dt_buf_.data(), dq.data(), dp.data(), (int32_t)pi.sum_dt_buf_.size(), 0,
                     {(*velocity)[0], (*velocity)[1], (*velocity)[2]},
                     {gravity_vector[0], gravity_vector[1], gravity_vector[2]}};
      double pose[7];
      msfl_adapter::ToArray(*pose_estimate_map_scan2world, pose);
      CHECK_EQ(msfl_scan2map_deskew(engine_.get(), &sc, &ss, &dk, pose, nullptr), MSFL_OK) << msfl_last_error();
      *pose_estimate_map_scan2world = msfl_adapter::FromArray(pose);
      return true;  // the speed-bias block is constant in the reference's problem (:94): *velocity is unchanged
    }
    const msfl_cloud mc = View(*cloud_map.cloud_corner_less_sharp, false), ms = View(*cloud_map.cloud_surf_less_flat, false);
    const msfl_cloud sc = View(*scan_curr.cloud_corner_less_sharp, false), ss = View(*scan_curr.cloud_surf_less_flat, false);
    CHECK_EQ(msfl_set_submap(engine_.get(), &mc, &ms), MSFL_OK) << msfl_last_error();  // the two kd-tree builds (:66-72)
    double pose[7];
    msfl_adapter::ToArray(*pose_estimate_map_scan2world, pose);
    CHECK_EQ(msfl_scan2map(engine_.get(), &sc, &ss, pose, nullptr), MSFL_OK) << msfl_last_error();
    *pose_estimate_map_scan2world = msfl_adapter::FromArray(pose);
    return true;  // :277
  }

 private:
  // The IMU-only predict that precedes the LiDAR factors when the estimator is initialised
  // (mapping_scan_matcher.cc:28-60): one IMUFactor between the previous state (constant) and pose_j / bias_j with the
  // two biases of bias_j held constant, 6 iterations; it sets *pose = pose_j and *velocity = bias_j.head<3>().  The
  // same problem is stated here with the reference's own factor and parameterisation classes, so the base class needs
  // no change beyond `virtual`.
  static void PredictWithImu(const RobotState &prev_state, Rigid3d *pose, Vector3d *velocity) {
    double pose_i[7], pose_j[7], bias_i[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bias_j[9];
    msfl_adapter::ToArray(Rigid3d{prev_state.p, prev_state.q}, pose_i);
    for (int k = 0; k < 3; ++k) bias_i[k] = prev_state.v[k];
    for (int k = 0; k < 7; ++k) pose_j[k] = pose_i[k];
    for (int k = 0; k < 9; ++k) bias_j[k] = bias_i[k];
    ceres::Problem problem;
    problem.AddResidualBlock(new IMUFactor(prev_state.imu_preintegration), nullptr, pose_i, bias_i, pose_j, bias_j);
    problem.SetParameterBlockConstant(pose_i);
    problem.SetParameterBlockConstant(bias_i);
    problem.AddParameterBlock(pose_j, 7, new PoseLocalParameterization);
    problem.AddParameterBlock(bias_j, 9, new ceres::SubsetParameterization(9, {3, 4, 5, 6, 7, 8}));
    ceres::Solver::Options options;
    options.max_num_iterations = 6;
    options.minimizer_progress_to_stdout = false;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    *pose = msfl_adapter::FromArray(pose_j);
    *velocity = Vector3d(bias_j[0], bias_j[1], bias_j[2]);
  }

  msfl_adapter::Engine engine_;
};

// Replaces the two HybridGrid members of LaserMapping (laser_mapping.h: hybrid_grid_map_corner_ / hybrid_grid_map_surf_,
// resolution 3 m, leaf = the two down-size filters' 0.2 / 0.4) AND the matcher for the LiDAR-only branch: one call per
// frame does LaserMapping::MatchScan2Map + InsertScan2Map (laser_mapping.cc:258-340) with every intermediate on the GPU.
class GpuMappingFrame {
 public:
  explicit GpuMappingFrame(float resolution = 3.0f, float leaf_corner = 0.2f, float leaf_surf = 0.4f) {
    if (msfl_map_create(engine_.get(), resolution, leaf_corner, &corner_) != MSFL_OK ||
        msfl_map_create(engine_.get(), resolution, leaf_surf, &surf_) != MSFL_OK)
      throw std::runtime_error(std::string("msfl_map_create: ") + msfl_last_error());
  }
  ~GpuMappingFrame() {
    msfl_map_destroy(corner_);
    msfl_map_destroy(surf_);
  }
  GpuMappingFrame(const GpuMappingFrame &) = delete;
  GpuMappingFrame &operator=(const GpuMappingFrame &) = delete;

  // odom_result.cloud_corner_less_sharp / cloud_surf_less_flat (un-down-sampled) in, pose_map_scan2world_ in-out;
  // returns false when the surround clouds were too small to match (:284-285, :313) -- the scan is inserted either way
  template <typename PointT>
  bool MatchAndInsert(const pcl::PointCloud<PointT> &corner_less_sharp, const pcl::PointCloud<PointT> &surf_less_flat,
                      Rigid3d *pose_map_scan2world) {
    const msfl_cloud c = msfl_adapter::View(corner_less_sharp, false), s = msfl_adapter::View(surf_less_flat, false);
    double pose[7];
    msfl_adapter::ToArray(*pose_map_scan2world, pose);
    int32_t matched = 0;
    CHECK_EQ(msfl_mapping_frame(engine_.get(), corner_, surf_, &c, &s, pose, &matched, nullptr), MSFL_OK) << msfl_last_error();
    *pose_map_scan2world = msfl_adapter::FromArray(pose);
    return matched != 0;
  }

 private:
  msfl_adapter::Engine engine_;
  msfl_map *corner_ = nullptr, *surf_ = nullptr;
};
