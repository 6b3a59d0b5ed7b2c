// stub_reference.h -- the SMALLEST stand-ins for the reference / PCL / Eigen / Ceres / glog declarations that
// msf_loam_b200/adapter/gpu_scan_matchers.h touches, so the adapter can be compiled (and its marshalling run against
// libmsfl.so) in this image, where none of those libraries exist.  TEST INFRASTRUCTURE: signatures and field names
// follow the reference headers cited next to each item; nothing here computes anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// ---- Eigen (only what Rigid3d / RobotState / IntegrationBase expose to the adapter)
namespace Eigen {
struct Vector3d {
  double v[3] = {0, 0, 0};
  Vector3d() = default;
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double &operator[](int i) { return v[i]; }
  const double &operator[](int i) const { return v[i]; }
  static Vector3d Zero() { return Vector3d(); }
};
struct Quaterniond {
  double qx = 0, qy = 0, qz = 0, qw = 1;
  Quaterniond() = default;
  Quaterniond(double w, double x, double y, double z) : qx(x), qy(y), qz(z), qw(w) {}  // Eigen ctor order: w, x, y, z
  double x() const { return qx; }
  double y() const { return qy; }
  double z() const { return qz; }
  double w() const { return qw; }
  static Quaterniond Identity() { return Quaterniond(); }
};
}  // namespace Eigen
using Vector3d = Eigen::Vector3d;        // common.h:73
using Quaterniond = Eigen::Quaterniond;  // common.h

// ---- glog
struct MsflStubCheck {
  bool ok;
  std::ostringstream os;
  explicit MsflStubCheck(bool b) : ok(b) {}
  ~MsflStubCheck() {
    if (!ok) { std::fprintf(stderr, "CHECK failed: %s\n", os.str().c_str()); std::abort(); }
  }
  template <typename T> MsflStubCheck &operator<<(const T &t) { os << t; return *this; }
};
#define CHECK_GE(a, b) MsflStubCheck((a) >= (b))
#define CHECK_EQ(a, b) MsflStubCheck((a) == (b))

// ---- PCL point types: pcl::PointXYZI (32 B: xyz + pad, intensity at 16) and the reference's PointXYZIRT (common.h:44-62)
namespace pcl {
struct alignas(16) PointXYZI {
  float x, y, z, _pad;
  float intensity;
  float _pad2[3];
};
template <typename P>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<P>>;
  std::vector<P> points;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(size_t n) { points.resize(n); }
  P &operator[](size_t i) { return points[i]; }
};
}  // namespace pcl
struct alignas(16) PointXYZIRT {
  float x, y, z, _pad;
  float intensity;
  std::uint16_t ring;
  float time;
};
static_assert(sizeof(pcl::PointXYZI) == 32 && offsetof(pcl::PointXYZI, intensity) == 16, "pcl::PointXYZI layout");
static_assert(sizeof(PointXYZIRT) == 32 && offsetof(PointXYZIRT, ring) == 20, "PointXYZIRT layout (common.h:44-50)");
using PointType = pcl::PointXYZI;        // common.h:64
using PointTypeOriginal = ::PointXYZIRT; // common.h:69

// ---- common/rigid_transform.h:36-96
template <typename F>
class Rigid3 {
 public:
  Rigid3() = default;
  Rigid3(const Eigen::Vector3d &t, const Eigen::Quaterniond &q) : t_(t), q_(q) {}
  const Eigen::Vector3d &translation() const { return t_; }
  const Eigen::Quaterniond &rotation() const { return q_; }

 private:
  Eigen::Vector3d t_;
  Eigen::Quaterniond q_;
};
using Rigid3d = Rigid3<double>;

// ---- common/time.h, common/timestamped_pointcloud.h:11-48
using Time = long long;
template <typename T>
struct TimestampedPointCloud {
  using PointCloudTypePtr = typename pcl::PointCloud<T>::Ptr;
  Time time = 0;
  PointCloudTypePtr cloud_full_res{new pcl::PointCloud<T>}, cloud_corner_sharp{new pcl::PointCloud<T>},
      cloud_corner_less_sharp{new pcl::PointCloud<T>}, cloud_surf_flat{new pcl::PointCloud<T>},
      cloud_surf_less_flat{new pcl::PointCloud<T>};
};

// ---- slam/imu_fusion/integration_base.h:8-73 (the three buffers GetDeltaQP reads), estimator.h:10-19
struct IntegrationBase {
  std::vector<double> sum_dt_buf_;
  std::vector<Eigen::Vector3d> delta_p_buf_;
  std::vector<Eigen::Quaterniond> delta_q_buf_;
};
struct RobotState {
  Time time = 0;
  Vector3d p, v;
  Quaterniond q;
  Vector3d bg, ba;
  std::shared_ptr<IntegrationBase> imu_preintegration;
};

// ---- Ceres (declarations only; Solve is a no-op here)
namespace ceres {
struct CostFunction { virtual ~CostFunction() = default; };
struct LossFunction { virtual ~LossFunction() = default; };
struct LocalParameterization { virtual ~LocalParameterization() = default; };
struct SubsetParameterization : LocalParameterization {
  SubsetParameterization(int, const std::vector<int> &) {}
};
struct Problem {
  std::vector<std::unique_ptr<CostFunction>> costs;
  std::vector<std::unique_ptr<LocalParameterization>> params;
  void AddResidualBlock(CostFunction *c, LossFunction *, double *, double *, double *, double *) { costs.emplace_back(c); }
  void SetParameterBlockConstant(double *) {}
  void AddParameterBlock(double *, int, LocalParameterization *p) { params.emplace_back(p); }
};
struct Solver {
  struct Options { int max_num_iterations = 50; bool minimizer_progress_to_stdout = false; };
  struct Summary {};
};
inline void Solve(const Solver::Options &, Problem *, Solver::Summary *) {}
}  // namespace ceres
struct IMUFactor : ceres::CostFunction {  // slam/imu_fusion/imu_factor.h
  explicit IMUFactor(std::shared_ptr<IntegrationBase>) {}
};
struct PoseLocalParameterization : ceres::LocalParameterization {};  // pose_local_parameterization.h:5

// ---- the two matcher base classes: odometry_scan_matcher.h:8-13, mapping_scan_matcher.h:12-22 (+ `virtual`)
class ScanMatcher {
 public:
  virtual ~ScanMatcher() = default;
};
class OdometryScanMatcher : public ScanMatcher {
 public:
  virtual bool MatchScan2Scan(const TimestampedPointCloud<PointTypeOriginal> &, const TimestampedPointCloud<PointTypeOriginal> &,
                              Rigid3d *) { return false; }
};
class MappingScanMatcher : public ScanMatcher {
 public:
  virtual bool MatchScan2Map(const TimestampedPointCloud<PointType> &, const TimestampedPointCloud<PointType> &, const bool,
                             const std::shared_ptr<IntegrationBase> &, const Vector3d &, const RobotState &, Rigid3d *,
                             Vector3d *) { return false; }
};
