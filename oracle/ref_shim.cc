// ref_shim.cc -- TEST INFRASTRUCTURE.  C entry points over the REFERENCE's own scan-matching code, compiled from the
// sources where they lie under /root/reference (oracle/Makefile, target `ref` -> _ref/libmsfl_ref.so):
//   src/slam/local/scan_matching/lidar_factor.cc           the four factors (:7-100)                      rows a-8, f-3
//   src/slam/imu_fusion/pose_local_parameterization.cc     Plus / ComputeJacobian (:6-27, utility.h:8-31)  row a-9
//   src/slam/local/scan_matching/odometry_scan_matcher.cc  OdometryScanMatcher::MatchScan2Scan (:43-285)   row a-5
//   src/slam/local/scan_matching/mapping_scan_matcher.cc   MappingScanMatcher::MatchScan2Map (:19-278)     rows a-6, a-7, f-3
//   src/slam/local/scan_matching/scan_matcher.cc           RefineByRejectOutliersWithThreshold (:13-38)    row a-10
//   src/slam/imu_fusion/scan_undistortion.cc               GetDeltaQP (:22-42)                             row f-3
// PCL, FLANN, Eigen, Ceres and glog are absent from the image: the reference sources are compiled UNMODIFIED against the
// stand-in headers in oracle/ref_stubs/.  What is the reference's and what is not:
//   reference's own : every association loop, gate, threshold, fit expression, factor, parameter-block set-up, iteration
//                     cap, the order of the two outer iterations, what is written back to the pose;
//   restated        : the third-party numerics underneath -- exact k-NN (the oracle's kd-tree), 3x3 eigen / 5x3 QR
//                     (the oracle's), Eigen's fixed-size algebra (ref_stubs/msfl_eigen_standin.h), and the Ceres
//                     trust-region loop (ceres::Solve below drives the oracle's msflo_lm_solve_cb with the reference's
//                     own CostFunction / LossFunction / LocalParameterization objects);
//   supplied here   : the two IMU side-car definitions the compiled files reference but whose sources (integration_base.cc,
//                     imu_factor.cc: IMU preintegration, out of scope) are not compiled -- an IntegrationBase constructor
//                     and IMUFactor::Evaluate; the IMU-only predict problem (mapping_scan_matcher.cc:35-60) is declined by
//                     the stand-in solver, so the Deskew entry takes the pose AFTER that predict, like msfl_scan2map_deskew.
// tests/test_ref_factors.py and tests/test_ref_matchers.py check the oracle -- and, on the GPU box, the CUDA path through
// the C ABI -- against these functions.  The same library also holds the reference's scan registration
// (ref_extract_shim.cc: src/msf_loam_node.cc, rows a-1 .. a-4) and its sub-map store (ref_map_shim.cc:
// src/slam/map/hybrid_grid.cc, row f-1).
#include <pcl/kdtree/kdtree_flann.h>

#include <atomic>
#include <cstring>
#include <thread>

#include "msfl_oracle.h"
#include "slam/imu_fusion/imu_factor.h"
#include "slam/local/scan_matching/mapping_scan_matcher.h"
#include "slam/local/scan_matching/odometry_scan_matcher.h"
#include "slam/imu_fusion/pose_local_parameterization.h"
#include "slam/local/scan_matching/lidar_factor.h"

namespace {
Eigen::Vector3d v3(const double *p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
}  // namespace

extern "C" {

// residual r[3] and the 3x7 row-major Jacobian w.r.t. the pose block [t xyz, q xyzw]
void msflref_edge_factor(const double pose[7], const double p[3], const double C[3], const double N[3], double r[3], double J[21]) {
  LidarEdgeFactorSE3 f(v3(p), v3(C), v3(N));
  const double *params[1] = {pose};
  double *jac[1] = {J};
  static_cast<const ceres::CostFunction &>(f).Evaluate(params, r, jac);
}

void msflref_plane_factor(const double pose[7], const double p[3], const double C[3], const double N[3], double r[1], double J[7]) {
  LidarPlaneFactorSE3 f(v3(p), v3(C), v3(N));
  const double *params[1] = {pose};
  double *jac[1] = {J};
  static_cast<const ceres::CostFunction &>(f).Evaluate(params, r, jac);
}

// Deskew variants: speed_bias[9] (velocity first); dq is x y z w; Jb = Jacobian w.r.t. the speed-bias block
void msflref_edge_factor_deskew(const double pose[7], const double speed_bias[9], const double p[3], const double C[3],
                                const double N[3], const double dp[3], const double dq[4], double dt, const double G[3],
                                double r[3], double J[21], double Jb[27]) {
  LidarEdgeFactorDeskewSE3 f(v3(p), v3(C), v3(N), v3(dp), Eigen::Quaterniond(dq[3], dq[0], dq[1], dq[2]), dt, v3(G));
  const double *params[2] = {pose, speed_bias};
  double *jac[2] = {J, Jb};
  static_cast<const ceres::CostFunction &>(f).Evaluate(params, r, jac);
}

void msflref_plane_factor_deskew(const double pose[7], const double speed_bias[9], const double p[3], const double C[3],
                                 const double N[3], const double dp[3], const double dq[4], double dt, const double G[3],
                                 double r[1], double J[7], double Jb[9]) {
  LidarPlaneFactorDeskewSE3 f(v3(p), v3(C), v3(N), v3(dp), Eigen::Quaterniond(dq[3], dq[0], dq[1], dq[2]), dt, v3(G));
  const double *params[2] = {pose, speed_bias};
  double *jac[2] = {J, Jb};
  static_cast<const ceres::CostFunction &>(f).Evaluate(params, r, jac);
}

// PoseLocalParameterization declares its overrides private: call them through the Ceres interface, as Ceres does
void msflref_pose_plus(const double x[7], const double delta[6], double out[7]) {
  PoseLocalParameterization lp;
  static_cast<const ceres::LocalParameterization &>(lp).Plus(x, delta, out);
}

void msflref_pose_plus_jacobian(const double x[7], double J[42]) {
  PoseLocalParameterization lp;
  static_cast<const ceres::LocalParameterization &>(lp).ComputeJacobian(x, J);
}

int msflref_pose_sizes(void) {
  PoseLocalParameterization lp;
  const ceres::LocalParameterization &b = lp;
  return b.GlobalSize() * 100 + b.LocalSize();
}
}

// ---------------------------------------------------------------------------------------------------------------
// side-car definitions (see the header comment)
// ---------------------------------------------------------------------------------------------------------------
IntegrationBase::IntegrationBase(const Eigen::Vector3d &acc0, const Eigen::Vector3d &gyr0, const Eigen::Vector3d &linearized_ba,
                                 const Eigen::Vector3d &linearized_bg)
    : dt_(0), acc0_(acc0), gyr0_(gyr0), linearized_acc_(acc0), linearized_gyr_(gyr0), linearized_ba_(linearized_ba),
      linearized_bg_(linearized_bg), sum_dt_(0) {}

bool IMUFactor::Evaluate(double const *const *, double *, double **) const { return false; }  // never evaluated: see Solve

namespace msfl_ref {
// per thread: the batch driver below runs independent scans on several host threads
KnnLog &knn_log() {
  static thread_local KnnLog log;
  return log;
}
struct SolveRecord {
  int supported, n_edge, n_plane;
  msflo_lm_log log;
};
static std::vector<SolveRecord> &solve_log() {
  static thread_local std::vector<SolveRecord> v;
  return v;
}
static bool g_keep_solve_log = true;  // off inside the timed batch driver
// benchmark schedule: the reference asks Ceres for max_num_iterations = 6 with its termination tests on; bench.py's
// workload is "2 x 5 attempts, fixed"  for both arms, which the stand-in solver can be told to follow
static int g_fixed_attempts = 0;  // > 0: early exit off, this many attempts per solve
}  // namespace msfl_ref

// ---------------------------------------------------------------------------------------------------------------
// ceres::Solve stand-in: the oracle's trust-region loop over the REFERENCE's cost / loss / parameterisation objects
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct EvalCtx {
  ceres::Problem *problem;
  double *pose_block;
  const ceres::LocalParameterization *plus;
};

// cost, H = J^T J, g = J^T r at `pose` the way Ceres assembles them: residual_block.cc (global Jacobian x local
// parameterisation Jacobian) then corrector.cc (loss-function rescaling of residual and Jacobian)
void eval_problem(const msflo_params *, const void *vctx, const double pose[7], double *cost_out, double H[36], double g[6]) {
  const EvalCtx *c = static_cast<const EvalCtx *>(vctx);
  const bool want = H || g;
  double Jp[42];
  if (want) c->plus->ComputeJacobian(pose, Jp);  // 7 x 6 row-major
  if (H) std::memset(H, 0, 36 * sizeof(double));
  if (g) std::memset(g, 0, 6 * sizeof(double));
  double cost = 0;
  for (const ceres::ResidualBlock *b : c->problem->residual_blocks()) {
    const int nres = b->cost->num_residuals();
    const double *params[4];
    double *jac[4];
    double r[16], Jpose[16 * 7];
    for (size_t i = 0; i < b->parameters.size(); ++i) {
      const bool is_pose = b->parameters[i] == c->pose_block;
      params[i] = is_pose ? pose : b->parameters[i];
      jac[i] = is_pose ? Jpose : nullptr;  // constant blocks: Ceres passes NULL for their Jacobians
    }
    b->cost->Evaluate(params, r, want ? jac : nullptr);
    double s = 0;
    for (int k = 0; k < nres; ++k) s += r[k] * r[k];
    double rho[3] = {s, 1.0, 0.0};
    if (b->loss) b->loss->Evaluate(s, rho);
    cost += 0.5 * rho[0];
    if (!want) continue;
    // Corrector::Corrector
    const double sqrt_rho1 = std::sqrt(rho[1]);
    double residual_scaling = sqrt_rho1, alpha_sq_norm = 0.0;
    if (!(s == 0.0 || rho[2] <= 0.0)) {
      const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
      const double alpha = 1.0 - std::sqrt(D);
      residual_scaling = sqrt_rho1 / (1 - alpha);
      alpha_sq_norm = alpha / s;
    }
    double Jl[16 * 6];
    for (int k = 0; k < nres; ++k)
      for (int j = 0; j < 6; ++j) {
        double a = 0;
        for (int m = 0; m < 7; ++m) a += Jpose[k * 7 + m] * Jp[m * 6 + j];
        Jl[k * 6 + j] = a;
      }
    if (alpha_sq_norm == 0.0) {  // Corrector::CorrectJacobian
      for (int k = 0; k < nres * 6; ++k) Jl[k] *= sqrt_rho1;
    } else {
      for (int j = 0; j < 6; ++j) {
        double rtj = 0;
        for (int k = 0; k < nres; ++k) rtj += Jl[k * 6 + j] * r[k];
        for (int k = 0; k < nres; ++k) Jl[k * 6 + j] = sqrt_rho1 * (Jl[k * 6 + j] - alpha_sq_norm * r[k] * rtj);
      }
    }
    for (int k = 0; k < nres; ++k) {
      const double rk = r[k] * residual_scaling;  // Corrector::CorrectResiduals
      if (g)
        for (int j = 0; j < 6; ++j) g[j] += Jl[k * 6 + j] * rk;
      if (H)
        for (int u = 0; u < 6; ++u)
          for (int v = 0; v < 6; ++v) H[u * 6 + v] += Jl[k * 6 + u] * Jl[k * 6 + v];
    }
  }
  *cost_out = cost;
}
}  // namespace

namespace ceres {
void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary) {
  msfl_ref::SolveRecord rec;
  std::memset(&rec, 0, sizeof rec);
  // the variable blocks the residuals touch
  std::set<double *> variable;
  bool sizes_ok = true;
  for (const ResidualBlock *b : problem->residual_blocks()) {
    if (b->cost->num_residuals() > 16 || b->parameters.size() > 4) sizes_ok = false;
    for (double *p : b->parameters) {
      const Problem::ParameterBlock *pb = problem->find(p);
      if (pb && !pb->constant) variable.insert(p);
    }
    if (b->cost->num_residuals() == 3) ++rec.n_edge;
    if (b->cost->num_residuals() == 1) ++rec.n_plane;
  }
  const Problem::ParameterBlock *pose = variable.size() == 1 ? problem->find(*variable.begin()) : nullptr;
  if (!sizes_ok || !pose || pose->size != 7 || !pose->parameterization || pose->parameterization->LocalSize() != 6) {
    // e.g. the IMU-only predict (four blocks, two of them free): IMU side-car, out of scope -- parameters untouched
    summary->message = "stand-in ceres::Solve: problem shape not supported, parameters left unchanged";
    if (msfl_ref::g_keep_solve_log) msfl_ref::solve_log().push_back(rec);
    return;
  }
  rec.supported = 1;
  msflo_params P;
  msflo_default_params(&P);  // Ceres' Solver::Options defaults (trust region, LM, Jacobi scaling, tolerances)
  P.max_num_iterations = options.max_num_iterations;
  if (msfl_ref::g_fixed_attempts > 0) {
    P.max_num_iterations = msfl_ref::g_fixed_attempts;
    P.early_exit = 0;
  }
  EvalCtx ctx{problem, pose->values, pose->parameterization};
  double x[7];
  std::memcpy(x, pose->values, sizeof x);
  msflo_lm_solve_cb(&P, eval_problem, &ctx, (int)problem->residual_blocks().size(), x, &rec.log);
  std::memcpy(pose->values, x, sizeof x);
  summary->message = "stand-in ceres::Solve: oracle trust-region loop over the reference's cost functions";
  if (msfl_ref::g_keep_solve_log) msfl_ref::solve_log().push_back(rec);
}
}  // namespace ceres

// ---------------------------------------------------------------------------------------------------------------
// matcher entry points: clouds are n x 4 float (x y z intensity), poses [t xyz, q xyzw]
// ---------------------------------------------------------------------------------------------------------------
namespace {
PointCloudPtr cloud_from(const float *xyzi, int n) {
  PointCloudPtr c(new PointCloud);
  c->points.resize(n);
  for (int i = 0; i < n; ++i) {
    PointType &p = c->points[i];
    p.x = xyzi[4 * i], p.y = xyzi[4 * i + 1], p.z = xyzi[4 * i + 2], p.intensity = xyzi[4 * i + 3];
  }
  c->width = n;
  return c;
}
PointCloudOriginalPtr cloud_from(const float *xyzi, const uint16_t *ring, int n) {
  PointCloudOriginalPtr c(new PointCloudOriginal);
  c->points.resize(n);
  for (int i = 0; i < n; ++i) {
    PointTypeOriginal &p = c->points[i];
    p.x = xyzi[4 * i], p.y = xyzi[4 * i + 1], p.z = xyzi[4 * i + 2], p.intensity = xyzi[4 * i + 3];
    p.ring = ring ? ring[i] : 0;
    p.time = 0.f;
  }
  c->width = n;
  return c;
}
Rigid3d rigid_from(const double pose[7]) {
  return Rigid3d(Eigen::Vector3d(pose[0], pose[1], pose[2]), Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]));
}
void rigid_to(const Rigid3d &T, double pose[7]) {
  pose[0] = T.translation().x(), pose[1] = T.translation().y(), pose[2] = T.translation().z();
  pose[3] = T.rotation().x(), pose[4] = T.rotation().y(), pose[5] = T.rotation().z(), pose[6] = T.rotation().w();
}
std::shared_ptr<IntegrationBase> preintegration_from(const double *sum_dt, const double *dq, const double *dp, int n) {
  const Eigen::Vector3d z = Eigen::Vector3d::Zero();
  std::shared_ptr<IntegrationBase> pre(new IntegrationBase(z, z, z, z));
  for (int i = 0; i < n; ++i) {
    pre->sum_dt_buf_.push_back(sum_dt[i]);
    pre->delta_q_buf_.push_back(Eigen::Quaterniond(dq[4 * i + 3], dq[4 * i], dq[4 * i + 1], dq[4 * i + 2]));
    pre->delta_p_buf_.push_back(Eigen::Vector3d(dp[3 * i], dp[3 * i + 1], dp[3 * i + 2]));
  }
  return pre;
}
int run_scan2map(const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf, const float *scan_corner,
                 int n_scan_corner, const float *scan_surf, int n_scan_surf, bool is_initialized,
                 const std::shared_ptr<IntegrationBase> &pre, const double V[3], const double G[3], double pose[7], double vel_out[3]) {
  TimestampedPointCloud<PointType> cloud_map, scan_curr;
  cloud_map.cloud_corner_less_sharp = cloud_from(map_corner, n_map_corner);
  cloud_map.cloud_surf_less_flat = cloud_from(map_surf, n_map_surf);
  scan_curr.cloud_corner_less_sharp = cloud_from(scan_corner, n_scan_corner);
  scan_curr.cloud_surf_less_flat = cloud_from(scan_surf, n_scan_surf);
  RobotState prev;  // the state the IMU-only predict starts from; that predict is declined, so pose_j = this pose
  prev.p = Eigen::Vector3d(pose[0], pose[1], pose[2]);
  prev.q = Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]);
  prev.v = Eigen::Vector3d(V[0], V[1], V[2]);
  prev.imu_preintegration = pre;
  Rigid3d T = rigid_from(pose);
  Vector3d velocity(V[0], V[1], V[2]);
  const Vector3d gravity(G[0], G[1], G[2]);
  MappingScanMatcher matcher;
  const bool ok = matcher.MatchScan2Map(cloud_map, scan_curr, is_initialized, pre, gravity, prev, &T, &velocity);
  rigid_to(T, pose);
  if (vel_out) vel_out[0] = velocity.x(), vel_out[1] = velocity.y(), vel_out[2] = velocity.z();
  return ok ? 1 : 0;
}
}  // namespace

extern "C" {

// MappingScanMatcher::MatchScan2Map, LiDAR-only branch (is_initialized == false).  The reference calls GetDeltaQP for
// every point even then (mapping_scan_matcher.cc:115,185) and CHECK-fails outside the preintegration window, so the
// harness hands it an identity preintegration that spans every dt.
int msflref_scan2map(const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf, const float *scan_corner,
                     int n_scan_corner, const float *scan_surf, int n_scan_surf, double pose[7]) {
  const double sum_dt[2] = {-1e6, 1e6}, dq[8] = {0, 0, 0, 1, 0, 0, 0, 1}, dp[6] = {0, 0, 0, 0, 0, 0}, zero[3] = {0, 0, 0};
  return run_scan2map(map_corner, n_map_corner, map_surf, n_map_surf, scan_corner, n_scan_corner, scan_surf, n_scan_surf, false,
                      preintegration_from(sum_dt, dq, dp, 2), zero, zero, pose, nullptr);
}

// B independent MatchScan2Map calls (LiDAR-only branch) against one submap on n_threads host threads -- the CPU arm of
// bench.py (--impl reference).  Every call is the reference's: it builds its two kd-trees from the map clouds like the
// reference does every frame (mapping_scan_matcher.cc:66-72).  fixed_attempts > 0 = the benchmark schedule (see above).
int msflref_scan2map_batch(const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf, int B,
                           const float *scan_corner, const int *corner_off, const float *scan_surf, const int *surf_off,
                           double *poses, int n_threads, int fixed_attempts) {
  TimestampedPointCloud<PointType> cloud_map;
  cloud_map.cloud_corner_less_sharp = cloud_from(map_corner, n_map_corner);
  cloud_map.cloud_surf_less_flat = cloud_from(map_surf, n_map_surf);
  const double sum_dt[2] = {-1e6, 1e6}, dq[8] = {0, 0, 0, 1, 0, 0, 0, 1}, dp[6] = {0, 0, 0, 0, 0, 0};
  const std::shared_ptr<IntegrationBase> pre = preintegration_from(sum_dt, dq, dp, 2);
  msfl_ref::g_fixed_attempts = fixed_attempts;
  msfl_ref::g_keep_solve_log = false;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int b = next.fetch_add(1); b < B; b = next.fetch_add(1)) {
      TimestampedPointCloud<PointType> scan_curr;
      scan_curr.cloud_corner_less_sharp = cloud_from(scan_corner + 4 * (size_t)corner_off[b], corner_off[b + 1] - corner_off[b]);
      scan_curr.cloud_surf_less_flat = cloud_from(scan_surf + 4 * (size_t)surf_off[b], surf_off[b + 1] - surf_off[b]);
      double *pose = poses + 7 * (size_t)b;
      RobotState prev;
      prev.p = Eigen::Vector3d(pose[0], pose[1], pose[2]);
      prev.q = Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]);
      prev.v = Eigen::Vector3d(0, 0, 0);
      prev.imu_preintegration = pre;
      Rigid3d T = rigid_from(pose);
      Vector3d velocity(0, 0, 0);
      MappingScanMatcher matcher;
      matcher.MatchScan2Map(cloud_map, scan_curr, false, pre, Vector3d(0, 0, 0), prev, &T, &velocity);
      rigid_to(T, pose);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (std::thread &t : pool) t.join();
  msfl_ref::g_fixed_attempts = 0;
  msfl_ref::g_keep_solve_log = true;
  return B;
}

// the Deskew branch (is_initialized == true); pose = the pose after the IMU-only predict, V = bias_j.head<3>()
int msflref_scan2map_deskew(const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                            const float *scan_corner, int n_scan_corner, const float *scan_surf, int n_scan_surf,
                            const double *sum_dt, const double *delta_q, const double *delta_p, int n_pre, const double V[3],
                            const double G[3], double pose[7], double vel_out[3]) {
  return run_scan2map(map_corner, n_map_corner, map_surf, n_map_surf, scan_corner, n_scan_corner, scan_surf, n_scan_surf, true,
                      preintegration_from(sum_dt, delta_q, delta_p, n_pre), V, G, pose, vel_out);
}

// GetDeltaQP (scan_undistortion.cc:22-42); dq out is x y z w
void msflref_get_delta_qp(const double *sum_dt, const double *delta_q, const double *delta_p, int n_pre, double dt, double dq[4],
                          double dp[3]) {
  const Rigid3d r = GetDeltaQP(preintegration_from(sum_dt, delta_q, delta_p, n_pre), dt);
  dq[0] = r.rotation().x(), dq[1] = r.rotation().y(), dq[2] = r.rotation().z(), dq[3] = r.rotation().w();
  dp[0] = r.translation().x(), dp[1] = r.translation().y(), dp[2] = r.translation().z();
}

// OdometryScanMatcher::MatchScan2Scan; returns 1 / 0 = the reference's bool
int msflref_scan2scan(const float *last_corner, const uint16_t *last_corner_ring, int n_last_corner, const float *last_surf,
                      const uint16_t *last_surf_ring, int n_last_surf, const float *curr_sharp, int n_curr_sharp,
                      const float *curr_flat, int n_curr_flat, double pose[7]) {
  TimestampedPointCloud<PointTypeOriginal> scan_last, scan_curr;
  scan_last.cloud_corner_less_sharp = cloud_from(last_corner, last_corner_ring, n_last_corner);
  scan_last.cloud_surf_less_flat = cloud_from(last_surf, last_surf_ring, n_last_surf);
  scan_curr.cloud_corner_sharp = cloud_from(curr_sharp, nullptr, n_curr_sharp);
  scan_curr.cloud_surf_flat = cloud_from(curr_flat, nullptr, n_curr_flat);
  Rigid3d T = rigid_from(pose);
  OdometryScanMatcher matcher;
  const bool ok = matcher.MatchScan2Scan(scan_last, scan_curr, &T);
  rigid_to(T, pose);
  return ok ? 1 : 0;
}

// TransformPoint (rigid_transform.h:132-138): the query transform of rows a-4 / a-6 / a-7
void msflref_transform_point(const double pose[7], const float in[3], float out[3]) {
  PointType p;
  p.x = in[0], p.y = in[1], p.z = in[2];
  const PointType q = TransformPoint(rigid_from(pose), p);
  out[0] = q.x, out[1] = q.y, out[2] = q.z;
}

// ---- what the stand-in solver and kd-tree saw during the calls above ----
void msflref_log_reset(int record_knn) {
  msfl_ref::solve_log().clear();
  msfl_ref::KnnLog &k = msfl_ref::knn_log();
  k.enabled = record_knn != 0;
  k.idx.clear(), k.d2.clear(), k.k_of.clear();
}
int msflref_n_solves(void) { return (int)msfl_ref::solve_log().size(); }
int msflref_solve_info(int i, int *supported, int *n_edge, int *n_plane, msflo_lm_log *log) {
  if (i < 0 || i >= (int)msfl_ref::solve_log().size()) return -1;
  const msfl_ref::SolveRecord &r = msfl_ref::solve_log()[i];
  *supported = r.supported, *n_edge = r.n_edge, *n_plane = r.n_plane;
  if (log) *log = r.log;
  return 0;
}
int msflref_knn_log_searches(void) { return (int)msfl_ref::knn_log().k_of.size(); }
int msflref_knn_log_entries(void) { return (int)msfl_ref::knn_log().idx.size(); }
void msflref_knn_log_copy(int *k_of, int *idx, float *d2) {
  const msfl_ref::KnnLog &k = msfl_ref::knn_log();
  std::memcpy(k_of, k.k_of.data(), k.k_of.size() * sizeof(int));
  std::memcpy(idx, k.idx.data(), k.idx.size() * sizeof(int));
  std::memcpy(d2, k.d2.data(), k.d2.size() * sizeof(float));
}
}
