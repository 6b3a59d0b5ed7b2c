// ref_factors_shim.cc -- TEST INFRASTRUCTURE.  C entry points over the REFERENCE's own factor / parameterisation
// classes, compiled from the sources where they lie under /root/reference (oracle/Makefile, target
// _ref/libmsfl_ref_factors.so):
//   src/slam/local/scan_matching/lidar_factor.cc     LidarEdgeFactorSE3 / LidarPlaneFactorSE3 (:7-44) and the Deskew
//                                                    variants (:46-100)
//   src/slam/imu_fusion/pose_local_parameterization.cc   Plus (:6-21, through Utility::deltaQ utility.h:8-31),
//                                                    ComputeJacobian (:23-27)
// Eigen and Ceres are absent from the image: the reference sources are compiled against the stand-in headers in
// oracle/ref_stubs/ (what that does and does not pin is stated in msfl_eigen_standin.h).  tests/test_ref_factors.py
// checks the oracle's restatement -- and, on the GPU box, msfl_accumulate -- against these functions.
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
