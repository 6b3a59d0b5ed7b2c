// Stand-in for <ceres/ceres.h> (TEST INFRASTRUCTURE): only the two abstract interfaces the reference's factor and
// parameterisation classes derive from (ceres/sized_cost_function.h, ceres/local_parameterization.h) -- enough to
// compile lidar_factor.cc and pose_local_parameterization.cc UNMODIFIED.  No solver: ceres::Solve is restated in
// oracle/msfl_oracle.c (msflo_lm_solve) and stays unpinned.
#ifndef MSFL_CERES_STANDIN_H
#define MSFL_CERES_STANDIN_H
namespace ceres {
class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
};
template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {};
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
}  // namespace ceres
#endif
