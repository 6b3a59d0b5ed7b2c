// Stand-in for <ceres/ceres.h> (TEST INFRASTRUCTURE): the slice of the Ceres API the reference's scan matchers use, so
// that odometry_scan_matcher.cc / mapping_scan_matcher.cc / scan_matcher.cc / lidar_factor.cc /
// pose_local_parameterization.cc compile UNMODIFIED.  Problem only records what the reference hands it; ceres::Solve
// (defined in oracle/ref_shim.cc) evaluates the reference's OWN CostFunction / LossFunction / LocalParameterization
// objects and runs the trust-region loop of the oracle (msflo_lm_solve_cb: the restatement of Ceres'
// TrustRegionMinimizer + LevenbergMarquardtStrategy).  So the solver loop stays "restated, not pinned"; everything it is
// fed -- which residual blocks, which loss, which parameter blocks are constant, the iteration cap -- is the reference's.
#ifndef MSFL_CERES_STANDIN_H
#define MSFL_CERES_STANDIN_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <set>
#include <string>
#include <vector>

#include "../glog/logging.h"  // the real ceres.h pulls glog in too (scan_undistortion.cc relies on it for CHECK)

namespace ceres {

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
  int num_residuals() const { return num_residuals_; }
  const std::vector<int32_t> &parameter_block_sizes() const { return parameter_block_sizes_; }

 protected:
  void set_num_residuals(int n) { num_residuals_ = n; }
  std::vector<int32_t> *mutable_parameter_block_sizes() { return &parameter_block_sizes_; }

 private:
  int num_residuals_ = 0;
  std::vector<int32_t> parameter_block_sizes_;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    *mutable_parameter_block_sizes() = std::vector<int32_t>{Ns...};
  }
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

// ceres/loss_function.cc HuberLoss::Evaluate
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a_ * r - b_;
      rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s;
      rho[1] = 1.0;
      rho[2] = 0.0;
    }
  }

 private:
  const double a_, b_;
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double *x, const double *delta, double *x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double *x, double *jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};

// only ever attached to the speed-bias block, which the reference then holds constant (mapping_scan_matcher.cc:92-94)
class SubsetParameterization : public LocalParameterization {
 public:
  SubsetParameterization(int size, const std::vector<int> &constant_parameters) : size_(size), constant_(constant_parameters) {}
  bool Plus(const double *x, const double *delta, double *x_plus_delta) const override {
    int k = 0;
    for (int i = 0; i < size_; ++i)
      x_plus_delta[i] = std::find(constant_.begin(), constant_.end(), i) != constant_.end() ? x[i] : x[i] + delta[k++];
    return true;
  }
  bool ComputeJacobian(const double *, double *) const override { return false; }
  int GlobalSize() const override { return size_; }
  int LocalSize() const override { return size_ - (int)constant_.size(); }

 private:
  int size_;
  std::vector<int> constant_;
};

struct ResidualBlock {
  CostFunction *cost;
  LossFunction *loss;
  std::vector<double *> parameters;
};
typedef ResidualBlock *ResidualBlockId;
struct CRSMatrix {};

class Problem {
 public:
  struct Options {};
  struct EvaluateOptions {
    bool apply_loss_function = true;
  };
  struct ParameterBlock {
    double *values;
    int size;
    LocalParameterization *parameterization;
    bool constant;
  };
  Problem() {}
  explicit Problem(const Options &) {}
  Problem(const Problem &) = delete;
  ~Problem() {  // Ceres' default ownership: the problem deletes cost functions, loss functions and parameterizations once
    std::set<CostFunction *> c;
    std::set<LossFunction *> l;
    std::set<LocalParameterization *> p;
    for (ResidualBlock *b : blocks_) {
      c.insert(b->cost);
      if (b->loss) l.insert(b->loss);
      delete b;
    }
    for (ParameterBlock &b : parameters_)
      if (b.parameterization) p.insert(b.parameterization);
    for (auto *x : c) delete x;
    for (auto *x : l) delete x;
    for (auto *x : p) delete x;
  }
  void AddParameterBlock(double *values, int size, LocalParameterization *parameterization = nullptr) {
    ParameterBlock *b = find(values);
    if (b) {
      if (parameterization) b->parameterization = parameterization;
      return;
    }
    parameters_.push_back(ParameterBlock{values, size, parameterization, false});
  }
  template <typename... Ts>
  ResidualBlockId AddResidualBlock(CostFunction *cost, LossFunction *loss, double *x0, Ts *... xs) {
    ResidualBlock *b = new ResidualBlock{cost, loss, std::vector<double *>{x0, xs...}};
    for (size_t i = 0; i < b->parameters.size(); ++i)
      if (!find(b->parameters[i])) parameters_.push_back(ParameterBlock{b->parameters[i], cost->parameter_block_sizes()[i], nullptr, false});
    blocks_.push_back(b);
    return b;
  }
  void SetParameterBlockConstant(double *values) {
    ParameterBlock *b = find(values);
    if (b) b->constant = true;
  }
  void GetResidualBlocks(std::vector<ResidualBlockId> *out) const { *out = blocks_; }
  void RemoveResidualBlock(ResidualBlockId id) { blocks_.erase(std::remove(blocks_.begin(), blocks_.end(), id), blocks_.end()); }
  bool Evaluate(const EvaluateOptions &, double *, std::vector<double> *, std::vector<double> *, CRSMatrix *) { return false; }

  const std::vector<ResidualBlock *> &residual_blocks() const { return blocks_; }
  ParameterBlock *find(double *values) {
    for (ParameterBlock &b : parameters_)
      if (b.values == values) return &b;
    return nullptr;
  }

 private:
  std::vector<ResidualBlock *> blocks_;
  std::vector<ParameterBlock> parameters_;
};

struct Solver {
  struct Options {
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
  };
  struct Summary {
    std::string message;
    std::string BriefReport() const { return message; }
  };
};

void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *summary);  // oracle/ref_shim.cc

}  // namespace ceres
#endif
