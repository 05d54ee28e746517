// oracle/shim_fixed: the part of the Ceres API the reference's registration touches.  Test infrastructure only.
// The interfaces (CostFunction, LocalParameterization, LossFunction, Problem, Solver::Options) are declarations; Ceres
// itself is not in this image, so ceres::Solve is THIRD-PARTY ARITHMETIC answered by the oracle's restatement of its
// trust-region Levenberg-Marquardt loop (lmsf_oracle_lm_solve_cb in oracle/lmsf_oracle.cpp: Jacobi scaling, LM
// diagonal, step acceptance, radius update, tolerances, Huber loss + corrector) run over the problem's OWN residual
// blocks: the reference's cost functions' Evaluate and its parameterization's Plus / ComputeJacobian are what it calls.
#pragma once
#include <memory>
#include <set>
#include <vector>

extern "C" {
typedef void (*lmsf_oracle_residual_cb)(void* user, int i, const double x[7], int want_jac, double* r, double J6[6]);
typedef void (*lmsf_oracle_plus_cb)(void* user, const double x[7], const double d[6], double out[7]);
int lmsf_oracle_lm_solve_cb(int n_res, lmsf_oracle_residual_cb residual, lmsf_oracle_plus_cb plus, void* user,
                            double huber_a, int max_iters, double x[7], int* steps, int* accepted, double* cost);
}

namespace ceres {

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};
template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual double huber_a() const = 0;
};
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a) {}
  double huber_a() const override { return a_; }

 private:
  double a_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY };

class Problem {  // one 7-parameter block with a 6-dimensional local parameterization, scalar residual blocks
 public:
  struct Options {};
  Problem() {}
  explicit Problem(const Options&) {}
  ~Problem() {
    for (LossFunction* l : losses_) delete l;  // Ceres takes ownership of cost functions, losses, parameterizations
  }
  void AddParameterBlock(double* values, int size, LocalParameterization* p) {
    if (size != 7) __builtin_trap();
    x_ = values;
    prm_.reset(p);
  }
  void AddResidualBlock(CostFunction* c, LossFunction* l, double* x) {
    if (x != x_) __builtin_trap();
    blocks_.emplace_back(c);
    if (l) losses_.insert(l);
    loss_ = l;
  }
  double* x_ = nullptr;
  std::unique_ptr<LocalParameterization> prm_;
  std::vector<std::unique_ptr<CostFunction>> blocks_;
  std::set<LossFunction*> losses_;
  LossFunction* loss_ = nullptr;
};

struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    bool check_gradients = false;
    double gradient_check_relative_precision = 1e-8;
  };
  struct Summary {
    int num_steps = 0, num_accepted = 0;
    double final_cost = 0;
  };
};

inline void shim_residual(void* user, int i, const double x[7], int want_jac, double* r, double J6[6]) {
  Problem* p = static_cast<Problem*>(user);
  double J7[7], P[42];
  double* jac[1] = {J7};
  const double* par[1] = {x};
  p->blocks_[i]->Evaluate(par, r, want_jac ? jac : nullptr);
  if (!want_jac) return;
  p->prm_->ComputeJacobian(x, P);  // 7 x 6 row-major: global -> local
  for (int j = 0; j < 6; ++j) {
    double s = 0;
    for (int k = 0; k < 7; ++k) s += J7[k] * P[k * 6 + j];
    J6[j] = s;
  }
}
inline void shim_plus(void* user, const double x[7], const double d[6], double out[7]) {
  static_cast<Problem*>(user)->prm_->Plus(x, d, out);
}

inline void Solve(const Solver::Options& o, Problem* p, Solver::Summary* s) {
  if (o.linear_solver_type != DENSE_QR || !p->prm_ || !p->loss_ || p->losses_.size() != 1) __builtin_trap();
  lmsf_oracle_lm_solve_cb((int)p->blocks_.size(), shim_residual, shim_plus, p, p->loss_->huber_a(), o.max_num_iterations,
                          p->x_, &s->num_steps, &s->num_accepted, &s->final_cost);
}

}  // namespace ceres
