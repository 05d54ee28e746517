// oracle/shim_fixed: the two Ceres interfaces the reference's cost functions and SE3 parameterization derive from
// (declarations only; Ceres itself is not in this image and its solver is restated in the oracle).
#pragma once
namespace ceres {
template <int kNumResiduals, int... Ns>
class SizedCostFunction {
 public:
  virtual ~SizedCostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
}  // namespace ceres
