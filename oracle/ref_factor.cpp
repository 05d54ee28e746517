// oracle/ref_factor.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN, unmodified Ceres cost
// functions and SE3 parameterization (Algorithm/PointClouds/registration/ceres_factor/edge_factor.hpp:33-61,
// surf_factor.hpp:32-56, Algorithm/Ceres/Parameterization/PoseSE3Parameterization.hpp:32-63, Math.hpp:19-72), compiled
// where they lie under /root/reference against oracle/shim_fixed/ (Eigen's fixed-size expressions coefficient by
// coefficient; quaternion product / rotation answered by oracle_math.h; ceres/ceres.h = the two interfaces).
// Pins row a5.4: residuals, Jacobian expressions and conventions (left perturbation, a / b order, signs), Plus.
#include <cmath>
#include <string>
#include <vector>

using namespace std;  // as in the node's translation unit (src/apps/include/utility.hpp:51)

#include "Algorithm/PointClouds/registration/ceres_factor/edge_factor.hpp"
#include "Algorithm/PointClouds/registration/ceres_factor/surf_factor.hpp"
#include "Algorithm/Ceres/Parameterization/PoseSE3Parameterization.hpp"

// kind 0: se3PointEdgeFactor(p, a = geom[0:3], b = geom[3:6]); kind 1: se3PointSurfFactor(p, n = geom[0:3], D = geom[3]).
// J6 = Evaluate's 1x7 Jacobian times the parameterization's 7x6 ComputeJacobian (what Ceres forms for the solver).
extern "C" int ref_factor_eval(int kind, const double x[7], const double p[3], const double geom[7], double* r,
                               double J6[6]) {
  Eigen::Vector3d pt(p[0], p[1], p[2]);
  double J7[7], P[42];
  double* jac[1] = {J7};
  const double* par[1] = {x};
  if (kind == 0) {
    Algorithm::se3PointEdgeFactor f(pt, Eigen::Vector3d(geom[0], geom[1], geom[2]), Eigen::Vector3d(geom[3], geom[4], geom[5]));
    f.Evaluate(par, r, jac);
  } else {
    Algorithm::se3PointSurfFactor f(pt, Eigen::Vector3d(geom[0], geom[1], geom[2]), geom[3]);
    f.Evaluate(par, r, jac);
  }
  Algorithm::PoseSE3Parameterization prm;
  prm.ComputeJacobian(x, P);
  for (int j = 0; j < 6; ++j) {
    double s = 0;
    for (int k = 0; k < 7; ++k) s += J7[k] * P[k * 6 + j];
    J6[j] = s;
  }
  return prm.GlobalSize() == 7 && prm.LocalSize() == 6 ? 0 : -1;
}

extern "C" int ref_se3_plus(const double x[7], const double d[6], double out[7]) {
  Algorithm::PoseSE3Parameterization prm;
  prm.Plus(x, d, out);
  return 0;
}
