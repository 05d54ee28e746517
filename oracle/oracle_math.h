// oracle_math.h — fixed-size fp64 linear algebra and SE3 helpers of the CPU ORACLE.
//
// TEST INFRASTRUCTURE ONLY (see lmsf_oracle.h).  The reference gets this
// arithmetic from Eigen (absent here, unpinned): SelfAdjointEigenSolver,
// ColPivHouseholderQR, Quaternion, AngleAxis.  The restatements below use only
// IEEE +,-,*,/,sqrt so that a build with -ffp-contract=off is reproducible
// bit for bit on any IEEE machine.
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

// ---------------------------------------------------------------- vectors
struct V3 {
  double x, y, z;
};
static inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

// ---------------------------------------------------------------- quaternion (x,y,z,w)
struct Quat {
  double x, y, z, w;
};
// Eigen::Quaternion::_transformVector: v + w*(2 u x v) + u x (2 u x v)
static inline V3 qrot(const Quat& q, V3 v) {
  V3 u{q.x, q.y, q.z};
  V3 uv = cross(u, v);
  uv = uv + uv;
  V3 c = cross(u, uv);
  return {v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z};
}
// Eigen quaternion product a*b
static inline Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
// Eigen::Quaternion::toRotationMatrix, row-major R[9]
static inline void q2R(const Quat& q, double R[9]) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}
// Eigen quaternion-from-matrix (internal::quaternionbase_assign_impl<Matrix3>)
static inline Quat R2q(const double R[9]) {
  Quat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t;
    q.y = (R[2] - R[6]) * t;
    q.z = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    q.x = v[0];
    q.y = v[1];
    q.z = v[2];
  }
  return q;
}

// ---------------------------------------------------------------- rigid transform (Isometry3d)
struct Iso {
  double R[9];
  double t[3];
};
static inline Iso iso_identity() {
  Iso a;
  std::memset(&a, 0, sizeof a);
  a.R[0] = a.R[4] = a.R[8] = 1.0;
  return a;
}
static inline Iso iso_mul(const Iso& a, const Iso& b) {
  Iso c;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      c.R[i * 3 + j] = a.R[i * 3 + 0] * b.R[0 * 3 + j] + a.R[i * 3 + 1] * b.R[1 * 3 + j] + a.R[i * 3 + 2] * b.R[2 * 3 + j];
    c.t[i] = a.R[i * 3 + 0] * b.t[0] + a.R[i * 3 + 1] * b.t[1] + a.R[i * 3 + 2] * b.t[2] + a.t[i];
  }
  return c;
}
// Isometry inverse: R^T, -R^T t
static inline Iso iso_inv(const Iso& a) {
  Iso c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) c.t[i] = -(c.R[i * 3 + 0] * a.t[0] + c.R[i * 3 + 1] * a.t[1] + c.R[i * 3 + 2] * a.t[2]);
  return c;
}
static inline Iso iso_from_pose7(const double p[7]) {
  Iso a;
  Quat q{p[0], p[1], p[2], p[3]};
  q2R(q, a.R);
  a.t[0] = p[4];
  a.t[1] = p[5];
  a.t[2] = p[6];
  return a;
}
static inline void iso_to_pose7(const Iso& a, double p[7]) {
  Quat q = R2q(a.R);
  p[0] = q.x;
  p[1] = q.y;
  p[2] = q.z;
  p[3] = q.w;
  p[4] = a.t[0];
  p[5] = a.t[1];
  p[6] = a.t[2];
}

// ---------------------------------------------------------------- symmetric eigen (cyclic Jacobi)
// Stands in for Eigen::SelfAdjointEigenSolver (EdgeFeatureMatch.hpp:63,
// edgeSurfFeatureRegistration.hpp:282): eigenvalues ascending, eigenvectors in
// the COLUMNS of V (row-major V[i*N+j] = component i of vector j).
template <int N>
static inline void symeig(const double* Ain, double* w, double* V) {
  double a[N][N], d[N], b[N], z[N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      a[i][j] = Ain[i * N + j];
      V[i * N + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int i = 0; i < N; ++i) {
    b[i] = d[i] = a[i][i];
    z[i] = 0.0;
  }
  for (int sweep = 1; sweep <= 60; ++sweep) {
    double sm = 0.0;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) sm += std::fabs(a[p][q]);
    if (sm == 0.0) break;
    double tresh = (sweep < 4) ? 0.2 * sm / (N * N) : 0.0;
    for (int p = 0; p < N - 1; ++p) {
      for (int q = p + 1; q < N; ++q) {
        double g = 100.0 * std::fabs(a[p][q]);
        if (sweep > 4 && std::fabs(d[p]) + g == std::fabs(d[p]) && std::fabs(d[q]) + g == std::fabs(d[q])) {
          a[p][q] = 0.0;
        } else if (std::fabs(a[p][q]) > tresh) {
          double h = d[q] - d[p], t;
          if (std::fabs(h) + g == std::fabs(h)) {
            t = a[p][q] / h;
          } else {
            double theta = 0.5 * h / a[p][q];
            t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
          h = t * a[p][q];
          z[p] -= h;
          z[q] += h;
          d[p] -= h;
          d[q] += h;
          a[p][q] = 0.0;
          for (int j = 0; j < p; ++j) {
            double gg = a[j][p], hh = a[j][q];
            a[j][p] = gg - s * (hh + gg * tau);
            a[j][q] = hh + s * (gg - hh * tau);
          }
          for (int j = p + 1; j < q; ++j) {
            double gg = a[p][j], hh = a[j][q];
            a[p][j] = gg - s * (hh + gg * tau);
            a[j][q] = hh + s * (gg - hh * tau);
          }
          for (int j = q + 1; j < N; ++j) {
            double gg = a[p][j], hh = a[q][j];
            a[p][j] = gg - s * (hh + gg * tau);
            a[q][j] = hh + s * (gg - hh * tau);
          }
          for (int j = 0; j < N; ++j) {
            double gg = V[j * N + p], hh = V[j * N + q];
            V[j * N + p] = gg - s * (hh + gg * tau);
            V[j * N + q] = hh + s * (gg - hh * tau);
          }
        }
      }
    }
    for (int i = 0; i < N; ++i) {
      b[i] += z[i];
      d[i] = b[i];
      z[i] = 0.0;
    }
  }
  // ascending selection sort, columns follow
  for (int i = 0; i < N - 1; ++i) {
    int k = i;
    for (int j = i + 1; j < N; ++j)
      if (d[j] < d[k]) k = j;
    if (k != i) {
      double tmp = d[i];
      d[i] = d[k];
      d[k] = tmp;
      for (int r = 0; r < N; ++r) {
        tmp = V[r * N + i];
        V[r * N + i] = V[r * N + k];
        V[r * N + k] = tmp;
      }
    }
  }
  for (int i = 0; i < N; ++i) w[i] = d[i];
}

// ---------------------------------------------------------------- column-pivoted Householder QR least squares
// Stands in for Eigen::ColPivHouseholderQR::solve (surfFeatureMatch.hpp:52 with
// M=5,N=3; edgeSurfFeatureRegistration.hpp:272 with M=N=6).  A row-major MxN.
// Rank-deficient columns (|R_kk| <= eps*M*|R_00|) get a zero coefficient.
template <int M, int N>
static inline void qr_solve(const double* Ain, const double* bin, double* x) {
  double A[M][N], b[M];
  int perm[N];
  for (int i = 0; i < M; ++i) {
    b[i] = bin[i];
    for (int j = 0; j < N; ++j) A[i][j] = Ain[i * N + j];
  }
  for (int j = 0; j < N; ++j) perm[j] = j;
  int rank = 0;
  double r00 = 0.0;
  const int K = (M < N) ? M : N;
  for (int k = 0; k < K; ++k) {
    int piv = k;
    double best = -1.0;
    for (int j = k; j < N; ++j) {
      double s = 0.0;
      for (int i = k; i < M; ++i) s += A[i][j] * A[i][j];
      if (s > best) {
        best = s;
        piv = j;
      }
    }
    if (piv != k) {
      for (int i = 0; i < M; ++i) {
        double t = A[i][k];
        A[i][k] = A[i][piv];
        A[i][piv] = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    double nrm = std::sqrt(best);
    if (k == 0) r00 = nrm;
    if (!(nrm > 2.220446049250313e-16 * M * r00) || nrm == 0.0) break;
    rank = k + 1;
    double alpha = (A[k][k] > 0.0) ? -nrm : nrm;
    // v = x - alpha e1, normalised so v[0] = 1
    double v0 = A[k][k] - alpha;
    double v[M];
    v[k] = 1.0;
    for (int i = k + 1; i < M; ++i) v[i] = A[i][k] / v0;
    double beta = -v0 / alpha;  // = 2 / (v^T v)
    A[k][k] = alpha;
    for (int i = k + 1; i < M; ++i) A[i][k] = 0.0;
    for (int j = k + 1; j < N; ++j) {
      double s = 0.0;
      for (int i = k; i < M; ++i) s += v[i] * A[i][j];
      s *= beta;
      for (int i = k; i < M; ++i) A[i][j] -= s * v[i];
    }
    double s = 0.0;
    for (int i = k; i < M; ++i) s += v[i] * b[i];
    s *= beta;
    for (int i = k; i < M; ++i) b[i] -= s * v[i];
  }
  double y[N];
  for (int j = 0; j < N; ++j) y[j] = 0.0;
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < rank; ++j) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  for (int j = 0; j < N; ++j) x[perm[j]] = y[j];
}

// 6x6 general inverse, Gauss-Jordan with partial pivoting (Eigen: PartialPivLU
// behind MatrixXd::inverse(), edgeSurfFeatureRegistration.hpp:303).
static inline bool inv6(const double* Ain, double* out) {
  const int N = 6;
  double a[N][2 * N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      a[i][j] = Ain[i * N + j];
      a[i][N + j] = (i == j) ? 1.0 : 0.0;
    }
  bool ok = true;
  for (int c = 0; c < N; ++c) {
    int p = c;
    for (int r = c + 1; r < N; ++r)
      if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 2 * N; ++j) {
        double t = a[c][j];
        a[c][j] = a[p][j];
        a[p][j] = t;
      }
    double d = a[c][c];
    if (d == 0.0) ok = false;
    for (int j = 0; j < 2 * N; ++j) a[c][j] /= d;
    for (int r = 0; r < N; ++r)
      if (r != c) {
        double f = a[r][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * N; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) out[i * N + j] = a[i][N + j];
  return ok;
}

// 6x6 SPD solve by Cholesky (used for the LM normal equations; Ceres DENSE_QR
// solves the same regularised least-squares problem by QR of [J; D]).
static inline bool chol_solve6(const double* A, const double* b, double* x) {
  const int N = 6;
  double L[N][N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * N + j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = std::sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double y[N];
  for (int i = 0; i < N; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < N; ++k) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
  return true;
}

// Math::GetTransformFromSe3 (include/Math.hpp:29-72): se3 = [omega; upsilon].
static inline void se3_exp(const double d[6], Quat& q, V3& t) {
  V3 om{d[0], d[1], d[2]}, up{d[3], d[4], d[5]};
  double theta = norm(om);
  double half = 0.5 * theta;
  double real = std::cos(half), imag;
  if (theta < 1e-10) {
    double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = std::sin(half) / theta;
  }
  q = Quat{imag * om.x, imag * om.y, imag * om.z, real};
  double J[9];
  if (theta < 1e-10) {
    q2R(q, J);
  } else {
    double O[9] = {0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3 + 0] * O[0 * 3 + j] + O[i * 3 + 1] * O[1 * 3 + j] + O[i * 3 + 2] * O[2 * 3 + j];
    double c1 = (1 - std::cos(theta)) / (theta * theta);
    double c2 = (theta - std::sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  t = V3{J[0] * up.x + J[1] * up.y + J[2] * up.z, J[3] * up.x + J[4] * up.y + J[5] * up.z,
         J[6] * up.x + J[7] * up.y + J[8] * up.z};
}

}  // namespace orc
