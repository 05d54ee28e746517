// dmath.cuh — small fixed-size fp64 algebra for the registration kernels and the
// host-side tracker (poses).  Everything here is IEEE +,-,*,/,sqrt only and the
// library is compiled with -fmad=false, so a 3x3 eigen decomposition or a 5x3
// least-squares fit gives the same bits on the device as the same sequence of
// operations does on any IEEE host.  The reference obtains these from Eigen
// (SelfAdjointEigenSolver EdgeFeatureMatch.hpp:63, ColPivHouseholderQR
// surfFeatureMatch.hpp:52 / edgeSurfFeatureRegistration.hpp:272, Quaternion,
// AngleAxis), which is not part of its tree.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define HD __host__ __device__ __forceinline__

namespace lm {

struct d3 {
  double x, y, z;
};
HD d3 mk3(double x, double y, double z) {
  d3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
HD d3 add3(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
HD d3 sub3(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
HD d3 scl3(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
HD double dot3(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HD d3 crs3(d3 a, d3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
HD double nrm3(d3 a) { return sqrt(dot3(a, a)); }

struct quat {
  double x, y, z, w;
};
// rotate v by q the way Eigen's operator* does: v + w*t + u x t with t = 2 u x v
HD d3 qrot(const quat& q, d3 v) {
  d3 u = mk3(q.x, q.y, q.z);
  d3 t = crs3(u, v);
  t = add3(t, t);
  d3 c = crs3(u, t);
  return mk3(v.x + q.w * t.x + c.x, v.y + q.w * t.y + c.y, v.z + q.w * t.z + c.z);
}
HD quat qmul(const quat& a, const quat& b) {
  quat r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}
// row-major rotation matrix of q (Eigen toRotationMatrix term order)
HD void quat_to_mat(const quat& q, double* R) {
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
HD quat mat_to_quat(const double* R) {
  quat q;
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0);
    q.w = 0.5 * s;
    s = 0.5 / s;
    q.x = (R[7] - R[5]) * s;
    q.y = (R[2] - R[6]) * s;
    q.z = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    double v[3];
    v[i] = 0.5 * s;
    s = 0.5 / s;
    q.w = (R[k * 3 + j] - R[j * 3 + k]) * s;
    v[j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
    v[k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
    q.x = v[0];
    q.y = v[1];
    q.z = v[2];
  }
  return q;
}

// rigid transform as the tracker keeps it (Eigen::Isometry3d): row-major R, t
struct rigid {
  double R[9];
  double t[3];
};
HD rigid rigid_identity() {
  rigid a;
  for (int i = 0; i < 9; ++i) a.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  a.t[0] = a.t[1] = a.t[2] = 0.0;
  return a;
}
HD rigid rigid_mul(const rigid& a, const rigid& b) {
  rigid c;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      c.R[i * 3 + j] = a.R[i * 3 + 0] * b.R[0 * 3 + j] + a.R[i * 3 + 1] * b.R[1 * 3 + j] + a.R[i * 3 + 2] * b.R[2 * 3 + j];
    c.t[i] = a.R[i * 3 + 0] * b.t[0] + a.R[i * 3 + 1] * b.t[1] + a.R[i * 3 + 2] * b.t[2] + a.t[i];
  }
  return c;
}
HD rigid rigid_inv(const rigid& a) {
  rigid c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) c.t[i] = -(c.R[i * 3 + 0] * a.t[0] + c.R[i * 3 + 1] * a.t[1] + c.R[i * 3 + 2] * a.t[2]);
  return c;
}
HD rigid rigid_from_pose(const double* p) {
  rigid a;
  quat q;
  q.x = p[0];
  q.y = p[1];
  q.z = p[2];
  q.w = p[3];
  quat_to_mat(q, a.R);
  a.t[0] = p[4];
  a.t[1] = p[5];
  a.t[2] = p[6];
  return a;
}
HD void rigid_to_pose(const rigid& a, double* p) {
  quat q = mat_to_quat(a.R);
  p[0] = q.x;
  p[1] = q.y;
  p[2] = q.z;
  p[3] = q.w;
  p[4] = a.t[0];
  p[5] = a.t[1];
  p[6] = a.t[2];
}

// Cyclic Jacobi eigen-decomposition of a symmetric NxN matrix (row-major).
// w ascending; column j of V (V[i*N+j]) is the eigenvector of w[j].
template <int N>
HD void jacobi_eig_generic(const double* Ain, double* w, double* V) {
  double a[N * N], d[N], b[N], z[N];
#pragma unroll
  for (int i = 0; i < N * N; ++i) {
    a[i] = Ain[i];
    V[i] = ((i / N) == (i % N)) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    b[i] = d[i] = a[i * N + i];
    z[i] = 0.0;
  }
  for (int sweep = 1; sweep <= 60; ++sweep) {
    double sm = 0.0;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) sm += fabs(a[p * N + q]);
    if (sm == 0.0) break;
    double tresh = (sweep < 4) ? 0.2 * sm / (N * N) : 0.0;
    for (int p = 0; p < N - 1; ++p) {
      for (int q = p + 1; q < N; ++q) {
        double apq = a[p * N + q];
        double g = 100.0 * fabs(apq);
        if (sweep > 4 && fabs(d[p]) + g == fabs(d[p]) && fabs(d[q]) + g == fabs(d[q])) {
          a[p * N + q] = 0.0;
        } else if (fabs(apq) > tresh) {
          double h = d[q] - d[p], t;
          if (fabs(h) + g == fabs(h)) {
            t = apq / h;
          } else {
            double theta = 0.5 * h / apq;
            t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          double c = 1.0 / sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
          h = t * apq;
          z[p] -= h;
          z[q] += h;
          d[p] -= h;
          d[q] += h;
          a[p * N + q] = 0.0;
          for (int j = 0; j < p; ++j) {
            double gg = a[j * N + p], hh = a[j * N + q];
            a[j * N + p] = gg - s * (hh + gg * tau);
            a[j * N + q] = hh + s * (gg - hh * tau);
          }
          for (int j = p + 1; j < q; ++j) {
            double gg = a[p * N + j], hh = a[j * N + q];
            a[p * N + j] = gg - s * (hh + gg * tau);
            a[j * N + q] = hh + s * (gg - hh * tau);
          }
          for (int j = q + 1; j < N; ++j) {
            double gg = a[p * N + j], hh = a[q * N + j];
            a[p * N + j] = gg - s * (hh + gg * tau);
            a[q * N + j] = hh + s * (gg - hh * tau);
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

// The 3x3 case written out on scalars: the same operations in the same order as jacobi_eig_generic<3>
// (bit-identical results, checked by test_dmath), but every value lives in a register — the generic
// version indexes small local arrays, which put ~40 us of dependent local-memory latency into every
// thread that fits one edge (EdgeFeatureMatch.hpp:63).
#define LM_JROT3(APQ, DP, DQ, ZP, ZQ, G1, H1, VP0, VQ0, VP1, VQ1, VP2, VQ2)                         \
  {                                                                                                \
    double apq = APQ;                                                                              \
    double g = 100.0 * fabs(apq);                                                                  \
    if (sweep > 4 && fabs(DP) + g == fabs(DP) && fabs(DQ) + g == fabs(DQ)) {                       \
      APQ = 0.0;                                                                                   \
    } else if (fabs(apq) > tresh) {                                                                \
      double h = DQ - DP, t;                                                                       \
      if (fabs(h) + g == fabs(h)) {                                                                \
        t = apq / h;                                                                               \
      } else {                                                                                     \
        double theta = 0.5 * h / apq;                                                              \
        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));                                       \
        if (theta < 0.0) t = -t;                                                                   \
      }                                                                                            \
      double c = 1.0 / sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);                          \
      h = t * apq;                                                                                 \
      ZP -= h;                                                                                     \
      ZQ += h;                                                                                     \
      DP -= h;                                                                                     \
      DQ += h;                                                                                     \
      APQ = 0.0;                                                                                   \
      {                                                                                            \
        double gg = G1, hh = H1;                                                                   \
        G1 = gg - s * (hh + gg * tau);                                                             \
        H1 = hh + s * (gg - hh * tau);                                                             \
      }                                                                                            \
      {                                                                                            \
        double gg = VP0, hh = VQ0;                                                                 \
        VP0 = gg - s * (hh + gg * tau);                                                            \
        VQ0 = hh + s * (gg - hh * tau);                                                            \
      }                                                                                            \
      {                                                                                            \
        double gg = VP1, hh = VQ1;                                                                 \
        VP1 = gg - s * (hh + gg * tau);                                                            \
        VQ1 = hh + s * (gg - hh * tau);                                                            \
      }                                                                                            \
      {                                                                                            \
        double gg = VP2, hh = VQ2;                                                                 \
        VP2 = gg - s * (hh + gg * tau);                                                            \
        VQ2 = hh + s * (gg - hh * tau);                                                            \
      }                                                                                            \
    }                                                                                              \
  }

HD void jacobi_eig3(const double* Ain, double* w, double* V) {
  double a01 = Ain[1], a02 = Ain[2], a12 = Ain[5];
  double d0 = Ain[0], d1 = Ain[4], d2 = Ain[8];
  double b0 = d0, b1 = d1, b2 = d2, z0 = 0.0, z1 = 0.0, z2 = 0.0;
  double v00 = 1.0, v01 = 0.0, v02 = 0.0, v10 = 0.0, v11 = 1.0, v12 = 0.0, v20 = 0.0, v21 = 0.0, v22 = 1.0;
  for (int sweep = 1; sweep <= 60; ++sweep) {
    double sm = 0.0;
    sm += fabs(a01);
    sm += fabs(a02);
    sm += fabs(a12);
    if (sm == 0.0) break;
    double tresh = (sweep < 4) ? 0.2 * sm / 9 : 0.0;
    // (p,q) = (0,1): the j > q term touches a[0][2], a[1][2]
    LM_JROT3(a01, d0, d1, z0, z1, a02, a12, v00, v01, v10, v11, v20, v21)
    // (0,2): the p < j < q term touches a[0][1], a[1][2]
    LM_JROT3(a02, d0, d2, z0, z2, a01, a12, v00, v02, v10, v12, v20, v22)
    // (1,2): the j < p term touches a[0][1], a[0][2]
    LM_JROT3(a12, d1, d2, z1, z2, a01, a02, v01, v02, v11, v12, v21, v22)
    b0 += z0;
    d0 = b0;
    z0 = 0.0;
    b1 += z1;
    d1 = b1;
    z1 = 0.0;
    b2 += z2;
    d2 = b2;
    z2 = 0.0;
  }
  // ascending selection sort, columns follow (same comparisons as the generic version)
#define LM_JSWAP(DA, DB, VA0, VB0, VA1, VB1, VA2, VB2) \
  {                                                    \
    double t_ = DA;                                    \
    DA = DB;                                           \
    DB = t_;                                           \
    t_ = VA0;                                          \
    VA0 = VB0;                                         \
    VB0 = t_;                                          \
    t_ = VA1;                                          \
    VA1 = VB1;                                         \
    VB1 = t_;                                          \
    t_ = VA2;                                          \
    VA2 = VB2;                                         \
    VB2 = t_;                                          \
  }
  {
    int k = 0;
    double dk = d0;
    if (d1 < dk) {
      k = 1;
      dk = d1;
    }
    if (d2 < dk) {
      k = 2;
      dk = d2;
    }
    if (k == 1) LM_JSWAP(d0, d1, v00, v01, v10, v11, v20, v21)
    if (k == 2) LM_JSWAP(d0, d2, v00, v02, v10, v12, v20, v22)
  }
  if (d2 < d1) LM_JSWAP(d1, d2, v01, v02, v11, v12, v21, v22)
#undef LM_JSWAP
  w[0] = d0;
  w[1] = d1;
  w[2] = d2;
  V[0] = v00;
  V[1] = v01;
  V[2] = v02;
  V[3] = v10;
  V[4] = v11;
  V[5] = v12;
  V[6] = v20;
  V[7] = v21;
  V[8] = v22;
}
#undef LM_JROT3

template <int N>
HD void jacobi_eig(const double* Ain, double* w, double* V) {
  jacobi_eig_generic<N>(Ain, w, V);
}
template <>
HD void jacobi_eig<3>(const double* Ain, double* w, double* V) {
  jacobi_eig3(Ain, w, V);
}

// Least squares min |A x - b| for a row-major MxN A by Householder QR with
// column pivoting (largest remaining column norm first); a column whose pivot
// falls under eps*M*|first pivot| ends the factorisation, its x stays 0.
template <int M, int N>
HD void cpqr_solve_generic(const double* Ain, const double* bin, double* x) {
  double A[M * N], b[M], v[M], y[N];
  int perm[N];
  for (int i = 0; i < M * N; ++i) A[i] = Ain[i];
  for (int i = 0; i < M; ++i) b[i] = bin[i];
  for (int j = 0; j < N; ++j) {
    perm[j] = j;
    y[j] = 0.0;
  }
  int rank = 0;
  double first = 0.0;
  const int K = (M < N) ? M : N;
  for (int k = 0; k < K; ++k) {
    int piv = k;
    double best = -1.0;
    for (int j = k; j < N; ++j) {
      double s = 0.0;
      for (int i = k; i < M; ++i) s += A[i * N + j] * A[i * N + j];
      if (s > best) {
        best = s;
        piv = j;
      }
    }
    if (piv != k) {
      for (int i = 0; i < M; ++i) {
        double t = A[i * N + k];
        A[i * N + k] = A[i * N + piv];
        A[i * N + piv] = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    double nrm = sqrt(best);
    if (k == 0) first = nrm;
    if (!(nrm > 2.220446049250313e-16 * M * first) || nrm == 0.0) break;
    rank = k + 1;
    double alpha = (A[k * N + k] > 0.0) ? -nrm : nrm;
    double v0 = A[k * N + k] - alpha;
    v[k] = 1.0;
    for (int i = k + 1; i < M; ++i) v[i] = A[i * N + k] / v0;
    double beta = -v0 / alpha;
    A[k * N + k] = alpha;
    for (int i = k + 1; i < M; ++i) A[i * N + k] = 0.0;
    for (int j = k + 1; j < N; ++j) {
      double s = 0.0;
      for (int i = k; i < M; ++i) s += v[i] * A[i * N + j];
      s *= beta;
      for (int i = k; i < M; ++i) A[i * N + j] -= s * v[i];
    }
    double s = 0.0;
    for (int i = k; i < M; ++i) s += v[i] * b[i];
    s *= beta;
    for (int i = k; i < M; ++i) b[i] -= s * v[i];
  }
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < rank; ++j) s -= A[k * N + j] * y[j];
    y[k] = s / A[k * N + k];
  }
  for (int j = 0; j < N; ++j) x[perm[j]] = y[j];
}

// Index-static form of the same factorisation (identical operations and order; pivoting moves data with
// conditional swaps) so that, fully unrolled, the 5x3 plane fit of SurfFeatureMatch (surfFeatureMatch.hpp:52)
// lives in registers.  Used for the 5x3 case only; test_dmath checks it bit for bit against the generic form.
template <int M, int N>
HD void cpqr_solve_static(const double* Ain, const double* bin, double* x) {
  // every array index below is a compile-time constant once the loops are unrolled (M*N small), so the
  // whole factorisation lives in registers; pivoting moves data with conditional swaps
  double A[M * N], b[M], v[M], y[N];
  int perm[N];
#pragma unroll
  for (int i = 0; i < M * N; ++i) A[i] = Ain[i];
#pragma unroll
  for (int i = 0; i < M; ++i) b[i] = bin[i];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    perm[j] = j;
    y[j] = 0.0;
  }
  int rank = 0;
  double first = 0.0;
  constexpr int K = (M < N) ? M : N;
  bool live = true;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (live) {
      int piv = k;
      double best = -1.0;
#pragma unroll
      for (int j = k; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = k; i < M; ++i) s += A[i * N + j] * A[i * N + j];
        if (s > best) {
          best = s;
          piv = j;
        }
      }
#pragma unroll
      for (int j = k + 1; j < N; ++j) {
        if (piv == j) {
#pragma unroll
          for (int i = 0; i < M; ++i) {
            double t = A[i * N + k];
            A[i * N + k] = A[i * N + j];
            A[i * N + j] = t;
          }
          int t = perm[k];
          perm[k] = perm[j];
          perm[j] = t;
        }
      }
      double nrm = sqrt(best);
      if (k == 0) first = nrm;
      if (!(nrm > 2.220446049250313e-16 * M * first) || nrm == 0.0) {
        live = false;
      } else {
        rank = k + 1;
        double alpha = (A[k * N + k] > 0.0) ? -nrm : nrm;
        double v0 = A[k * N + k] - alpha;
        v[k] = 1.0;
#pragma unroll
        for (int i = k + 1; i < M; ++i) v[i] = A[i * N + k] / v0;
        double beta = -v0 / alpha;
        A[k * N + k] = alpha;
#pragma unroll
        for (int i = k + 1; i < M; ++i) A[i * N + k] = 0.0;
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
          double s = 0.0;
#pragma unroll
          for (int i = k; i < M; ++i) s += v[i] * A[i * N + j];
          s *= beta;
#pragma unroll
          for (int i = k; i < M; ++i) A[i * N + j] -= s * v[i];
        }
        double s = 0.0;
#pragma unroll
        for (int i = k; i < M; ++i) s += v[i] * b[i];
        s *= beta;
#pragma unroll
        for (int i = k; i < M; ++i) b[i] -= s * v[i];
      }
    }
  }
#pragma unroll
  for (int k = K - 1; k >= 0; --k) {
    if (k < rank) {
      double s = b[k];
#pragma unroll
      for (int j = k + 1; j < K; ++j)
        if (j < rank) s -= A[k * N + j] * y[j];
      y[k] = s / A[k * N + k];
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int jj = 0; jj < N; ++jj)
      if (perm[j] == jj) x[jj] = y[j];
  }
}

template <int M, int N>
HD void cpqr_solve(const double* Ain, const double* bin, double* x) {
  cpqr_solve_generic<M, N>(Ain, bin, x);
}
template <>
HD void cpqr_solve<5, 3>(const double* Ain, const double* bin, double* x) {
  cpqr_solve_static<5, 3>(Ain, bin, x);
}

// 6x6 inverse, Gauss-Jordan with partial pivoting
HD bool invert6(const double* Ain, double* out) {
  double a[6 * 12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      a[i * 12 + j] = Ain[i * 6 + j];
      a[i * 12 + 6 + j] = (i == j) ? 1.0 : 0.0;
    }
  bool ok = true;
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r)
      if (fabs(a[r * 12 + c]) > fabs(a[p * 12 + c])) p = r;
    if (p != c)
      for (int j = 0; j < 12; ++j) {
        double t = a[c * 12 + j];
        a[c * 12 + j] = a[p * 12 + j];
        a[p * 12 + j] = t;
      }
    double d = a[c * 12 + c];
    if (d == 0.0) ok = false;
    for (int j = 0; j < 12; ++j) a[c * 12 + j] /= d;
    for (int r = 0; r < 6; ++r)
      if (r != c) {
        double f = a[r * 12 + c];
        if (f != 0.0)
          for (int j = 0; j < 12; ++j) a[r * 12 + j] -= f * a[c * 12 + j];
      }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out[i * 6 + j] = a[i * 12 + 6 + j];
  return ok;
}

// 6x6 SPD solve (Cholesky); false when a pivot is not positive
HD bool spd_solve6(const double* A, const double* b, double* x) {
  double L[36], y[6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = sqrt(s);
      } else {
        L[i * 6 + j] = s / L[j * 6 + j];
      }
    }
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
    y[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s / L[i * 6 + i];
  }
  return true;
}

// se3 -> (quaternion, translation), Math::GetTransformFromSe3 (include/Math.hpp:29-72)
HD void se3_to_qt(const double* d, quat& q, d3& t) {
  d3 om = mk3(d[0], d[1], d[2]), up = mk3(d[3], d[4], d[5]);
  double theta = nrm3(om);
  double half = 0.5 * theta;
  double re = cos(half), im;
  if (theta < 1e-10) {
    double t2 = theta * theta, t4 = t2 * t2;
    im = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    im = sin(half) / theta;
  }
  q.x = im * om.x;
  q.y = im * om.y;
  q.z = im * om.z;
  q.w = re;
  double J[9];
  if (theta < 1e-10) {
    quat_to_mat(q, J);
  } else {
    double O[9] = {0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3 + 0] * O[0 * 3 + j] + O[i * 3 + 1] * O[1 * 3 + j] + O[i * 3 + 2] * O[2 * 3 + j];
    double c1 = (1 - cos(theta)) / (theta * theta);
    double c2 = (theta - sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  t = mk3(J[0] * up.x + J[1] * up.y + J[2] * up.z, J[3] * up.x + J[4] * up.y + J[5] * up.z,
          J[6] * up.x + J[7] * up.y + J[8] * up.z);
}

// PoseSE3Parameterization::Plus (PoseSE3Parameterization.hpp:32-46): left-multiplicative update
HD void se3_plus(const double* x, const double* d, double* out) {
  quat dq;
  d3 dt;
  se3_to_qt(d, dq, dt);
  quat q;
  q.x = x[0];
  q.y = x[1];
  q.z = x[2];
  q.w = x[3];
  quat qp = qmul(dq, q);
  d3 tp = add3(qrot(dq, mk3(x[4], x[5], x[6])), dt);
  out[0] = qp.x;
  out[1] = qp.y;
  out[2] = qp.z;
  out[3] = qp.w;
  out[4] = tp.x;
  out[5] = tp.y;
  out[6] = tp.z;
}

}  // namespace lm
