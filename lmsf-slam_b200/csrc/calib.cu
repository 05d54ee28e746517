// calib.cu — multi-LiDAR extrinsic initialisation, host side of the calibration loop.
//
// Restates Algorithm::HandEyeCalibrationBase (Algorithm/calibration/handeye_calibration_base.hpp:36-246) as it is
// driven by MultiLidarSystem::process() in calibration status 0 (System/ML_System.hpp:243-283): the per-sweep motion
// increments of the primary and of an auxiliary LiDAR tracker are screened (checkScrewMotion :208-243), accumulated,
// stored (at most 300 pairs, the smallest rotation replaced first, AddPose :73-108), and the extrinsic rotation is the
// right singular vector of the stacked (L(q_primary) - R(q_sub)) blocks (CalibExRotation :115-153), the translation the
// least-squares solution of (R_primary - I) t = q_ext * t_sub - t_primary (calibExTranslationNonPlanar :165-190).
// Status 1 (:286-323), the per-sweep refinement, is lmsf_tracker_register_aux on the primary's context.
//
// This is a few hundred flops per sweep on 4x4 / 3x3 matrices: host code, like the reference's — there is no device
// work to take over, the device part of the loop (the trackers and the auxiliary registration) is in api.cu.
// The singular values / vectors of the (4N x 4) matrix are taken from the symmetric eigen-decomposition of Q^T Q
// (cyclic Jacobi), Eigen's JacobiSVD gives the same subspace; the 3-unknown least squares goes through its normal
// equations (full column rank assumed, as the reference's SVD solve does for a well-posed problem).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <queue>
#include <utility>
#include <vector>

#include "../../include/lmsf_b200.h"

namespace {

struct Q4 {
  double x, y, z, w;
};
struct V3 {
  double x, y, z;
};
struct PoseQT {  // Slam3D::Pose (Common/pose.hpp): unit quaternion + translation
  Q4 q{0, 0, 0, 1};
  V3 t{0, 0, 0};
};

Q4 qnormalize(Q4 q) {
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Q4{q.x / n, q.y / n, q.z / n, q.w / n};
}
Q4 qmul(const Q4& a, const Q4& b) {  // Eigen: a * b
  return Q4{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
V3 qrot(const Q4& q, const V3& v) {  // Eigen: q * v = v + 2 w (u x v) + 2 u x (u x v)
  V3 u{q.x, q.y, q.z};
  V3 uv{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
  uv = V3{uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  V3 uuv{u.y * uv.z - u.z * uv.y, u.z * uv.x - u.x * uv.z, u.x * uv.y - u.y * uv.x};
  return V3{v.x + q.w * uv.x + uuv.x, v.y + q.w * uv.y + uuv.y, v.z + q.w * uv.z + uuv.z};
}
// Pose * Pose (pose.hpp poseTransform / operator*): Pose(q1 q2, q1 t2 + t1), the constructor normalises q
PoseQT pmul(const PoseQT& a, const PoseQT& b) {
  PoseQT r;
  r.q = qnormalize(qmul(a.q, b.q));
  V3 rt = qrot(a.q, b.t);
  r.t = V3{rt.x + a.t.x, rt.y + a.t.y, rt.z + a.t.z};
  return r;
}
// Eigen::AngleAxisd(q): angle in [0, pi], axis = vec / |vec| (x axis for a null rotation)
void angle_axis(Q4 q, double& angle, V3& axis) {
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  if (q.w < 0) n = -n;
  if (n != 0.0) {
    angle = 2.0 * atan2(n, fabs(q.w));
    axis = V3{q.x / n, q.y / n, q.z / n};
  } else {
    angle = 0.0;
    axis = V3{1, 0, 0};
  }
}
void q_to_R(const Q4& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz), R[1] = txy - twz, R[2] = txz + twy;
  R[3] = txy + twz, R[4] = 1 - (txx + tzz), R[5] = tyz - twx;
  R[6] = txz - twy, R[7] = tyz + twx, R[8] = 1 - (txx + tyy);
}

// cyclic Jacobi for a symmetric n x n matrix (n <= 4): eigenvalues ascending, eigenvectors in the columns of V
template <int N>
void sym_eig(const double* Ain, double* w, double* V) {
  double A[N * N];
  for (int i = 0; i < N * N; ++i) A[i] = Ain[i];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    if (off < 1e-300) break;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) {
        if (A[p * N + q] == 0.0) continue;
        double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * A[p * N + q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
  int order[N];
  for (int i = 0; i < N; ++i) order[i] = i;
  std::sort(order, order + N, [&](int a, int b) { return A[a * N + a] < A[b * N + b]; });
  double Vs[N * N];
  for (int j = 0; j < N; ++j) {
    w[j] = A[order[j] * N + order[j]];
    for (int i = 0; i < N; ++i) Vs[i * N + j] = V[i * N + order[j]];
  }
  for (int i = 0; i < N * N; ++i) V[i] = Vs[i];
}

typedef std::pair<unsigned short, std::pair<PoseQT, PoseQT>> Entry;
struct RotCmp {  // rotCmp (:24-31): the pair with the LARGEST w (smallest rotation) on top
  bool operator()(const Entry& r, const Entry& l) const { return l.second.first.q.w > r.second.first.q.w; }
};

}  // namespace

struct lmsf_handeye {
  static constexpr double EPSILON_R = 0.05, EPSILON_T = 0.1;
  static constexpr size_t N_POSE = 300;
  double rot_cov_thre = 0.25;
  std::priority_queue<Entry, std::vector<Entry>, RotCmp> priority_pose;
  std::queue<Entry> new_pose_pair;
  std::vector<std::pair<PoseQT, PoseQT>> pose_storage;
  std::vector<double> Q;  // N_POSE * 4 rows x 4
  Q4 ext_q{0, 0, 0, 1};
  V3 ext_t{0, 0, 0};
  double sv[4] = {0, 0, 0, 0};  // singular values, descending like Eigen's
  bool calib_done = false;
  PoseQT acc_pri, acc_sub;
  lmsf_handeye() : Q(N_POSE * 4 * 4, 0.0) { pose_storage.reserve(N_POSE); }

  // checkScrewMotion (:208-243)
  bool check_screw(const PoseQT& pri, const PoseQT& sub) {
    double a_pri, a_sub;
    V3 ax_pri, ax_sub;
    angle_axis(pri.q, a_pri, ax_pri);
    angle_axis(sub.q, a_sub, ax_sub);
    double r_dis = fabs(a_pri - a_sub);
    double t_dis = fabs((pri.t.x * ax_pri.x + pri.t.y * ax_pri.y + pri.t.z * ax_pri.z) -
                        (sub.t.x * ax_sub.x + sub.t.y * ax_sub.y + sub.t.z * ax_sub.z));
    if (r_dis > EPSILON_R || t_dis > EPSILON_T) {
      acc_pri = PoseQT();
      acc_sub = PoseQT();
      return false;
    }
    acc_pri = pmul(acc_pri, pri);
    acc_sub = pmul(acc_sub, sub);
    double a1, a2;
    V3 dummy;
    angle_axis(acc_pri.q, a1, dummy);
    angle_axis(acc_sub.q, a2, dummy);
    return a1 > 0 || a2 > 0;  // (the reference falls off the end otherwise: undefined; taken as "not yet")
  }

  // AddPose (:73-108)
  bool add_pose(const PoseQT& pri, const PoseQT& sub) {
    if (!check_screw(pri, sub)) return false;
    if (pose_storage.size() < N_POSE) {
      Entry e((unsigned short)pose_storage.size(), std::make_pair(acc_pri, acc_sub));
      new_pose_pair.push(e);
      priority_pose.push(e);
      pose_storage.emplace_back(acc_pri, acc_sub);
    } else {
      unsigned short pos = priority_pose.top().first;
      pose_storage[pos] = std::make_pair(acc_pri, acc_sub);
      Entry e(pos, std::make_pair(acc_pri, acc_sub));
      new_pose_pair.push(e);
      priority_pose.pop();
      priority_pose.push(e);
    }
    acc_pri = PoseQT();
    acc_sub = PoseQT();
    return pose_storage.size() >= 3;
  }

  // CalibExRotation (:115-153)
  bool calib_rotation() {
    while (!new_pose_pair.empty()) {
      Entry e = new_pose_pair.front();
      new_pose_pair.pop();
      const Q4& p = e.second.first.q;
      const Q4& s = e.second.second.q;
      // Math::QuanternionLeftProductMatrix(p) - QuanternionRightProductMatrix(s)  (Math.hpp:79-95), [w x y z] order
      const double L[16] = {p.w, -p.x, -p.y, -p.z, p.x, p.w, -p.z, p.y, p.y, p.z, p.w, -p.x, p.z, -p.y, p.x, p.w};
      const double R[16] = {s.w, -s.x, -s.y, -s.z, s.x, s.w, s.z, -s.y, s.y, -s.z, s.w, s.x, s.z, s.y, -s.x, s.w};
      for (int i = 0; i < 16; ++i) Q[(size_t)e.first * 16 + i] = L[i] - R[i];
    }
    double G[16] = {0};  // Q^T Q over all N_POSE blocks (unused blocks are zero, as in the reference's 1200 x 4 matrix)
    for (size_t r = 0; r < N_POSE * 4; ++r)
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) G[i * 4 + j] += Q[r * 4 + i] * Q[r * 4 + j];
    double w[4], V[16];
    sym_eig<4>(G, w, V);
    for (int k = 0; k < 4; ++k) sv[k] = sqrt(std::max(0.0, w[3 - k]));  // descending
    double x[4] = {V[0 * 4 + 0], V[1 * 4 + 0], V[2 * 4 + 0], V[3 * 4 + 0]};  // smallest singular value: [w, x, y, z]
    if (x[0] < 0)
      for (int k = 0; k < 4; ++k) x[k] = -x[k];
    if (sv[2] > rot_cov_thre) {
      ext_q = qnormalize(Q4{x[1], x[2], x[3], x[0]});
      return true;
    }
    return false;
  }

  // calibExTranslationNonPlanar (:165-190)
  bool calib_translation() {
    double AtA[9] = {0}, Atb[3] = {0};
    for (size_t i = 0; i < pose_storage.size(); ++i) {
      double Rm[9];
      q_to_R(pose_storage[i].first.q, Rm);
      Rm[0] -= 1.0, Rm[4] -= 1.0, Rm[8] -= 1.0;
      V3 rt = qrot(ext_q, pose_storage[i].second.t);
      const double b[3] = {rt.x - pose_storage[i].first.t.x, rt.y - pose_storage[i].first.t.y,
                           rt.z - pose_storage[i].first.t.z};
      for (int r = 0; r < 3; ++r)
        for (int a = 0; a < 3; ++a) {
          Atb[a] += Rm[r * 3 + a] * b[r];
          for (int c = 0; c < 3; ++c) AtA[a * 3 + c] += Rm[r * 3 + a] * Rm[r * 3 + c];
        }
    }
    // minimum-norm least squares through the eigen-decomposition of A^T A (what an SVD solve returns)
    double w[3], V[9];
    sym_eig<3>(AtA, w, V);
    double x[3] = {0, 0, 0};
    const double tol = 1e-12 * std::max(w[2], 1e-300);
    for (int k = 0; k < 3; ++k) {
      if (!(w[k] > tol)) continue;
      double proj = (V[0 * 3 + k] * Atb[0] + V[1 * 3 + k] * Atb[1] + V[2 * 3 + k] * Atb[2]) / w[k];
      for (int i = 0; i < 3; ++i) x[i] += V[i * 3 + k] * proj;
    }
    ext_t = V3{x[0], x[1], x[2]};
    calib_done = true;
    return true;
  }
};

extern "C" {

int lmsf_handeye_create(lmsf_handeye** out) {
  if (!out) return LMSF_ERR_INVALID;
  *out = new (std::nothrow) lmsf_handeye();
  return *out ? LMSF_OK : LMSF_ERR_INVALID;
}

void lmsf_handeye_destroy(lmsf_handeye* h) { delete h; }

static PoseQT from7(const double p[7]) {
  PoseQT r;
  r.q = qnormalize(Q4{p[0], p[1], p[2], p[3]});
  r.t = V3{p[4], p[5], p[6]};
  return r;
}

int lmsf_handeye_add_pose(lmsf_handeye* h, const double delta_primary[7], const double delta_sub[7], int* enough) {
  if (!h || !delta_primary || !delta_sub || !enough) return LMSF_ERR_INVALID;
  *enough = h->add_pose(from7(delta_primary), from7(delta_sub)) ? 1 : 0;
  return LMSF_OK;
}

int lmsf_handeye_calibrate(lmsf_handeye* h, double extrinsic[7], double singular_values[4], int* ok) {
  if (!h || !extrinsic || !ok) return LMSF_ERR_INVALID;
  *ok = 0;
  if (h->calib_rotation() && h->calib_translation()) *ok = 1;
  if (singular_values)
    for (int k = 0; k < 4; ++k) singular_values[k] = h->sv[k];
  if (*ok) {
    extrinsic[0] = h->ext_q.x, extrinsic[1] = h->ext_q.y, extrinsic[2] = h->ext_q.z, extrinsic[3] = h->ext_q.w;
    extrinsic[4] = h->ext_t.x, extrinsic[5] = h->ext_t.y, extrinsic[6] = h->ext_t.z;
  }
  return LMSF_OK;
}

int lmsf_handeye_size(lmsf_handeye* h, int* n) {
  if (!h || !n) return LMSF_ERR_INVALID;
  *n = (int)h->pose_storage.size();
  return LMSF_OK;
}

}  // extern "C"
