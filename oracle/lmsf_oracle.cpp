// lmsf_oracle.cpp — CPU ORACLE of the LMSF-SLAM scan-to-map registration hot path.
//
// TEST INFRASTRUCTURE ONLY.  PARITY PINNED IN PART (see lmsf_oracle.h): feature
// extraction (rows a1.*), the common point-cloud process (f4) and the exact kNN
// (a3.2), the alignment score (f2) and the matchers' control flow (a4) are checked
// bit for bit against the reference's own code compiled into oracle/_ref; voxel
// grid, Eigen's solver arithmetic, the GN / LM solves and the tracker (a2, a5, a6)
// rest on absent third-party arithmetic and are PARITY UNPINNED restatements of
// the source text.  Every function cites the reference file:line it follows (paths
// relative to src/MultiSensorFusionEstimator3D/include/).
//
// Floating-point contract: build with -ffp-contract=off (the reference builds
// -O3 without -march, i.e. no FMA; CMakeLists.txt:10-12).  float expressions in
// the reference stay float here, double stay double.
#include "lmsf_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <vector>

#include "oracle_math.h"

using namespace orc;

namespace {

struct P4 {
  float x, y, z, i;
};

// ===================================================================== kNN
struct Nb5 {
  float d[5];
  int id[5];
  int n;
  void clear() { n = 0; }
  float worst() const { return n < 5 ? std::numeric_limits<float>::infinity() : d[4]; }
  // ascending by (distance, index)
  void add(float dd, int ii) {
    if (n == 5) {
      if (dd > d[4] || (dd == d[4] && ii > id[4])) return;
    }
    int k = (n < 5) ? n : 4;
    while (k > 0 && (d[k - 1] > dd || (d[k - 1] == dd && id[k - 1] > ii))) {
      d[k] = d[k - 1];
      id[k] = id[k - 1];
      --k;
    }
    d[k] = dd;
    id[k] = ii;
    if (n < 5) ++n;
  }
};

// FLANN L2_Simple<float>: ((dx*dx)+dy*dy)+dz*dz accumulated in fp32
static inline float sqdist(const P4& a, float qx, float qy, float qz) {
  float dx = a.x - qx, dy = a.y - qy, dz = a.z - qz;
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

// Exact kd-tree standing in for pcl::KdTreeFLANN (FLANN KDTreeSingleIndex,
// leaf 15, eps 0; FeatureMatch/FeatureMatchBase.hpp:40-44).
struct KdTree {
  struct Node {
    int lo, hi, dim, left, right;
    float split;
  };
  const P4* pts = nullptr;
  std::vector<int> idx;
  std::vector<Node> nodes;
  void build(const P4* p, int n) {
    pts = p;
    idx.resize(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    if (n > 0) rec(0, n);
  }
  int rec(int lo, int hi) {
    int me = (int)nodes.size();
    nodes.push_back(Node{lo, hi, -1, -1, -1, 0.f});
    if (hi - lo <= 15) return me;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = lo; k < hi; ++k) {
      const P4& p = pts[idx[k]];
      const float c[3] = {p.x, p.y, p.z};
      for (int d = 0; d < 3; ++d) {
        mn[d] = std::min(mn[d], c[d]);
        mx[d] = std::max(mx[d], c[d]);
      }
    }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    if (!(mx[dim] - mn[dim] > 0.f)) return me;  // all coincident: keep as a leaf
    int mid = (lo + hi) / 2;
    auto coord = [&](int id) { return dim == 0 ? pts[id].x : (dim == 1 ? pts[id].y : pts[id].z); };
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                     [&](int a, int b) { return coord(a) < coord(b) || (coord(a) == coord(b) && a < b); });
    float split = coord(idx[mid]);
    int l = rec(lo, mid);
    int r = rec(mid, hi);
    nodes[me].dim = dim;
    nodes[me].split = split;
    nodes[me].left = l;
    nodes[me].right = r;
    return me;
  }
  void search(int node, float qx, float qy, float qz, Nb5& nb) const {
    const Node& nd = nodes[node];
    if (nd.dim < 0) {
      for (int k = nd.lo; k < nd.hi; ++k) nb.add(sqdist(pts[idx[k]], qx, qy, qz), idx[k]);
      return;
    }
    float q = nd.dim == 0 ? qx : (nd.dim == 1 ? qy : qz);
    float diff = q - nd.split;
    int nearc = diff < 0.f ? nd.left : nd.right, farc = diff < 0.f ? nd.right : nd.left;
    search(nearc, qx, qy, qz, nb);
    if (diff * diff <= nb.worst()) search(farc, qx, qy, qz, nb);
  }
};

struct MapIndex {
  std::vector<P4> pts;
  KdTree tree;
  bool set = false;
  void assign(const P4* p, int n, int mode) {
    pts.assign(p, p + n);
    if (mode == 0) tree.build(pts.data(), n);
    set = true;
  }
  // nearestKSearch(point, 5): exact, ascending by (d2, index)
  void knn(float qx, float qy, float qz, int mode, Nb5& nb) const {
    nb.clear();
    if (mode == 0) {
      if (!pts.empty()) tree.search(0, qx, qy, qz, nb);
    } else {
      for (int i = 0; i < (int)pts.size(); ++i) nb.add(sqdist(pts[i], qx, qy, qz), i);
    }
  }
};

// ===================================================================== matchers
struct EdgeInfo {
  V3 n;
  double r;
  V3 a, b;
};
struct SurfInfo {
  V3 n;
  double D, r;
};

// EdgeFeatureMatch::Match (FeatureMatch/EdgeFeatureMatch.hpp:33-87)
static bool match_edge(const MapIndex& m, int mode, float px, float py, float pz, EdgeInfo& out) {
  if (!m.set) return false;
  Nb5 nb;
  m.knn(px, py, pz, mode, nb);
  if (nb.n < 5) return false;            // sqdist[4] is out of range in the reference: reject
  if (!(nb.d[4] < 1.0f)) return false;   // search_thresh_ (FeatureMatchBase.hpp:29), squared
  V3 pt[5], c{0, 0, 0};
  for (int j = 0; j < 5; ++j) {
    const P4& q = m.pts[nb.id[j]];
    pt[j] = V3{(double)q.x, (double)q.y, (double)q.z};
    c = c + pt[j];
  }
  c = V3{c.x / 5.0, c.y / 5.0, c.z / 5.0};
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 5; ++j) {
    V3 z = pt[j] - c;
    const double zz[3] = {z.x, z.y, z.z};
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) S[r * 3 + cc] = S[r * 3 + cc] + zz[r] * zz[cc];
  }
  double w[3], V[9];
  symeig<3>(S, w, V);
  if (!(w[2] > 3 * w[1])) return false;
  V3 u{V[0 * 3 + 2], V[1 * 3 + 2], V[2 * 3 + 2]};
  V3 a = 0.1 * u + c;
  V3 b = (-0.1) * u + c;
  V3 p{(double)px, (double)py, (double)pz};
  V3 nu = cross(p - a, p - b);
  V3 de = a - b;
  double den = norm(de);
  out.r = norm(nu) / den;
  V3 g = cross(de, nu);
  double gn = norm(g);
  out.n = (gn > 0.0) ? V3{g.x / gn, g.y / gn, g.z / gn} : g;  // Eigen normalized() of a zero vector
  out.a = a;
  out.b = b;
  return true;
}

// SurfFeatureMatch::Match (FeatureMatch/surfFeatureMatch.hpp:32-87)
static bool match_surf(const MapIndex& m, int mode, float px, float py, float pz, SurfInfo& out) {
  if (!m.set) return false;
  Nb5 nb;
  m.knn(px, py, pz, mode, nb);
  if (nb.n < 5) return false;
  if (!(nb.d[4] < 1.0f)) return false;
  double A[15], B[5] = {-1, -1, -1, -1, -1}, nn[3];
  for (int j = 0; j < 5; ++j) {
    const P4& q = m.pts[nb.id[j]];
    A[j * 3 + 0] = q.x;
    A[j * 3 + 1] = q.y;
    A[j * 3 + 2] = q.z;
  }
  qr_solve<5, 3>(A, B, nn);
  V3 n{nn[0], nn[1], nn[2]};
  double len = norm(n);
  double D = 1 / len;
  n = V3{n.x / len, n.y / len, n.z / len};
  for (int j = 0; j < 5; ++j) {
    if (std::fabs(n.x * A[j * 3 + 0] + n.y * A[j * 3 + 1] + n.z * A[j * 3 + 2] + D) > 0.2) return false;
  }
  V3 p{(double)px, (double)py, (double)pz};
  float distance = (float)(dot(n, p) + D);  // `float distance` (surfFeatureMatch.hpp:72)
  out.r = std::fabs(distance);
  if (distance >= 0) {
    out.n = n;
    out.D = D;
  } else {
    out.n = neg(n);
    out.D = -D;
  }
  return true;
}

// ===================================================================== feature extraction
struct Features {
  std::vector<P4> edge, surf;
};

// splitScan (FeatureExtract/LOAMFeatureProcessor_base.hpp:290-343).  ring id or -1.
static int ring_of(const P4& p, int n_scans, float min_d, float max_d) {
  float s = p.x * p.x + p.y * p.y;  // float products and sum (:300-301)
  // `sqrt(float)` resolves to std::sqrt(float): `using namespace std;` (src/apps/include/utility.hpp:51, reached through
  // ros_utils.hpp) precedes this header in the node's translation unit (MultiLidarSLAM_node.cpp:10-17) — confirmed by
  // compiling the reference's own header (oracle/ref_loam.cpp), which this function must match bit for bit
  double distance = (double)std::sqrt(s);
  if (distance > max_d || distance < min_d) return -1;
  double angle = std::atan(p.z / distance) * 180 / M_PI;
  int id = 0;
  if (n_scans == 16) {
    id = int((angle + 15) / 2 + 0.5);
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 32) {
    id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
    // the reference tests `N_SCANS_ < 0` (:320, a typo); a negative id is UB there, rejected here
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 64) {
    if (angle >= -8.83)
      id = int((2 - angle) * 3.0 + 0.5);
    else
      id = n_scans / 2 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return -1;
  } else {
    return -1;
  }
  return id;
}

// checkBadEdgePoint (:216-282)
static void check_bad(const std::vector<P4>& pc, std::vector<int>& disable) {
  int P = (int)pc.size();
  for (int j = 5; j < P - 6; j++) {
    double a0 = (double)std::atan2(pc[j].x, pc[j].y);  // two float arguments: std::atan2(float, float) (:223-224)
    double a1 = (double)std::atan2(pc[j + 1].x, pc[j + 1].y);
    double da = std::fabs(a0 - a1);
    if (da > M_PI) da = M_PI * 2 - da;
    if (da > 0.0175) {
      for (int k = -5; k <= 5; ++k) disable[j + k] = 1;
      j = j + 4;
      continue;
    }
    float s0 = pc[j].x * pc[j].x + pc[j].y * pc[j].y + pc[j].z * pc[j].z;
    float s1 = pc[j + 1].x * pc[j + 1].x + pc[j + 1].y * pc[j + 1].y + pc[j + 1].z * pc[j + 1].z;
    double d0 = (double)std::sqrt(s0), d1 = (double)std::sqrt(s1);  // std::sqrt(float) (:247-252)
    double ang;
    if (d0 < d1)
      ang = std::atan2(d0 * da, d1 - d0);
    else
      ang = std::atan2(d1 * da, d0 - d1);
    if (ang <= 0.17) {
      if (d0 < d1) {
        for (int k = 1; k <= 5; ++k) disable[j + k] = 1;
        j = j + 4;
      } else {
        for (int k = 0; k <= 5; ++k) disable[j - k] = 1;
      }
    }
  }
}

// curvature (:97-118): float sum left to right, then double squares
static inline double curvature(const std::vector<P4>& pc, int j) {
  float sx = pc[j - 5].x + pc[j - 4].x + pc[j - 3].x + pc[j - 2].x + pc[j - 1].x - 10 * pc[j].x + pc[j + 1].x +
             pc[j + 2].x + pc[j + 3].x + pc[j + 4].x + pc[j + 5].x;
  float sy = pc[j - 5].y + pc[j - 4].y + pc[j - 3].y + pc[j - 2].y + pc[j - 1].y - 10 * pc[j].y + pc[j + 1].y +
             pc[j + 2].y + pc[j + 3].y + pc[j + 4].y + pc[j + 5].y;
  float sz = pc[j - 5].z + pc[j - 4].z + pc[j - 3].z + pc[j - 2].z + pc[j - 1].z - 10 * pc[j].z + pc[j + 1].z +
             pc[j + 2].z + pc[j + 3].z + pc[j + 4].z + pc[j + 5].z;
  double dx = sx, dy = sy, dz = sz;
  return dx * dx + dy * dy + dz * dz;
}

// LOAMFeatureProcessorBase::Process (:59-126) + featureExtractionFromSector (:145-207)
static void extract(const lmsf_oracle_params& prm, const P4* in, int n, uint8_t* label, Features& out) {
  const int R = prm.n_scans;
  std::vector<std::vector<P4>> rings(R);
  std::vector<std::vector<int>> rid(R);
  if (label) std::memset(label, 0, n);
  for (int i = 0; i < n; ++i) {
    int id = ring_of(in[i], R, prm.min_range, prm.max_range);
    if (id < 0) continue;
    rings[id].push_back(in[i]);
    rid[id].push_back(i);
  }
  out.edge.clear();
  out.surf.clear();
  for (int r = 0; r < R; ++r) {
    const std::vector<P4>& pc = rings[r];
    int P = (int)pc.size();
    if (P < 20) continue;
    int total = P - 10;
    if (total < 6) continue;
    int len = (int)((total / 6) + 0.5);
    std::vector<int> disable(P, 0), is_edge(P, 0);
    if (prm.remove_bad_points) check_bad(pc, disable);
    for (int k = 0; k < 6; ++k) {
      int s = 5 + len * k, e = s + len - 1;
      if (k == 5) e = P - 6;
      std::vector<std::pair<double, int>> cur;
      cur.reserve(e - s + 1);
      for (int j = s; j <= e; ++j) cur.emplace_back(curvature(pc, j), j);
      // std::sort ascending by value (:152); ties are implementation-defined in the
      // reference, canonical here: (value, index)
      std::sort(cur.begin(), cur.end());
      int picked = 0;
      for (int i = (int)cur.size() - 1; i >= 0; --i) {
        int ind = cur[i].second;
        if (disable[ind] == 0) {
          if (cur[i].first <= prm.edge_thresh) break;
          picked++;
          if (picked <= 20) {
            out.edge.push_back(pc[ind]);
            is_edge[ind] = 1;
            if (label) label[rid[r][ind]] = 1;
          } else {
            break;
          }
          for (int m = 1; m <= 5; ++m) {
            int nn = ind + m >= P ? P - 1 : ind + m;
            disable[nn] = 1;
          }
          for (int m = -1; m >= -5; --m) {
            int nn = ind + m < 0 ? 0 : ind + m;
            disable[nn] = 1;
          }
        }
      }
      for (int i = 0; i < (int)cur.size(); ++i) {
        int ind = cur[i].second;
        if (is_edge[ind] == 0) {
          out.surf.push_back(pc[ind]);
          if (label) label[rid[r][ind]] = 2;
        }
      }
    }
  }
}

// ===================================================================== voxel grid
// pcl::VoxelGrid<PointXYZI>::applyFilter as called by FilterBase::Filter
// (Filter/filter_base.hpp:34-45); PCL is absent: semantics restated from PCL
// 1.7-1.12.  All arithmetic fp32.
static void voxel(const P4* in, int n, float leaf, std::vector<P4>& out, int32_t* vox_of_pt) {
  out.clear();
  if (vox_of_pt)
    for (int i = 0; i < n; ++i) vox_of_pt[i] = -1;
  float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nfin = 0;
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z)) continue;
    ++nfin;
    const float c[3] = {in[i].x, in[i].y, in[i].z};
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], c[d]);
      mx[d] = std::max(mx[d], c[d]);
    }
  }
  if (nfin == 0) return;
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
    // "Leaf size is too small for the input dataset": output = input
    out.assign(in, in + n);
    if (vox_of_pt)
      for (int i = 0; i < n; ++i) vox_of_pt[i] = i;
    return;
  }
  int minb[3], maxb[3], div[3];
  for (int d = 0; d < 3; ++d) {
    minb[d] = (int)std::floor(mn[d] * inv);
    maxb[d] = (int)std::floor(mx[d] * inv);
    div[d] = maxb[d] - minb[d] + 1;
  }
  int mul[3] = {1, div[0], div[0] * div[1]};
  std::vector<std::pair<int, int>> iv;
  iv.reserve(nfin);
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z)) continue;
    int i0 = (int)(std::floor(in[i].x * inv) - (float)minb[0]);
    int i1 = (int)(std::floor(in[i].y * inv) - (float)minb[1]);
    int i2 = (int)(std::floor(in[i].z * inv) - (float)minb[2]);
    iv.emplace_back(i0 * mul[0] + i1 * mul[1] + i2 * mul[2], i);
  }
  std::sort(iv.begin(), iv.end());  // canonical: (voxel, point index)
  size_t k = 0;
  while (k < iv.size()) {
    size_t e = k;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    while (e < iv.size() && iv[e].first == iv[k].first) {
      const P4& p = in[iv[e].second];
      sx += p.x;
      sy += p.y;
      sz += p.z;
      si += p.i;
      if (vox_of_pt) vox_of_pt[iv[e].second] = (int32_t)out.size();
      ++e;
    }
    float cnt = (float)(e - k);
    out.push_back(P4{sx / cnt, sy / cnt, sz / cnt, si / cnt});
    k = e;
  }
}

// ===================================================================== LM (Ceres restated)
struct EdgeBlk {
  V3 pl, a, b;
};
struct SurfBlk {
  V3 pl, n;
  double D;
};
struct Normal {
  double H[36], g[6], cost;
};

// ceres::HuberLoss(a) + Corrector (rho'' <= 0 => scale r and J by sqrt(rho'))
static inline void huber(double a, double s, double& rho0, double& scale) {
  double b = a * a;
  if (s > b) {
    double r = std::sqrt(s);
    rho0 = 2 * a * r - b;
    double rho1 = std::max(DBL_MIN, a / r);
    scale = std::sqrt(rho1);
  } else {
    rho0 = s;
    scale = 1.0;
  }
}

static inline void accum(Normal& N, const double J[6], double r, double huber_a) {
  double rho0, sc;
  huber(huber_a, r * r, rho0, sc);
  N.cost += 0.5 * rho0;
  double Js[6], rs = sc * r;
  for (int i = 0; i < 6; ++i) Js[i] = sc * J[i];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) N.H[i * 6 + j] += Js[i] * Js[j];
    N.g[i] += Js[i] * rs;
  }
}

// se3PointEdgeFactor::Evaluate (ceres_factor/edge_factor.hpp:33-61),
// se3PointSurfFactor::Evaluate (ceres_factor/surf_factor.hpp:32-56);
// local Jacobian = first 6 columns (PoseSE3Parameterization.hpp:54-60).
static inline void edge_factor(const Quat& q, const V3& t, const EdgeBlk& e, bool want_jac, double& r, double J[6]) {
  V3 lp = qrot(q, e.pl) + t;
  V3 nu = cross(lp - e.a, lp - e.b);
  V3 de = e.a - e.b;
  double den = norm(de), nun = norm(nu);
  r = nun / den;
  for (int i = 0; i < 6; ++i) J[i] = 0;
  if (want_jac) {
    V3 u{nu.x / nun, nu.y / nun, nu.z / nun};
    V3 g = cross(de, u);   // -(u^T skew(de))
    V3 jr = cross(lp, g);  //  g^T (-skew(lp))
    J[0] = jr.x / den;
    J[1] = jr.y / den;
    J[2] = jr.z / den;
    J[3] = g.x / den;
    J[4] = g.y / den;
    J[5] = g.z / den;
  }
}

static inline void surf_factor(const Quat& q, const V3& t, const SurfBlk& s, bool want_jac, double& r, double J[6]) {
  V3 lp = qrot(q, s.pl) + t;
  r = dot(s.n, lp) + s.D;
  for (int i = 0; i < 6; ++i) J[i] = 0;
  if (want_jac) {
    V3 jr = cross(lp, s.n);
    J[0] = jr.x;
    J[1] = jr.y;
    J[2] = jr.z;
    J[3] = s.n.x;
    J[4] = s.n.y;
    J[5] = s.n.z;
  }
}

static void lm_evaluate(const std::vector<EdgeBlk>& eb, const std::vector<SurfBlk>& sb, const double x[7],
                        double huber_a, bool want_jac, Normal& N) {
  std::memset(&N, 0, sizeof N);
  Quat q{x[0], x[1], x[2], x[3]};
  V3 t{x[4], x[5], x[6]};
  double r, J[6];
  for (const EdgeBlk& e : eb) {
    edge_factor(q, t, e, want_jac, r, J);
    accum(N, J, r, huber_a);
  }
  for (const SurfBlk& s : sb) {
    surf_factor(q, t, s, want_jac, r, J);
    accum(N, J, r, huber_a);
  }
}

// PoseSE3Parameterization::Plus (Algorithm/Ceres/Parameterization/PoseSE3Parameterization.hpp:32-46)
static void se3_plus(const double x[7], const double d[6], double out[7]) {
  Quat dq;
  V3 dt;
  se3_exp(d, dq, dt);
  Quat q{x[0], x[1], x[2], x[3]};
  Quat qp = qmul(dq, q);
  V3 tp = qrot(dq, V3{x[4], x[5], x[6]}) + dt;
  out[0] = qp.x;
  out[1] = qp.y;
  out[2] = qp.z;
  out[3] = qp.w;
  out[4] = tp.x;
  out[5] = tp.y;
  out[6] = tp.z;
}

static double norm7(const double* x) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s += x[i] * x[i];
  return std::sqrt(s);
}

// The Huber width travels through the ABI as a float, the reference writes the double literal `HuberLoss(0.1)`
// (ceres_edgeSurfFeatureRegistration.hpp:107): (double)0.1f = 0.10000000149 moves the pose by 1e-9 m against the
// reference's own code (oracle/ref_lm.cpp found it).  A float parameter is therefore read as the decimal constant it was
// written as: its shortest 6-significant-digit decimal, converted to double.
static double decimal_of_float(float f) {
  char buf[48];
  std::snprintf(buf, sizeof buf, "%.6g", (double)f);
  return std::strtod(buf, nullptr);
}

// ceres::Solve with DENSE_QR, Levenberg-Marquardt trust region, Jacobi scaling,
// max_num_iterations (ceres_edgeSurfFeatureRegistration.hpp:116-123).  Ceres is
// absent and unpinned: TrustRegionMinimizer / LevenbergMarquardtStrategy of
// Ceres 1.14 restated.  The regularised least-squares step
// is solved through its normal equations (Cholesky) instead of QR of [J; D].
// `evaluate(x, want_jac, Normal&)` and `plus(x, delta, out)` are the problem: the library's own blocks (lm_solve
// below) or a caller's residual callbacks (lmsf_oracle_lm_solve_cb, used to run the reference's own cost functions).
template <class Eval, class Plus>
static void lm_solve_t(Eval&& evaluate, Plus&& plus, bool empty, int max_iters, double x[7], int& steps, int& accepted,
                       double& cost_out) {
  Normal N;
  evaluate(x, true, N);
  double cost = N.cost;
  cost_out = cost;
  steps = 0;
  accepted = 0;
  if (empty) return;
  double scale[6];
  for (int j = 0; j < 6; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(N.H[j * 6 + j]));
  double radius = 1e4, decrease = 2.0;
  double x_norm = norm7(x);
  int invalid = 0;
  auto gmax = [&]() {
    double m = 0;
    for (int j = 0; j < 6; ++j) m = std::max(m, std::fabs(N.g[j]));
    return m;
  };
  if (gmax() <= 1e-10) return;
  for (int iter = 1; iter <= max_iters; ++iter) {
    ++steps;
    double Hs[36], gs[6], A[36], y[6], step[6];
    for (int i = 0; i < 6; ++i) {
      gs[i] = scale[i] * N.g[i];
      for (int j = 0; j < 6; ++j) Hs[i * 6 + j] = scale[i] * N.H[i * 6 + j] * scale[j];
    }
    std::memcpy(A, Hs, sizeof A);
    for (int j = 0; j < 6; ++j) {
      double dj = std::min(std::max(Hs[j * 6 + j], 1e-6), 1e32);
      A[j * 6 + j] += dj / radius;
    }
    bool ok = chol_solve6(A, gs, y);
    double mcc = 0.0;
    if (ok) {
      for (int j = 0; j < 6; ++j) {
        step[j] = -y[j];
        if (!std::isfinite(step[j])) ok = false;
      }
    }
    if (ok) {
      double sg = 0, sHs = 0;
      for (int i = 0; i < 6; ++i) {
        sg += step[i] * gs[i];
        double row = 0;
        for (int j = 0; j < 6; ++j) row += Hs[i * 6 + j] * step[j];
        sHs += step[i] * row;
      }
      mcc = -(sg + 0.5 * sHs);  // -(J s)^T (r + J s / 2)
    }
    if (!ok || !(mcc > 0.0)) {
      if (++invalid >= 5) break;
      radius *= 0.5;
      if (radius < 1e-32) break;
      continue;
    }
    invalid = 0;
    double delta[6], cand[7];
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    plus(x, delta, cand);
    Normal Nc;
    evaluate(cand, false, Nc);
    double diff[7];
    for (int i = 0; i < 7; ++i) diff[i] = x[i] - cand[i];
    if (norm7(diff) <= 1e-8 * (x_norm + 1e-8)) break;   // parameter tolerance: candidate NOT taken
    double cc = cost - Nc.cost;
    if (std::fabs(cc) <= 1e-6 * cost) break;             // function tolerance: candidate NOT taken
    double rho = cc / mcc;
    if (rho > 1e-3) {
      std::memcpy(x, cand, sizeof(double) * 7);
      x_norm = norm7(x);
      evaluate(x, true, N);
      cost = N.cost;
      ++accepted;
      double u = 2.0 * rho - 1.0;  // Ceres: pow(2 rho - 1, 3); u*u*u differs from it by <= 1 ulp
      radius = radius / std::max(1.0 / 3.0, 1.0 - u * u * u);
      radius = std::min(1e16, radius);
      decrease = 2.0;
      if (gmax() <= 1e-10) break;
    } else {
      radius = radius / decrease;
      decrease *= 2.0;
    }
    if (radius < 1e-32) break;
  }
  cost_out = cost;
}

static void lm_solve(const std::vector<EdgeBlk>& eb, const std::vector<SurfBlk>& sb, double huber_a, int max_iters,
                     double x[7], int& steps, int& accepted, double& cost_out) {
  lm_solve_t([&](const double* xx, bool jac, Normal& N) { lm_evaluate(eb, sb, xx, huber_a, jac, N); },
             [&](const double* xx, const double* d, double* out) { se3_plus(xx, d, out); }, eb.empty() && sb.empty(),
             max_iters, x, steps, accepted, cost_out);
}

}  // namespace

// ===================================================================== context
struct lmsf_oracle_ctx {
  lmsf_oracle_params prm;
  MapIndex map[2];
  int lm_count;
  // tracker (LidarTracker/LidarTrackerLocalMap.hpp:46-59)
  bool init = false;
  Iso prev, curr, motion, last_kf;
  double last_kf_time = 0;
  std::deque<std::vector<P4>> win[2];
  std::vector<P4> map_cat[2];
};

namespace {

static int nthreads(const lmsf_oracle_ctx* c) { return c->prm.threads > 1 ? c->prm.threads : 1; }

// pointAssociateToMap (edgeSurfFeatureRegistration.hpp:342-350, ceres_...:235-244)
static inline void to_map(const Quat& q, const V3& t, const P4& p, float& x, float& y, float& z) {
  V3 w = qrot(q, V3{(double)p.x, (double)p.y, (double)p.z}) + t;
  x = (float)w.x;
  y = (float)w.y;
  z = (float)w.z;
}

// EdgeSurfFeatureRegistration::Solve / GNOptimization (edgeSurfFeatureRegistration.hpp:113-330)
static void solve_gn(lmsf_oracle_ctx* c, const P4* edge, int ne, const P4* surf, int ns, double pose[7],
                     lmsf_oracle_reg_stats* st) {
  Quat q{pose[0], pose[1], pose[2], pose[3]};
  V3 t{pose[4], pose[5], pose[6]};
  const int mode = c->prm.knn_mode;
  bool degenerate = false;
  double Map[36];
  std::vector<uint8_t> ok_s(ns), ok_e(ne);
  std::vector<SurfInfo> si(ns);
  std::vector<EdgeInfo> ei(ne);
  int it = 0, n_e = 0, n_s = 0, converged = 0;
  double cost = 0;
  for (it = 0; it < c->prm.gn_max_iters; ++it) {
#pragma omp parallel for num_threads(nthreads(c)) schedule(static)
    for (int i = 0; i < ns; ++i) {
      float x, y, z;
      to_map(q, t, surf[i], x, y, z);
      ok_s[i] = match_surf(c->map[1], mode, x, y, z, si[i]);
    }
#pragma omp parallel for num_threads(nthreads(c)) schedule(static)
    for (int i = 0; i < ne; ++i) {
      float x, y, z;
      to_map(q, t, edge[i], x, y, z);
      ok_e[i] = match_edge(c->map[0], mode, x, y, z, ei[i]);
    }
    n_e = 0;
    n_s = 0;
    for (int i = 0; i < ne; ++i) n_e += ok_e[i];
    for (int i = 0; i < ns; ++i) n_s += ok_s[i];
    // counters are uint16_t in the reference (:58-59) and wrap above 65 535; 32-bit here
    if (n_e + n_s < 10) continue;  // GNOptimization returns false (:221-225)
    double R[9];
    q2R(q, R);
    double H[36], g[6];
    std::memset(H, 0, sizeof H);
    std::memset(g, 0, sizeof g);
    cost = 0;
    auto row = [&](const P4& pl, const V3& grad, double res) {
      float residual = (float)res;  // `float residual` (:234)
      double px = pl.x, py = pl.y, pz = pl.z;
      // -R * skew(p)
      double sk[9] = {0, -pz, py, pz, 0, -px, -py, px, 0};
      double A[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i * 3 + j] = -(R[i * 3 + 0] * sk[0 * 3 + j] + R[i * 3 + 1] * sk[1 * 3 + j] + R[i * 3 + 2] * sk[2 * 3 + j]);
      double J[6];
      for (int j = 0; j < 3; ++j) J[j] = grad.x * A[0 * 3 + j] + grad.y * A[1 * 3 + j] + grad.z * A[2 * 3 + j];
      J[3] = grad.x;
      J[4] = grad.y;
      J[5] = grad.z;
      double rr = residual;
      for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) H[i * 6 + j] += J[i] * J[j];
        g[i] += J[i] * rr;
      }
      cost += 0.5 * rr * rr;
    };
    for (int i = 0; i < ne; ++i)
      if (ok_e[i]) row(edge[i], ei[i].n, ei[i].r);
    for (int i = 0; i < ns; ++i)
      if (ok_s[i]) row(surf[i], si[i].n, si[i].r);
    double mg[6], X[6];
    for (int i = 0; i < 6; ++i) mg[i] = -g[i];
    qr_solve<6, 6>(H, mg, X);
    if (it == 0) {
      double w[6], V[36], V2[36], Vi[36];
      symeig<6>(H, w, V);
      std::memcpy(V2, V, sizeof V);
      degenerate = false;
      float thresh = 100;
      for (int i = 5; i >= 0; --i) {  // literal: starts at the LARGEST eigenvalue (:289-302)
        if (w[i] < thresh) {
          for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.0;
          degenerate = true;
        } else {
          break;
        }
      }
      inv6(V, Vi);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += Vi[i * 6 + k] * V2[k * 6 + j];
          Map[i * 6 + j] = s;
        }
    }
    if (degenerate) {
      double Y[6];
      for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Map[i * 6 + k] * X[k];
        Y[i] = s;
      }
      std::memcpy(X, Y, sizeof X);
    }
    t.x += X[3];
    t.y += X[4];
    t.z += X[5];
    V3 dr{X[0], X[1], X[2]};
    double dn = norm(dr);
    V3 axis = dn > 0.0 ? V3{dr.x / dn, dr.y / dn, dr.z / dn} : dr;
    double ang = dn / 2;  // AngleAxisd(|d|/2, d^) (:317): half-angle quirk
    double sh = std::sin(0.5 * ang), ch = std::cos(0.5 * ang);
    Quat dq{sh * axis.x, sh * axis.y, sh * axis.z, ch};
    q = qmul(q, dq);
    float deltaR = (float)(dn / 2);
    float deltaT = (float)std::sqrt(std::pow(X[3] * 100, 2) + std::pow(X[4] * 100, 2) + std::pow(X[5] * 100, 2));
    if (deltaR < 0.0009 && deltaT < 0.05) {
      converged = 1;
      break;
    }
  }
  pose[0] = q.x;
  pose[1] = q.y;
  pose[2] = q.z;
  pose[3] = q.w;
  pose[4] = t.x;
  pose[5] = t.y;
  pose[6] = t.z;
  if (st) {
    std::memset(st, 0, sizeof *st);
    st->outer_iters = converged ? it + 1 : it;
    st->n_edge_matched = n_e;
    st->n_surf_matched = n_s;
    st->converged = converged;
    st->degenerate = degenerate;
    st->final_cost = cost;
  }
}

// CeresEdgeSurfFeatureRegistration::Solve (ceres_edgeSurfFeatureRegistration.hpp:96-130)
static void solve_lm(lmsf_oracle_ctx* c, const P4* edge, int ne, const P4* surf, int ns, double pose[7],
                     lmsf_oracle_reg_stats* st) {
  if (c->lm_count > 2) c->lm_count--;
  const int mode = c->prm.knn_mode;
  double x[7];
  std::memcpy(x, pose, sizeof x);
  std::vector<uint8_t> ok_s(ns), ok_e(ne);
  std::vector<SurfInfo> si(ns);
  std::vector<EdgeInfo> ei(ne);
  std::vector<EdgeBlk> eb;
  std::vector<SurfBlk> sb;
  int total = 0, acc = 0;
  double cost = 0;
  for (int it = 0; it < c->lm_count; ++it) {
    Quat q{x[0], x[1], x[2], x[3]};
    V3 t{x[4], x[5], x[6]};
    bool has_e = c->map[0].set, has_s = c->map[1].set;
#pragma omp parallel for num_threads(nthreads(c)) schedule(static)
    for (int i = 0; i < ne; ++i) {
      float px, py, pz;
      to_map(q, t, edge[i], px, py, pz);
      ok_e[i] = has_e && match_edge(c->map[0], mode, px, py, pz, ei[i]);
    }
#pragma omp parallel for num_threads(nthreads(c)) schedule(static)
    for (int i = 0; i < ns; ++i) {
      float px, py, pz;
      to_map(q, t, surf[i], px, py, pz);
      ok_s[i] = has_s && match_surf(c->map[1], mode, px, py, pz, si[i]);
    }
    eb.clear();
    sb.clear();
    for (int i = 0; i < ne; ++i)
      if (ok_e[i]) eb.push_back(EdgeBlk{V3{(double)edge[i].x, (double)edge[i].y, (double)edge[i].z}, ei[i].a, ei[i].b});
    for (int i = 0; i < ns; ++i)
      if (ok_s[i]) sb.push_back(SurfBlk{V3{(double)surf[i].x, (double)surf[i].y, (double)surf[i].z}, si[i].n, si[i].D});
    int s = 0, a = 0;
    lm_solve(eb, sb, decimal_of_float(c->prm.huber_delta), c->prm.lm_inner_iters, x, s, a, cost);
    total += s;
    acc += a;
  }
  std::memcpy(pose, x, sizeof x);
  if (st) {
    std::memset(st, 0, sizeof *st);
    st->outer_iters = c->lm_count;
    st->n_edge_matched = (int)eb.size();
    st->n_surf_matched = (int)sb.size();
    st->lm_steps_total = total;
    st->lm_steps_accepted = acc;
    st->final_cost = cost;
  }
}

static void do_register(lmsf_oracle_ctx* c, const P4* edge, int ne, const P4* surf, int ns, int solver, double pose[7],
                        lmsf_oracle_reg_stats* st) {
  if (solver == 0)
    solve_gn(c, edge, ne, surf, ns, pose, st);
  else
    solve_lm(c, edge, ne, surf, ns, pose, st);
}

// pcl::transformPointCloud(in, out, Matrix4d) (LidarTrackerLocalMap.hpp:217):
// double arithmetic, result stored to float; intensity copied.
static void transform_cloud(const std::vector<P4>& in, const Iso& T, std::vector<P4>& out) {
  out.resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    double x = in[i].x, y = in[i].y, z = in[i].z;
    out[i].x = (float)(T.R[0] * x + T.R[1] * y + T.R[2] * z + T.t[0]);
    out[i].y = (float)(T.R[3] * x + T.R[4] * y + T.R[5] * z + T.t[1]);
    out[i].z = (float)(T.R[6] * x + T.R[7] * y + T.R[8] * z + T.t[2]);
    out[i].i = in[i].i;
  }
}

// updateLocalMap (LidarTrackerLocalMap.hpp:205-232) over the (missing from the
// tree) PointCloudLocalMapBase: sliding window of `window` frames; the map is
// their concatenation oldest -> newest.  AddFrameForMotion appends and evicts
// the oldest frame beyond the window; AddFrameForTime replaces the newest
// frame (refresh without growing).  Inferred contract (SURVEY.md §8 a6').
static void update_map(lmsf_oracle_ctx* c, const Features& f, const Iso& T, int type) {
  const std::vector<P4>* clouds[2] = {&f.edge, &f.surf};
  const float leaf[2] = {c->prm.map_leaf_edge, c->prm.map_leaf_surf};
  for (int k = 0; k < 2; ++k) {
    if (clouds[k]->empty()) continue;
    std::vector<P4> tc;
    transform_cloud(*clouds[k], T, tc);
    if (type == 1) {
      c->win[k].push_back(std::move(tc));
      if ((int)c->win[k].size() > c->prm.window) c->win[k].pop_front();
    } else {
      if (!c->win[k].empty()) c->win[k].pop_back();
      c->win[k].push_back(std::move(tc));
    }
    std::vector<P4> cat;
    for (const auto& fr : c->win[k]) cat.insert(cat.end(), fr.begin(), fr.end());
    if (leaf[k] > 0.f) {
      std::vector<P4> ds;
      voxel(cat.data(), (int)cat.size(), leaf[k], ds, nullptr);
      cat.swap(ds);
    }
    c->map_cat[k].swap(cat);
    if (!c->map_cat[k].empty())  // ceres_edgeSurfFeatureRegistration.hpp:58
      c->map[k].assign(c->map_cat[k].data(), (int)c->map_cat[k].size(), c->prm.knn_mode);
  }
}

// needUpdataLocalMap (LidarTrackerLocalMap.hpp:239-262)
static int need_update(lmsf_oracle_ctx* c, const Iso& curr, double stamp) {
  if (stamp - c->last_kf_time > c->prm.kf_time) return 2;
  Iso d = iso_mul(iso_inv(c->last_kf), curr);
  double dt = std::sqrt(d.t[0] * d.t[0] + d.t[1] * d.t[1] + d.t[2] * d.t[2]);
  Quat q = R2q(d.R);
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  double w = q.w / n;
  double da = std::acos(w) * 2;
  if (dt > c->prm.kf_trans || da > c->prm.kf_rot) return 1;
  return 0;
}

// The node's removeNaNFromPointCloud (src/apps/src/MultiLidarSLAM_node.cpp:126-133) followed by
// RotaryLidarPreProcess<PointXYZI>::Process (Preprocess/RotaryLidar_preprocessing.hpp:31-71, findStartEndAngle :80-94,
// setPoint :100-104): intensity := relative time of the point in the sweep.  atan2 of two floats is the float overload
// (the C library's atan2f); comparisons and the +- 2 pi happen in double and are stored back to float, as the
// reference's expressions promote them; rel_time is float arithmetic.
static void rotary_preprocess(const P4* in, int n, float period, std::vector<P4>& out) {
  out.clear();
  for (int i = 0; i < n; ++i)
    if (std::isfinite(in[i].x) && std::isfinite(in[i].y) && std::isfinite(in[i].z)) out.push_back(in[i]);
  if (out.empty()) return;
  float start_ori = -std::atan2(out.front().y, out.front().x);
  float end_ori = -std::atan2(out.back().y, out.back().x) + 2 * M_PI;
  if (end_ori - start_ori > 3 * M_PI)
    end_ori -= 2 * M_PI;
  else if (end_ori - start_ori < M_PI)
    end_ori += 2 * M_PI;
  bool half_passed = false;
  for (size_t i = 0; i < out.size(); ++i) {
    float ori = -std::atan2(out[i].y, out[i].x);
    if (!half_passed) {
      if (ori < start_ori - M_PI / 2)
        ori += 2 * M_PI;
      else if (ori > start_ori + M_PI * 3 / 2)
        ori -= 2 * M_PI;
      if (ori - start_ori > M_PI) half_passed = true;
    } else {
      ori += 2 * M_PI;
      if (ori < end_ori - M_PI * 3 / 2)
        ori += 2 * M_PI;
      else if (ori > end_ori + M_PI / 2)
        ori -= 2 * M_PI;
    }
    out[i].i = (ori - start_ori) / (end_ori - start_ori) * period;
  }
}

// rotary_scan_period > 0: the sweep the extraction sees is the preprocessed one.  The points keep their places (labels are
// by input index); the non-finite ones, which removeNaN drops, are passed over by the extraction anyway (ring_of).
static const P4* preprocessed(const lmsf_oracle_params& prm, const P4* in, int n, std::vector<P4>& work) {
  if (!(prm.rotary_scan_period > 0.f) || n <= 0) return in;
  std::vector<P4> dense;
  rotary_preprocess(in, n, prm.rotary_scan_period, dense);
  work.assign(in, in + n);
  size_t k = 0;
  for (int i = 0; i < n; ++i)
    if (std::isfinite(in[i].x) && std::isfinite(in[i].y) && std::isfinite(in[i].z)) work[i].i = dense[k++].i;
  return work.data();
}

static void scan_filter(lmsf_oracle_ctx* c, Features& f) {
  if (c->prm.scan_leaf_edge > 0.f) {
    std::vector<P4> ds;
    voxel(f.edge.data(), (int)f.edge.size(), c->prm.scan_leaf_edge, ds, nullptr);
    f.edge.swap(ds);
  }
  if (c->prm.scan_leaf_surf > 0.f) {
    std::vector<P4> ds;
    voxel(f.surf.data(), (int)f.surf.size(), c->prm.scan_leaf_surf, ds, nullptr);
    f.surf.swap(ds);
  }
}

// LidarTrackerLocalMap::Solve (LidarTrackerLocalMap.hpp:107-160)
static void tracker_solve(lmsf_oracle_ctx* c, const Features& f, double stamp, double delta[7], double pose_out[7],
                          lmsf_oracle_track_stats* st) {
  lmsf_oracle_track_stats s;
  std::memset(&s, 0, sizeof s);
  s.n_edge = (int)f.edge.size();
  s.n_surf = (int)f.surf.size();
  if (!c->init) {
    update_map(c, f, iso_identity(), 1);
    c->curr = c->prev = c->motion = c->last_kf = iso_identity();
    c->last_kf_time = stamp;
    c->init = true;
    s.first = 1;
    s.keyframe = 1;
  } else {
    bool ident = delta[0] == 0 && delta[1] == 0 && delta[2] == 0 && delta[3] == 1 && delta[4] == 0 && delta[5] == 0 &&
                 delta[6] == 0;
    if (ident)
      c->curr = iso_mul(c->prev, c->motion);
    else
      c->curr = iso_mul(c->prev, iso_from_pose7(delta));
    // RegistrationLocalMap (:168-177): SetInputTarget + Solve(Isometry3d&)
    double p[7];
    iso_to_pose7(c->curr, p);  // Quaterniond(T.rotation())
    do_register(c, f.edge.data(), (int)f.edge.size(), f.surf.data(), (int)f.surf.size(), c->prm.solver, p, &s.reg);
    c->curr = iso_from_pose7(p);  // T.linear() = q.toRotationMatrix()
    c->motion = iso_mul(iso_inv(c->prev), c->curr);
    iso_to_pose7(c->motion, delta);
    c->prev = c->curr;
    int ut = need_update(c, c->curr, stamp);
    s.keyframe = ut;
    if (ut) {
      c->last_kf = c->curr;
      c->last_kf_time = stamp;
      update_map(c, f, c->curr, ut);
    }
  }
  s.map_edge = (int)c->map_cat[0].size();
  s.map_surf = (int)c->map_cat[1].size();
  iso_to_pose7(c->curr, pose_out);
  if (st) *st = s;
}

}  // namespace

// ===================================================================== C interface
extern "C" {

int lmsf_oracle_params_default(lmsf_oracle_params* p) {
  if (!p) return -1;
  std::memset(p, 0, sizeof *p);
  p->n_scans = 16;
  p->min_range = 2.f;
  p->max_range = 80.f;
  p->edge_thresh = 1.f;
  p->remove_bad_points = 1;
  p->max_points = 262144;
  p->window = 10;
  p->solver = 1;
  p->gn_max_iters = 10;
  p->lm_outer_start = 10;
  p->lm_inner_iters = 4;
  p->huber_delta = 0.1f;
  p->kf_trans = 0.3;
  p->kf_rot = 0.1;
  p->kf_time = 10.0;
  p->max_map_points = 0;
  p->knn_mode = 0;
  p->threads = 1;
  return 0;
}

int lmsf_oracle_ctx_create(int, const lmsf_oracle_params* p, lmsf_oracle_ctx** out) {
  if (!p || !out) return -1;
  if (p->n_scans != 16 && p->n_scans != 32 && p->n_scans != 64) return -1;
  lmsf_oracle_ctx* c = new lmsf_oracle_ctx();
  c->prm = *p;
  c->lm_count = p->lm_outer_start;
  *out = c;
  return 0;
}

void lmsf_oracle_ctx_destroy(lmsf_oracle_ctx* c) { delete c; }

const char* lmsf_oracle_strerror(int code) { return code == 0 ? "ok" : "oracle: invalid argument"; }

int lmsf_oracle_extract_features(lmsf_oracle_ctx* c, const float* xyzi, int n, uint8_t* label_out, float* edge_xyzi,
                                 int* n_edge, float* surf_xyzi, int* n_surf) {
  if (!c || n < 0 || (n > 0 && !xyzi)) return -1;
  Features f;
  std::vector<P4> work;
  extract(c->prm, preprocessed(c->prm, reinterpret_cast<const P4*>(xyzi), n, work), n, label_out, f);
  if (edge_xyzi && !f.edge.empty()) std::memcpy(edge_xyzi, f.edge.data(), f.edge.size() * sizeof(P4));
  if (surf_xyzi && !f.surf.empty()) std::memcpy(surf_xyzi, f.surf.data(), f.surf.size() * sizeof(P4));
  if (n_edge) *n_edge = (int)f.edge.size();
  if (n_surf) *n_surf = (int)f.surf.size();
  return 0;
}

int lmsf_oracle_rotary_preprocess(lmsf_oracle_ctx* c, const float* xyzi, int n, float scan_period, float* out_xyzi,
                                  int* n_out) {
  if (!c || n < 0 || !n_out || !(scan_period > 0.f) || (n > 0 && (!xyzi || !out_xyzi))) return -1;
  std::vector<P4> out;
  rotary_preprocess(reinterpret_cast<const P4*>(xyzi), n, scan_period, out);
  if (!out.empty()) std::memcpy(out_xyzi, out.data(), out.size() * sizeof(P4));
  *n_out = (int)out.size();
  return 0;
}

int lmsf_oracle_voxel_downsample(lmsf_oracle_ctx* c, const float* xyzi, int n, float leaf, float* out_xyzi, int* n_out,
                                 int32_t* voxel_of_point) {
  if (!c || n < 0 || !(leaf > 0.f) || !n_out) return -1;
  std::vector<P4> out;
  voxel(reinterpret_cast<const P4*>(xyzi), n, leaf, out, voxel_of_point);
  if (out_xyzi && !out.empty()) std::memcpy(out_xyzi, out.data(), out.size() * sizeof(P4));
  *n_out = (int)out.size();
  return 0;
}

// PointCloudCommonProcess::Process (processing/common_processing.hpp:87-111): removeNaN, VoxelGrid, (no outlier
// removal), DistanceFilter (Filter/distance_filter.hpp:24-44: fp32 Vector3f norm widened to double, open interval)
int lmsf_oracle_common_process(lmsf_oracle_ctx* c, const float* xyzi, int n, int remove_nan, float leaf, float dist_near,
                               float dist_far, float* out_xyzi, int* n_out) {
  if (!c || n < 0 || !n_out || leaf < 0.f) return -1;
  const P4* in = reinterpret_cast<const P4*>(xyzi);
  std::vector<P4> cur(in, in + n), nxt;
  if (remove_nan) {
    nxt.clear();
    for (const P4& p : cur)
      if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) nxt.push_back(p);
    cur.swap(nxt);
  }
  if (leaf > 0.f && !cur.empty()) {
    nxt.clear();
    voxel(cur.data(), (int)cur.size(), leaf, nxt, nullptr);
    cur.swap(nxt);
  }
  if (!(dist_near == 0.f && dist_far == 0.f)) {
    nxt.clear();
    for (const P4& p : cur) {
      float s = p.x * p.x + p.y * p.y;
      s = s + p.z * p.z;
      double d = (double)std::sqrt(s);
      if (d > (double)dist_near && d < (double)dist_far) nxt.push_back(p);
    }
    cur.swap(nxt);
  }
  if (out_xyzi && !cur.empty()) std::memcpy(out_xyzi, cur.data(), cur.size() * sizeof(P4));
  *n_out = (int)cur.size();
  return 0;
}

int lmsf_oracle_map_set(lmsf_oracle_ctx* c, int kind, const float* xyzi, int n) {
  if (!c || kind < 0 || kind > 1 || n < 0) return -1;
  if (n == 0) return 0;
  c->map_cat[kind].assign(reinterpret_cast<const P4*>(xyzi), reinterpret_cast<const P4*>(xyzi) + n);
  c->map[kind].assign(c->map_cat[kind].data(), n, c->prm.knn_mode);
  return 0;
}

int lmsf_oracle_knn5(lmsf_oracle_ctx* c, int kind, const float* q, int nq, int32_t* idx5, float* d2_5) {
  if (!c || kind < 0 || kind > 1 || !c->map[kind].set) return -5;
  const int mode = c->prm.knn_mode;
#pragma omp parallel for num_threads(nthreads(c)) schedule(static)
  for (int i = 0; i < nq; ++i) {
    Nb5 nb;
    c->map[kind].knn(q[3 * i], q[3 * i + 1], q[3 * i + 2], mode, nb);
    for (int k = 0; k < 5; ++k) {
      bool in = k < nb.n && nb.d[k] < 1.0f;
      idx5[5 * i + k] = in ? nb.id[k] : -1;
      d2_5[5 * i + k] = in ? nb.d[k] : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

// PointCloudAlignmentEvaluate::AlignmentScore (registration/alignEvaluate.hpp:55-87) against the map `kind`
// (SetTargetPoints :42-46).  pcl::transformPointCloud with a Matrix4f: fp32, ((m0 x + m1 y) + m2 z) + m3
// (PCL's non-SSE dense path); nearestKSearch(point, 1) without a radius; fitness accumulated in double in point
// order.  Sequential on purpose: the accumulation order is the reference's.
int lmsf_oracle_align_score(lmsf_oracle_ctx* c, int kind, const float* xyzi, int n, const float T[16],
                            double inlier_thresh, double inlier_ratio_thresh, double* score, double* overlap,
                            int32_t* n_inlier) {
  if (!c || kind < 0 || kind > 1 || !c->map[kind].set || !score || !overlap) return -5;
  if (n_inlier) *n_inlier = 0;
  if (n == 0) {
    *score = std::numeric_limits<double>::max();
    *overlap = 0;
    return 0;
  }
  const int mode = c->prm.knn_mode;
  double fitness = 0.0;
  int nr = 0;
  for (int i = 0; i < n; ++i) {
    const float x = xyzi[4 * i], y = xyzi[4 * i + 1], z = xyzi[4 * i + 2];
    const float qx = T[0] * x + T[1] * y + T[2] * z + T[3];
    const float qy = T[4] * x + T[5] * y + T[6] * z + T[7];
    const float qz = T[8] * x + T[9] * y + T[10] * z + T[11];
    Nb5 nb;
    c->map[kind].knn(qx, qy, qz, mode, nb);  // the nearest of the five is the 1-NN
    if (nb.n > 0 && (double)nb.d[0] <= inlier_thresh) {
      fitness += (double)nb.d[0];
      nr++;
    }
  }
  const double ratio = (double)nr / (double)n;
  *overlap = ratio;
  *score = (ratio > inlier_ratio_thresh) ? fitness / nr : std::numeric_limits<double>::max();
  if (n_inlier) *n_inlier = nr;
  return 0;
}

int lmsf_oracle_match(lmsf_oracle_ctx* c, int kind, const float* q, int nq, uint8_t* ok, double* out10) {
  if (!c || kind < 0 || kind > 1 || !c->map[kind].set) return -5;
  const int mode = c->prm.knn_mode;
#pragma omp parallel for num_threads(nthreads(c)) schedule(static)
  for (int i = 0; i < nq; ++i) {
    double* o = out10 + 10 * i;
    for (int k = 0; k < 10; ++k) o[k] = 0.0;
    if (kind == 0) {
      EdgeInfo e;
      ok[i] = match_edge(c->map[0], mode, q[3 * i], q[3 * i + 1], q[3 * i + 2], e);
      if (ok[i]) {
        const double v[10] = {e.n.x, e.n.y, e.n.z, e.r, e.a.x, e.a.y, e.a.z, e.b.x, e.b.y, e.b.z};
        std::memcpy(o, v, sizeof v);
      }
    } else {
      SurfInfo s;
      ok[i] = match_surf(c->map[1], mode, q[3 * i], q[3 * i + 1], q[3 * i + 2], s);
      if (ok[i]) {
        const double v[10] = {s.n.x, s.n.y, s.n.z, s.r, s.D, 0, 0, 0, 0, 0};
        std::memcpy(o, v, sizeof v);
      }
    }
  }
  return 0;
}

int lmsf_oracle_register(lmsf_oracle_ctx* c, const float* edge_xyzi, int n_e, const float* surf_xyzi, int n_s,
                         int solver, double pose[7], lmsf_oracle_reg_stats* st) {
  if (!c || !pose || n_e < 0 || n_s < 0) return -1;
  do_register(c, reinterpret_cast<const P4*>(edge_xyzi), n_e, reinterpret_cast<const P4*>(surf_xyzi), n_s, solver, pose,
              st);
  return 0;
}

int lmsf_oracle_set_lm_outer(lmsf_oracle_ctx* c, int count) {
  if (!c || count < 0) return -1;
  c->lm_count = count;
  return 0;
}

int lmsf_oracle_tracker_step_features(lmsf_oracle_ctx* c, const float* edge_xyzi, int n_e, const float* surf_xyzi,
                                      int n_s, double stamp, double delta[7], double pose_out[7],
                                      lmsf_oracle_track_stats* st) {
  if (!c || !delta || !pose_out) return -1;
  Features f;
  f.edge.assign(reinterpret_cast<const P4*>(edge_xyzi), reinterpret_cast<const P4*>(edge_xyzi) + n_e);
  f.surf.assign(reinterpret_cast<const P4*>(surf_xyzi), reinterpret_cast<const P4*>(surf_xyzi) + n_s);
  scan_filter(c, f);
  tracker_solve(c, f, stamp, delta, pose_out, st);
  return 0;
}

int lmsf_oracle_tracker_step(lmsf_oracle_ctx* c, const float* xyzi, int n, double stamp, double delta[7],
                             double pose_out[7], lmsf_oracle_track_stats* st) {
  if (!c || !delta || !pose_out) return -1;
  Features f;
  std::vector<P4> work;
  extract(c->prm, preprocessed(c->prm, reinterpret_cast<const P4*>(xyzi), n, work), n, nullptr, f);
  scan_filter(c, f);
  tracker_solve(c, f, stamp, delta, pose_out, st);
  return 0;
}

int lmsf_oracle_set_threads(lmsf_oracle_ctx* c, int threads) {
  if (!c) return -1;
  c->prm.threads = threads;
  return 0;
}

int lmsf_oracle_tracker_reset(lmsf_oracle_ctx* c) {
  if (!c) return -1;
  c->init = false;
  for (int k = 0; k < 2; ++k) {
    c->win[k].clear();
    c->map_cat[k].clear();
    c->map[k] = MapIndex();
  }
  c->lm_count = c->prm.lm_outer_start;
  return 0;
}

int lmsf_oracle_tracker_register_aux(lmsf_oracle_ctx* c, const float* xyzi, int n, double pose[7],
                                     lmsf_oracle_reg_stats* st) {
  if (!c || !pose) return -1;
  Features f;
  std::vector<P4> work;
  extract(c->prm, preprocessed(c->prm, reinterpret_cast<const P4*>(xyzi), n, work), n, nullptr, f);
  scan_filter(c, f);
  // Solve(Isometry3d&): quaternion -> matrix -> quaternion round trip of the adapter
  Iso T = iso_from_pose7(pose);
  double p[7];
  iso_to_pose7(T, p);
  do_register(c, f.edge.data(), (int)f.edge.size(), f.surf.data(), (int)f.surf.size(), c->prm.solver, p, st);
  T = iso_from_pose7(p);
  iso_to_pose7(T, pose);
  return 0;
}

int lmsf_oracle_get_map(lmsf_oracle_ctx* c, int kind, float* xyzi, int cap, int* n) {
  if (!c || kind < 0 || kind > 1 || !n) return -1;
  int m = (int)c->map_cat[kind].size();
  *n = m;
  if (xyzi) {
    if (cap < m) return -4;
    if (m) std::memcpy(xyzi, c->map_cat[kind].data(), (size_t)m * sizeof(P4));
  }
  return 0;
}

int lmsf_oracle_symeig3(const double a[9], double w[3], double v[9]) {
  symeig<3>(a, w, v);
  return 0;
}
int lmsf_oracle_symeig6(const double a[36], double w[6], double v[36]) {
  symeig<6>(a, w, v);
  return 0;
}
int lmsf_oracle_lstsq53(const double a[15], const double b[5], double x[3]) {
  qr_solve<5, 3>(a, b, x);
  return 0;
}
int lmsf_oracle_solve6(const double a[36], const double b[6], double x[6]) {
  qr_solve<6, 6>(a, b, x);
  return 0;
}
// test hooks: one cost-function evaluation (kind 0: geom = a, b; kind 1: geom = n, D) and the SE3 plus
int lmsf_oracle_factor_eval(int kind, const double x[7], const double p[3], const double geom[7], double* r,
                            double J6[6]) {
  Quat q{x[0], x[1], x[2], x[3]};
  V3 t{x[4], x[5], x[6]};
  V3 pl{p[0], p[1], p[2]};
  if (kind == 0)
    edge_factor(q, t, EdgeBlk{pl, V3{geom[0], geom[1], geom[2]}, V3{geom[3], geom[4], geom[5]}}, true, *r, J6);
  else
    surf_factor(q, t, SurfBlk{pl, V3{geom[0], geom[1], geom[2]}, geom[3]}, true, *r, J6);
  return 0;
}
// test hook: the same trust-region loop over a caller's problem — n_res residual blocks, each giving (r, 1x6 local
// Jacobian) at x through `residual`, the manifold plus through `plus`; Huber loss + corrector and the accumulation of the
// normal equations in block order are the oracle's (accum).  Used to run the reference's own cost functions and
// parameterization inside the restated ceres::Solve (oracle/shim_fixed/ceres/ceres.h).
int lmsf_oracle_lm_solve_cb(int n_res, lmsf_oracle_residual_cb residual, lmsf_oracle_plus_cb plus, void* user,
                            double huber_a, int max_iters, double x[7], int* steps, int* accepted, double* cost) {
  int st = 0, ac = 0;
  double c = 0;
  lm_solve_t(
      [&](const double* xx, bool jac, Normal& N) {
        std::memset(&N, 0, sizeof N);
        double r, J[6];
        for (int i = 0; i < n_res; ++i) {
          for (int k = 0; k < 6; ++k) J[k] = 0;
          residual(user, i, xx, jac ? 1 : 0, &r, J);
          accum(N, J, r, huber_a);
        }
      },
      [&](const double* xx, const double* d, double* out) { plus(user, xx, d, out); }, n_res == 0, max_iters, x, st, ac, c);
  if (steps) *steps = st;
  if (accepted) *accepted = ac;
  if (cost) *cost = c;
  return 0;
}
int lmsf_oracle_se3_plus(const double x[7], const double d[6], double out[7]) {
  se3_plus(x, d, out);
  return 0;
}
int lmsf_oracle_se3_exp(const double d[6], double q[4], double t[3]) {
  Quat qq;
  V3 tt;
  se3_exp(d, qq, tt);
  q[0] = qq.x;
  q[1] = qq.y;
  q[2] = qq.z;
  q[3] = qq.w;
  t[0] = tt.x;
  t[1] = tt.y;
  t[2] = tt.z;
  return 0;
}
int lmsf_oracle_lm_solve(const double* edge9, int n_e, const double* surf7, int n_s, double huber_a, int max_iters,
                         double x[7], int* steps, int* accepted, double* cost) {
  std::vector<EdgeBlk> eb(n_e);
  std::vector<SurfBlk> sb(n_s);
  for (int i = 0; i < n_e; ++i) {
    const double* e = edge9 + 9 * i;
    eb[i] = EdgeBlk{V3{e[0], e[1], e[2]}, V3{e[3], e[4], e[5]}, V3{e[6], e[7], e[8]}};
  }
  for (int i = 0; i < n_s; ++i) {
    const double* s = surf7 + 7 * i;
    sb[i] = SurfBlk{V3{s[0], s[1], s[2]}, V3{s[3], s[4], s[5]}, s[6]};
  }
  int st = 0, ac = 0;
  double c = 0;
  lm_solve(eb, sb, huber_a, max_iters, x, st, ac, c);
  if (steps) *steps = st;
  if (accepted) *accepted = ac;
  if (cost) *cost = c;
  return 0;
}

}  // extern "C"
