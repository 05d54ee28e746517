// lmsf_oracle_sc.cpp — CPU ORACLE of the loop-closure descriptor path ("next" row f1, BASELINE config 5).
//
// TEST INFRASTRUCTURE ONLY (see lmsf_oracle.h).  PINNED against the reference's own code compiled into oracle/_ref
// (tests/test_oracle_sc.py, bit for bit): descriptor, ring key, SC distance and shift against its ScanContext class
// (libref_sc.so; Eigen's mean / norm / dot taken as sequential sums on both sides — the one documented divergence, <= 1 ulp
// of a double), the ring-key search against its KDTreeVectorOfVectorsAdaptor over its vendored nanoflann, and the search
// end to end (tree-rebuild schedule, top-10, selection, threshold, yaw) against its SceneRecognitionScanContext class.
// Restates, with file:line under src/MultiSensorFusionEstimator3D/include/:
//   ScanContext::MakeScanContext / MakeRingkeyFromScanContext / DistanceBtnScanContext / distDirectSC /
//   fastAlignUsingVkey / circshift / xy2theta
//       Algorithm/PointClouds/processing/GlobalDescriptor/scanContext/Scancontext.hpp:59-318
//   SceneRecognitionScanContext::descFindSimilar  LoopDetection/SceneRecognitionScanContext.hpp:260-333
//   (ring-key kNN k = 10 by nanoflann metric_L2 = L2_Adaptor, nanoflann.hpp:375-414: fp32, groups of four).
// Canonical tie-break where nanoflann's is traversal dependent: (distance, id).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "lmsf_oracle.h"

namespace {
const int NR = 20, NS = 60;
const double MAX_RADIUS = 80.0, LIDAR_HEIGHT = 2.0;

// xy2theta (Scancontext.hpp:304-314); atan on the float quotient is std::atan(float) — `using namespace std;`
// (src/apps/include/utility.hpp:51) precedes this header in the node's translation unit, as for the feature extractor
// (see ring_of in lmsf_oracle.cpp) — its float result widened for the product with 180 / M_PI, returned as float
float xy2theta(float x, float y) {
  if (x >= 0 && y >= 0) return (float)((180 / M_PI) * std::atan(y / x));
  if (x < 0 && y >= 0) return (float)(180 - ((180 / M_PI) * std::atan(y / (-x))));
  if (x < 0 && y < 0) return (float)(180 + ((180 / M_PI) * std::atan(y / x)));
  if (x >= 0 && y < 0) return (float)(360 - ((180 / M_PI) * std::atan((-y) / x)));
  return std::numeric_limits<float>::quiet_NaN();
}

// int(ceil(v)) as x86 cvttsd2si gives it for NaN / out of range: INT_MIN
int ceil_to_int(double v) {
  double c = std::ceil(v);
  if (!(c > -2147483648.0 && c < 2147483648.0)) return std::numeric_limits<int>::min();
  return (int)c;
}

// MakeScanContext (:59-104): desc[ring][sector] = max (z + 2.0) as float, 0 where empty
void make_sc(const float* xyzi, int n, float* desc) {
  const float NO_POINT = -1000.f;
  for (int i = 0; i < NR * NS; ++i) desc[i] = NO_POINT;
  for (int i = 0; i < n; ++i) {
    float x = xyzi[4 * i], y = xyzi[4 * i + 1];
    float z = (float)((double)xyzi[4 * i + 2] + LIDAR_HEIGHT);
    float azim_range = std::sqrt(x * x + y * y);  // std::sqrt(float) (:78)
    float azim_angle = xy2theta(x, y);
    if (azim_range > MAX_RADIUS) continue;
    int ring = std::max(std::min(NR, ceil_to_int(((double)azim_range / MAX_RADIUS) * NR)), 1);
    int sector = std::max(std::min(NS, ceil_to_int(((double)azim_angle / 360.0) * NS)), 1);
    float& d = desc[(ring - 1) * NS + (sector - 1)];
    if ((double)d < (double)z) d = z;
  }
  for (int i = 0; i < NR * NS; ++i)
    if (desc[i] == NO_POINT) desc[i] = 0.f;
}

// MakeRingkeyFromScanContext (:112-126): row means (double), stored as float (eig2vec)
void ring_key(const float* desc, float* key) {
  for (int r = 0; r < NR; ++r) {
    double s = 0;
    for (int c = 0; c < NS; ++c) s += (double)desc[r * NS + c];
    key[r] = (float)(s / NS);
  }
}

void sector_key(const float* desc, double* vk) {
  for (int c = 0; c < NS; ++c) {
    double s = 0;
    for (int r = 0; r < NR; ++r) s += (double)desc[r * NS + c];
    vk[c] = s / NR;
  }
}

// distDirectSC (:213-232) of sc1 against sc2 circularly shifted by `shift` columns
double dist_direct(const float* a, const float* b, int shift) {
  int eff = 0;
  double sum = 0;
  for (int c = 0; c < NS; ++c) {
    int cb = (c - shift + NS) % NS;  // circshift: column cb of b lands on column c
    double na = 0, nb = 0, dot = 0;
    for (int r = 0; r < NR; ++r) {
      double va = a[r * NS + c], vb = b[r * NS + cb];
      na += va * va;
      nb += vb * vb;
      dot += va * vb;
    }
    na = std::sqrt(na);
    nb = std::sqrt(nb);
    if (na == 0 || nb == 0) continue;
    sum = sum + dot / (na * nb);
    eff++;
  }
  return 1.0 - sum / eff;
}

// DistanceBtnScanContext (:133-172)
void sc_distance(const float* a, const float* b, double* dist, int* shift_out) {
  double v1[NS], v2[NS];
  sector_key(a, v1);
  sector_key(b, v2);
  int best = 0;
  double bestn = 10000000;
  for (int s = 0; s < NS; ++s) {  // fastAlignUsingVkey (:243-263)
    double q = 0;
    for (int c = 0; c < NS; ++c) {
      double d = v1[c] - v2[(c - s + NS) % NS];
      q += d * d;
    }
    double nrm = std::sqrt(q);
    if (nrm < bestn) {
      best = s;
      bestn = nrm;
    }
  }
  const int radius = (int)std::round(0.5 * 0.1 * NS);
  std::vector<int> space{best};
  for (int i = 1; i < radius + 1; ++i) {
    space.push_back((best + i + NS) % NS);
    space.push_back((best - i + NS) % NS);
  }
  std::sort(space.begin(), space.end());
  int arg = 0;
  double mind = 10000000;
  for (int s : space) {
    double d = dist_direct(a, b, s);
    if (d < mind) {
      arg = s;
      mind = d;
    }
  }
  *dist = mind;
  *shift_out = arg;
}

// nanoflann L2_Adaptor::evalMetric (nanoflann.hpp:383-408) without the early exit: fp32, groups of four
float key_dist(const float* a, const float* b) {
  float result = 0.f;
  int d = 0;
  for (; d + 3 < NR; d += 4) {
    float d0 = a[d] - b[d], d1 = a[d + 1] - b[d + 1], d2 = a[d + 2] - b[d + 2], d3 = a[d + 3] - b[d + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; d < NR; ++d) {
    float d0 = a[d] - b[d];
    result += d0 * d0;
  }
  return result;
}
}  // namespace

extern "C" {

int lmsf_oracle_sc_make(const float* xyzi, int n, float* desc1200, float* key20) {
  if (n < 0 || !desc1200 || !key20) return -1;
  make_sc(xyzi, n, desc1200);
  ring_key(desc1200, key20);
  return 0;
}

int lmsf_oracle_sc_distance(const float* desc_a, const float* desc_b, double* dist, int* shift) {
  sc_distance(desc_a, desc_b, dist, shift);
  return 0;
}

// ring-key 10-NN among keys[0, limit) — ascending by (distance, id)
int lmsf_oracle_sc_knn(const float* keys, int limit, const float* q_keys, int nq, int32_t* idx10, float* d10) {
  for (int q = 0; q < nq; ++q) {
    std::vector<std::pair<float, int>> all(limit);
    for (int i = 0; i < limit; ++i) all[i] = {key_dist(q_keys + 20 * q, keys + 20 * (size_t)i), i};
    int k = std::min(10, limit);
    std::partial_sort(all.begin(), all.begin() + k, all.end());
    for (int j = 0; j < 10; ++j) {
      idx10[10 * q + j] = j < k ? all[j].second : -1;
      d10[10 * q + j] = j < k ? all[j].first : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

// descFindSimilar (SceneRecognitionScanContext.hpp:260-333): loop id (or -1), its SC distance and column shift
int lmsf_oracle_sc_search(const float* keys, const float* descs, int limit, const float* q_keys, const float* q_descs,
                          int nq, double thresh, int32_t* loop_id, double* loop_dist, int32_t* loop_shift) {
  if (limit < 1) return -1;  // fewer than ten keys: nanoflann fills only `limit` slots; the rest are skipped here
  std::vector<int32_t> idx(10 * (size_t)nq);
  std::vector<float> d(10 * (size_t)nq);
  lmsf_oracle_sc_knn(keys, limit, q_keys, nq, idx.data(), d.data());
  for (int q = 0; q < nq; ++q) {
    double mind = 10000000;
    int align = 0, nn = 0;
    for (int j = 0; j < 10; ++j) {
      int id = idx[10 * q + j];
      if (id < 0) continue;
      double dist;
      int sh;
      sc_distance(q_descs + 1200 * (size_t)q, descs + 1200 * (size_t)id, &dist, &sh);
      if (dist < mind) {
        mind = dist;
        align = sh;
        nn = id;
      }
    }
    loop_dist[q] = mind;
    loop_shift[q] = align;
    loop_id[q] = (mind < thresh) ? nn : -1;
  }
  return 0;
}
}
