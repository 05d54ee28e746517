// scancontext.cu — loop-closure descriptor path (BASELINE config 5) on the device.
//
// Replaces, for batches of queries and a database that lives in HBM (optionally one shard of it per GPU):
//   ScanContext::MakeScanContext / MakeRingkeyFromScanContext
//       Algorithm/PointClouds/processing/GlobalDescriptor/scanContext/Scancontext.hpp:59-104, :112-126
//   ScanContext::DistanceBtnScanContext (+ makeSectorkeyFromScanContext, fastAlignUsingVkey, distDirectSC, circshift)
//       Scancontext.hpp:133-172, :191-318
//   SceneRecognitionScanContext::AddKeyFramePoints / descFindSimilar  LoopDetection/SceneRecognitionScanContext.hpp:61-94, :260-333
//       (nanoflann ring-key tree, k = 10, metric_L2 = L2_Adaptor nanoflann.hpp:375-414 — replaced by an exact brute-force scan)
//
// Layout: a descriptor is 20 x 60 fp32, row-major (ring, sector) — the reference keeps MatrixXd, but every entry is a
// float (pt.z is a float), so fp32 storage is lossless; a ring key is 20 fp32 (eig2vec).  The database is two dense
// arrays keys[cap][20] and descs[cap][1200].
//
// Arithmetic: the operation sequence of the reference, expression by expression (fp32 products for the range, double
// for the bin indices, fp32 groups-of-four for the ring-key distance, double for everything in the SC distance), with
// sequential sums where Eigen uses a vectorised reduction (documented in DESIGN.md).  Built with -fmad=false.
#include <math.h>

#include "common.cuh"
#include "fmath.cuh"

namespace lm {

constexpr int SC_NR = 20, SC_NS = 60, SC_CELLS = SC_NR * SC_NS, SC_K = 10;
constexpr float SC_NO_POINT = -1000.0f;

// one ring-key candidate of one query as exchanged between database shards (24 bytes)
struct __align__(8) ScCand {
  double sc_dist;   // DistanceBtnScanContext of (query, candidate)
  float key_dist;   // squared ring-key distance
  int id;           // global keyframe id, -1 = empty slot
  int shift;        // column shift of the best alignment
  int pad;
};

static_assert(sizeof(ScCand) == sizeof(lmsf_sc_cand) && sizeof(ScCand) == 24, "candidate record layout");

struct ScDb {
  int cap = 0, n = 0;
  float* keys = nullptr;   // [cap][20]
  float* descs = nullptr;  // [cap][1200]
  int* bins = nullptr;     // [1200] ordered-int running maxima of the cloud being described
  // query scratch, grown on demand
  int q_cap = 0;
  float* q_keys = nullptr;   // [q_cap][20]
  float* q_descs = nullptr;  // [q_cap][1200]
  int p_cap = 0;             // partial lists per query the scratch below holds
  float* part_d = nullptr;   // [p_cap][q_cap][10]
  int* part_i = nullptr;
  float* top_d = nullptr;    // [q_cap][10]
  int* top_i = nullptr;      // [q_cap][10] local ids
  float* tau = nullptr;      // [q_cap] per-query bound of the 10th ring-key distance (threshold pass)
  ScCand* cand = nullptr;    // [q_cap][10]
  int* out_id = nullptr;     // [q_cap]
  double* out_dist = nullptr;
  int* out_shift = nullptr;
  float4* cloud = nullptr;   // staging for host clouds
  int cloud_cap = 0;
};

static ScDb* scdb_of(Ctx* c) {
  if (!c->scdb) c->scdb = new ScDb();
  return (ScDb*)c->scdb;
}

void scdb_free(Ctx* c) {
  ScDb* d = (ScDb*)c->scdb;
  if (!d) return;
  cudaFree(d->keys);
  cudaFree(d->descs);
  cudaFree(d->bins);
  cudaFree(d->q_keys);
  cudaFree(d->q_descs);
  cudaFree(d->part_d);
  cudaFree(d->part_i);
  cudaFree(d->top_d);
  cudaFree(d->top_i);
  cudaFree(d->tau);
  cudaFree(d->cand);
  cudaFree(d->out_id);
  cudaFree(d->out_dist);
  cudaFree(d->out_shift);
  cudaFree(d->cloud);
  delete d;
  c->scdb = nullptr;
}

// ------------------------------------------------------------------ MakeScanContext
// order-preserving float -> int (signed compare); NaN never reaches it
__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// int(ceil(v)) as x86 cvttsd2si: INT_MIN for NaN and out-of-range values
__device__ __forceinline__ int ceil_to_int(double v) {
  double c = ceil(v);
  if (!(c > -2147483648.0 && c < 2147483648.0)) return (int)0x80000000;
  return (int)c;
}

// xy2theta (Scancontext.hpp:304-318): atan of the float quotient is std::atan(float) = the C library's atanf in the
// reference's translation unit (`using namespace std;`, utility.hpp:51; fmath.cuh returns its bits), widened for the
// product with 180 / M_PI, returned as float
__device__ __forceinline__ float xy2theta(float x, float y) {
  const double k = 180 / M_PI;
  if (x >= 0 && y >= 0) return (float)(k * (double)atanf_fdlibm(y / x));
  if (x < 0 && y >= 0) return (float)(180 - (k * (double)atanf_fdlibm(y / (-x))));
  if (x < 0 && y < 0) return (float)(180 + (k * (double)atanf_fdlibm(y / x)));
  if (x >= 0 && y < 0) return (float)(360 - (k * (double)atanf_fdlibm((-y) / x)));
  return __int_as_float(0x7fc00000);
}

__global__ void __launch_bounds__(256) k_sc_clear(int* __restrict__ bins) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < SC_CELLS) bins[i] = f2ord(SC_NO_POINT);
}

// per-point polar bin + running maximum of z + LIDAR_HEIGHT (:69-93): shared-memory maxima per block, then one
// global atomicMax per touched bin.  max is order independent, so the result equals the sequential loop's.
__global__ void __launch_bounds__(256) k_sc_bin(const float4* __restrict__ pts, int n, int* __restrict__ bins) {
  __shared__ int s_bins[SC_CELLS];
  const int init = f2ord(SC_NO_POINT);
  for (int i = threadIdx.x; i < SC_CELLS; i += blockDim.x) s_bins[i] = init;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    float x = p.x, y = p.y;
    float z = (float)((double)p.z + 2.0);
    float xx = x * x, yy = y * y;
    float rng = (float)sqrt((double)(xx + yy));  // == std::sqrt(float) (:78): rounding a double sqrt to float is exact
    float ang = xy2theta(x, y);
    if ((double)rng > 80.0) continue;
    int ring = max(min(SC_NR, ceil_to_int(((double)rng / 80.0) * SC_NR)), 1);
    int sector = max(min(SC_NS, ceil_to_int(((double)ang / 360.0) * SC_NS)), 1);
    if (!(z > SC_NO_POINT)) continue;  // desc < z is false for NaN and for anything <= the initial value
    atomicMax(&s_bins[(ring - 1) * SC_NS + (sector - 1)], f2ord(z));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SC_CELLS; i += blockDim.x)
    if (s_bins[i] != init) atomicMax(&bins[i], s_bins[i]);
}

// NO_POINT -> 0 (:96-104), then the ring key = row means (:112-126; sum in double, stored as float by eig2vec)
__global__ void __launch_bounds__(256) k_sc_finish(const int* __restrict__ bins, float* __restrict__ desc,
                                                   float* __restrict__ key) {
  __shared__ float s_d[SC_CELLS];
  for (int i = threadIdx.x; i < SC_CELLS; i += blockDim.x) {
    float v = ord2f(bins[i]);
    if (v == SC_NO_POINT) v = 0.0f;
    s_d[i] = v;
    desc[i] = v;
  }
  __syncthreads();
  if (threadIdx.x < SC_NR) {
    double s = 0;
    for (int cidx = 0; cidx < SC_NS; ++cidx) s += (double)s_d[threadIdx.x * SC_NS + cidx];
    key[threadIdx.x] = (float)(s / SC_NS);
  }
}

// ------------------------------------------------------------------ ring-key 10-NN (exact scan)
struct Top10 {
  float d[SC_K];
  int id[SC_K];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int k = 0; k < SC_K; ++k) {
      d[k] = __int_as_float(0x7f800000);
      id[k] = 0x7fffffff;
    }
  }
  // ten keys are known to exist with distance <= bound: accept distance <= bound, ties included
  __device__ __forceinline__ void reset_inclusive(float bound) {
#pragma unroll
    for (int k = 0; k < SC_K; ++k) {
      d[k] = bound;
      id[k] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ void add(float dd, int ii) {
    if (dd < d[SC_K - 1] || (dd == d[SC_K - 1] && ii < id[SC_K - 1])) {
      d[SC_K - 1] = dd;
      id[SC_K - 1] = ii;
#pragma unroll
      for (int k = SC_K - 1; k > 0; --k) {
        bool sw = d[k] < d[k - 1] || (d[k] == d[k - 1] && id[k] < id[k - 1]);
        if (sw) {
          float td = d[k];
          d[k] = d[k - 1];
          d[k - 1] = td;
          int ti = id[k];
          id[k] = id[k - 1];
          id[k - 1] = ti;
        }
      }
    }
  }
};

constexpr int KQ = 32;       // queries per block (one per lane)
constexpr int KW = 8;        // warps per block: each scans every 8th key of a chunk
constexpr int KCHUNK = 256;  // keys staged in shared memory per iteration (20 KB)

// L2_Adaptor::evalMetric (nanoflann.hpp:383-408): fp32, four differences at a time
__device__ __forceinline__ float key_dist(const float (&q)[SC_NR], const float* __restrict__ k) {
  float result = 0.0f;
#pragma unroll
  for (int d = 0; d < SC_NR; d += 4) {
    float4 v = *reinterpret_cast<const float4*>(k + d);
    float d0 = q[d] - v.x, d1 = q[d + 1] - v.y, d2 = q[d + 2] - v.z, d3 = q[d + 3] - v.w;
    float t = d0 * d0;
    t = t + d1 * d1;
    t = t + d2 * d2;
    t = t + d3 * d3;
    result = result + t;
  }
  return result;
}

// grid (query tiles, parts): block (tile, part) scans keys [part*per, min(limit, (part+1)*per)) for 32 queries and
// writes their partial top-10 lists (ascending by (distance, id)).
// tau (optional): per query an upper bound of its 10th smallest distance (the 10th of a sample of the same keys).
// Every (warp, part) list is independent, and a list that starts empty inserts on most of its first few hundred
// keys — with ~150 lists per query the sorted insert, not the distance, dominated the scan (83 % of the keys took
// the insert path in some lane).  Starting every list at tau leaves only the ~0.2 % of keys that can still matter.
__global__ void __launch_bounds__(KQ* KW) k_sc_scan(const float* __restrict__ keys, int limit, int per,
                                                     const float* __restrict__ q_keys, int nq,
                                                     const float* __restrict__ tau,
                                                     float* __restrict__ part_d, int* __restrict__ part_i) {
  __shared__ __align__(16) float s_keys[KCHUNK * SC_NR];
  __shared__ float s_d[KW][SC_K][KQ];
  __shared__ int s_i[KW][SC_K][KQ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * KQ + lane;
  const int lo = blockIdx.y * per, hi = min(limit, lo + per);
  float qk[SC_NR];
#pragma unroll
  for (int d = 0; d < SC_NR; ++d) qk[d] = (q < nq) ? q_keys[(size_t)q * SC_NR + d] : 0.0f;
  Top10 top;
  if (tau != nullptr && q < nq)
    top.reset_inclusive(tau[q]);
  else
    top.reset();
  for (int base = lo; base < hi; base += KCHUNK) {
    const int cnt = min(KCHUNK, hi - base);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(keys + (size_t)base * SC_NR);
    float4* dst = reinterpret_cast<float4*>(s_keys);
    for (int i = threadIdx.x; i < cnt * (SC_NR / 4); i += blockDim.x) dst[i] = __ldg(&src[i]);
    __syncthreads();
    for (int k = warp; k < cnt; k += KW) top.add(key_dist(qk, s_keys + k * SC_NR), base + k);
  }
#pragma unroll
  for (int k = 0; k < SC_K; ++k) {
    s_d[warp][k][lane] = top.d[k];
    s_i[warp][k][lane] = top.id[k];
  }
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < KW; ++w)
#pragma unroll
      for (int k = 0; k < SC_K; ++k) top.add(s_d[w][k][lane], s_i[w][k][lane]);
    if (q < nq) {
      size_t o = ((size_t)blockIdx.y * nq + q) * SC_K;
#pragma unroll
      for (int k = 0; k < SC_K; ++k) {
        part_d[o + k] = top.d[k];
        part_i[o + k] = top.id[k];
      }
    }
  }
}

// merge the per-part lists of every query: ascending (distance, id); empty slots -> id -1, distance +inf.
// One warp per query: every lane folds the entries lane, lane + 32, ... of the parts x 10 candidates into a sorted
// list of its own, then ten rounds of a warp-wide lexicographic minimum over the list heads pop the result.
// tau_out (optional): the 10th distance of the merged list (+inf when fewer than ten), for the threshold pass.
__global__ void __launch_bounds__(128) k_sc_merge_parts(const float* __restrict__ part_d, const int* __restrict__ part_i,
                                                        int parts, int nq, float* __restrict__ top_d,
                                                        int* __restrict__ top_i, float* __restrict__ tau_out) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  Top10 top;
  top.reset();
  const int total = parts * SC_K;
  for (int e = lane; e < total; e += 32) {
    const int p = e / SC_K, k = e - p * SC_K;
    const size_t o = ((size_t)p * nq + q) * SC_K + k;
    const int id = part_i[o];
    if (id != 0x7fffffff) top.add(part_d[o], id);
  }
  float out_d = __int_as_float(0x7f800000);
  int out_i = -1;
  float tenth = __int_as_float(0x7f800000);
#pragma unroll
  for (int r = 0; r < SC_K; ++r) {
    float d = top.d[0];
    int id = top.id[0];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, s);
      const int oi = __shfl_xor_sync(0xffffffffu, id, s);
      if (od < d || (od == d && oi < id)) {
        d = od;
        id = oi;
      }
    }
    // the lane that owns the minimum pops its head (ids are unique across lanes; placeholders never win over a real
    // entry and, once only placeholders are left, every lane pops — harmless)
    if (top.id[0] == id && top.d[0] == d) {
#pragma unroll
      for (int k = 0; k + 1 < SC_K; ++k) {
        top.d[k] = top.d[k + 1];
        top.id[k] = top.id[k + 1];
      }
      top.d[SC_K - 1] = __int_as_float(0x7f800000);
      top.id[SC_K - 1] = 0x7fffffff;
    }
    const bool real = id != 0x7fffffff;
    if (lane == r) {
      out_d = real ? d : __int_as_float(0x7f800000);
      out_i = real ? id : -1;
    }
    if (r == SC_K - 1 && real) tenth = d;
  }
  if (lane < SC_K) {
    top_d[(size_t)q * SC_K + lane] = out_d;
    top_i[(size_t)q * SC_K + lane] = out_i;
  }
  if (tau_out != nullptr && lane == 0) tau_out[q] = tenth;
}

// ------------------------------------------------------------------ DistanceBtnScanContext
// One block of 64 threads per (query, candidate) pair.  a = the query descriptor (_sc1), b = the candidate (_sc2).
__device__ void sc_distance_block(const float* __restrict__ a_g, const float* __restrict__ b_g, double* dist_out,
                                  int* shift_out) {
  __shared__ float a[SC_CELLS], b[SC_CELLS];
  __shared__ double v1[SC_NS], v2[SC_NS], na[SC_NS], nb[SC_NS], nrm[SC_NS];
  __shared__ double sim[7][SC_NS];
  __shared__ double dsh[7];
  __shared__ int space[7];
  const int t = threadIdx.x;
  for (int i = t; i < SC_CELLS; i += blockDim.x) {
    a[i] = a_g[i];
    b[i] = b_g[i];
  }
  __syncthreads();
  if (t < SC_NS) {
    // makeSectorkeyFromScanContext (:191-203): column means; column norms for distDirectSC (:213-232)
    double s1 = 0, s2 = 0, q1 = 0, q2 = 0;
    for (int r = 0; r < SC_NR; ++r) {
      double x = a[r * SC_NS + t], y = b[r * SC_NS + t];
      s1 += x;
      s2 += y;
      q1 += x * x;
      q2 += y * y;
    }
    v1[t] = s1 / SC_NR;
    v2[t] = s2 / SC_NR;
    na[t] = sqrt(q1);
    nb[t] = sqrt(q2);
  }
  __syncthreads();
  if (t < SC_NS) {
    // fastAlignUsingVkey (:243-263): norm of vkey1 - circshift(vkey2, t)
    double q = 0;
    for (int cidx = 0; cidx < SC_NS; ++cidx) {
      double d = v1[cidx] - v2[(cidx - t + SC_NS) % SC_NS];
      q += d * d;
    }
    nrm[t] = sqrt(q);
  }
  __syncthreads();
  if (t == 0) {
    int best = 0;
    double bestn = 10000000;
    for (int s = 0; s < SC_NS; ++s)
      if (nrm[s] < bestn) {
        best = s;
        bestn = nrm[s];
      }
    // SEARCH_RADIUS = round(0.5 * 0.1 * 60) = 3 (:141); the search space sorted ascending (:155)
    int sp[7];
    sp[0] = best;
    for (int i = 1; i <= 3; ++i) {
      sp[2 * i - 1] = (best + i + SC_NS) % SC_NS;
      sp[2 * i] = (best - i + SC_NS) % SC_NS;
    }
    for (int i = 1; i < 7; ++i) {
      int v = sp[i], j = i - 1;
      while (j >= 0 && sp[j] > v) {
        sp[j + 1] = sp[j];
        --j;
      }
      sp[j + 1] = v;
    }
    for (int i = 0; i < 7; ++i) space[i] = sp[i];
  }
  __syncthreads();
  if (t < SC_NS) {
    for (int j = 0; j < 7; ++j) {
      int cb = (t - space[j] + SC_NS) % SC_NS;  // circshift (:284-302): column cb of b lands on column t
      double dot = 0;
      for (int r = 0; r < SC_NR; ++r) dot += (double)a[r * SC_NS + t] * (double)b[r * SC_NS + cb];
      // NaN marks "skip this sector pair" (:222-223)
      sim[j][t] = (na[t] == 0 || nb[cb] == 0) ? __longlong_as_double(0x7ff8000000000001ll) : dot / (na[t] * nb[cb]);
    }
  }
  __syncthreads();
  if (t < 7) {
    int eff = 0;
    double sum = 0;
    for (int cidx = 0; cidx < SC_NS; ++cidx) {
      int cb = (cidx - space[t] + SC_NS) % SC_NS;
      if (na[cidx] == 0 || nb[cb] == 0) continue;
      sum = sum + sim[t][cidx];
      eff++;
    }
    dsh[t] = 1.0 - sum / eff;
  }
  __syncthreads();
  if (t == 0) {
    int arg = 0;
    double mind = 10000000;
    for (int j = 0; j < 7; ++j)
      if (dsh[j] < mind) {
        arg = space[j];
        mind = dsh[j];
      }
    *dist_out = mind;
    *shift_out = arg;
  }
}

// explicit pairs (test hook): a[p], b[p] are 1200-float descriptors
__global__ void __launch_bounds__(64) k_sc_dist_pairs(const float* __restrict__ a, const float* __restrict__ b,
                                                      double* __restrict__ dist, int* __restrict__ shift) {
  size_t p = blockIdx.x;
  sc_distance_block(a + p * SC_CELLS, b + p * SC_CELLS, &dist[p], &shift[p]);
}

// candidates of a query batch against this shard: block (k, q) fills cand[q][k]
__global__ void __launch_bounds__(64) k_sc_dist_cand(const float* __restrict__ q_descs, const float* __restrict__ descs,
                                                     const float* __restrict__ top_d, const int* __restrict__ top_i,
                                                     int id_base, ScCand* __restrict__ cand) {
  const size_t o = (size_t)blockIdx.y * SC_K + blockIdx.x;
  const int id = top_i[o];
  __shared__ double d;
  __shared__ int sh;
  if (id >= 0) sc_distance_block(q_descs + (size_t)blockIdx.y * SC_CELLS, descs + (size_t)id * SC_CELLS, &d, &sh);
  __syncthreads();
  if (threadIdx.x == 0) {
    ScCand r;
    r.sc_dist = id >= 0 ? d : 10000000.0;
    r.key_dist = top_d[o];
    r.id = id >= 0 ? id + id_base : -1;
    r.shift = id >= 0 ? sh : 0;
    r.pad = 0;
    cand[o] = r;
  }
}

// ---- two-round exchange (sharded search): round 1 gathers every shard's ring-key top-10 WITHOUT ScanContext distances,
// every rank derives the same global top-10 and scores only the candidates it owns (about 10 / n_ranks per query
// instead of 10), round 2 gathers the scored records and k_sc_pick selects as before.
// the shard's ring-key top-10 as unscored candidate records
__global__ void __launch_bounds__(256) k_sc_keys_to_cand(const float* __restrict__ top_d, const int* __restrict__ top_i,
                                                         int n, int id_base, ScCand* __restrict__ cand) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const int id = top_i[o];
  ScCand r;
  r.sc_dist = 10000000.0;
  r.key_dist = top_d[o];
  r.id = id >= 0 ? id + id_base : -1;
  r.shift = 0;
  r.pad = 0;
  cand[o] = r;
}
// the global ring-key top-10 by (key distance, id) of every query out of n_ranks x 10 unscored records -> sel[q][10]
__global__ void __launch_bounds__(128) k_sc_select_global(const ScCand* __restrict__ all, int n_ranks, int nq,
                                                          ScCand* __restrict__ sel) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float kd[SC_K];
  int kid[SC_K];
#pragma unroll
  for (int k = 0; k < SC_K; ++k) {
    kd[k] = __int_as_float(0x7f800000);
    kid[k] = 0x7fffffff;
  }
  for (int r = 0; r < n_ranks; ++r)
    for (int j = 0; j < SC_K; ++j) {
      const ScCand& cnd = all[(r * nq + q) * SC_K + j];
      if (cnd.id < 0) continue;
      float dd = cnd.key_dist;
      int ii = cnd.id;
      if (dd < kd[SC_K - 1] || (dd == kd[SC_K - 1] && ii < kid[SC_K - 1])) {
        kd[SC_K - 1] = dd;
        kid[SC_K - 1] = ii;
#pragma unroll
        for (int k = SC_K - 1; k > 0; --k) {
          bool sw = kd[k] < kd[k - 1] || (kd[k] == kd[k - 1] && kid[k] < kid[k - 1]);
          if (sw) {
            float td = kd[k];
            kd[k] = kd[k - 1];
            kd[k - 1] = td;
            int ti = kid[k];
            kid[k] = kid[k - 1];
            kid[k - 1] = ti;
          }
        }
      }
    }
#pragma unroll
  for (int k = 0; k < SC_K; ++k) {
    ScCand r;
    r.sc_dist = 10000000.0;
    r.key_dist = kd[k];
    r.id = kid[k] == 0x7fffffff ? -1 : kid[k];
    r.shift = 0;
    r.pad = 0;
    sel[(size_t)q * SC_K + k] = r;
  }
}
// block (k, q): the ScanContext distance of sel[q][k] if this shard owns it (global ids [id_base, id_base + n_local));
// a candidate of another shard is written as an empty slot — its owner reports it
__global__ void __launch_bounds__(64) k_sc_dist_owned(const float* __restrict__ q_descs, const float* __restrict__ descs,
                                                      const ScCand* __restrict__ sel, int id_base, int n_local,
                                                      ScCand* __restrict__ out) {
  const size_t o = (size_t)blockIdx.y * SC_K + blockIdx.x;
  ScCand r = sel[o];
  const int local = r.id - id_base;
  const bool mine = r.id >= 0 && local >= 0 && local < n_local;
  __shared__ double d;
  __shared__ int sh;
  if (mine) sc_distance_block(q_descs + (size_t)blockIdx.y * SC_CELLS, descs + (size_t)local * SC_CELLS, &d, &sh);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mine) {
      r.sc_dist = d;
      r.shift = sh;
    } else {
      r.id = -1;
    }
    out[o] = r;
  }
}

// descFindSimilar's selection (:296-323) over the candidates of `n_ranks` shards: the global ring-key top-10 by
// (key distance, id) first, then the first strict minimum of the SC distance in that order, then the threshold.
__global__ void __launch_bounds__(128) k_sc_pick(const ScCand* __restrict__ all, int n_ranks, int nq, double thresh,
                                                 int* __restrict__ loop_id, double* __restrict__ loop_dist,
                                                 int* __restrict__ loop_shift) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  // selection of the ten smallest (key_dist, id): positions into `all`, kept sorted
  float kd[SC_K];
  int kid[SC_K], pos[SC_K];
#pragma unroll
  for (int k = 0; k < SC_K; ++k) {
    kd[k] = __int_as_float(0x7f800000);
    kid[k] = 0x7fffffff;
    pos[k] = -1;
  }
  for (int r = 0; r < n_ranks; ++r)
    for (int j = 0; j < SC_K; ++j) {
      int p = (r * nq + q) * SC_K + j;
      const ScCand& cnd = all[p];
      if (cnd.id < 0) continue;
      float dd = cnd.key_dist;
      int ii = cnd.id;
      if (dd < kd[SC_K - 1] || (dd == kd[SC_K - 1] && ii < kid[SC_K - 1])) {
        kd[SC_K - 1] = dd;
        kid[SC_K - 1] = ii;
        pos[SC_K - 1] = p;
#pragma unroll
        for (int k = SC_K - 1; k > 0; --k) {
          bool sw = kd[k] < kd[k - 1] || (kd[k] == kd[k - 1] && kid[k] < kid[k - 1]);
          if (sw) {
            float td = kd[k];
            kd[k] = kd[k - 1];
            kd[k - 1] = td;
            int ti = kid[k];
            kid[k] = kid[k - 1];
            kid[k - 1] = ti;
            int tp = pos[k];
            pos[k] = pos[k - 1];
            pos[k - 1] = tp;
          }
        }
      }
    }
  double mind = 10000000;
  int align = 0, nn = 0;
#pragma unroll
  for (int k = 0; k < SC_K; ++k) {
    if (pos[k] < 0) continue;
    const ScCand& cnd = all[pos[k]];
    if (cnd.sc_dist < mind) {
      mind = cnd.sc_dist;
      align = cnd.shift;
      nn = cnd.id;
    }
  }
  loop_dist[q] = mind;
  loop_shift[q] = align;
  loop_id[q] = (mind < thresh) ? nn : -1;
}

// ------------------------------------------------------------------ host side
static int ensure_db(Ctx* c, ScDb* d, int cap) {
  if (cap <= d->cap) return LMSF_OK;
  int ncap = d->cap ? d->cap : 1024;
  while (ncap < cap) ncap *= 2;
  float *nk = nullptr, *nd = nullptr;
  LM_CUDA(cudaMalloc(&nk, (size_t)ncap * SC_NR * sizeof(float)));
  LM_CUDA(cudaMalloc(&nd, (size_t)ncap * SC_CELLS * sizeof(float)));
  if (d->n > 0) {
    LM_CUDA(cudaMemcpyAsync(nk, d->keys, (size_t)d->n * SC_NR * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    LM_CUDA(cudaMemcpyAsync(nd, d->descs, (size_t)d->n * SC_CELLS * sizeof(float), cudaMemcpyDeviceToDevice,
                            c->stream));
    LM_CUDA(cudaStreamSynchronize(c->stream));
  }
  cudaFree(d->keys);
  cudaFree(d->descs);
  d->keys = nk;
  d->descs = nd;
  d->cap = ncap;
  if (!d->bins) LM_CUDA(cudaMalloc(&d->bins, SC_CELLS * sizeof(int)));
  return LMSF_OK;
}

static int scan_parts(int nq, int limit) {
  int tiles = div_up(nq, KQ);
  int parts = div_up(148 * 4, tiles);
  int max_parts = div_up(limit, KCHUNK);  // at least one chunk of keys per part
  if (parts > max_parts) parts = max_parts;
  if (parts > 64) parts = 64;
  if (parts < 1) parts = 1;
  return parts;
}

static int ensure_query(Ctx* c, ScDb* d, int nq, int parts) {
  if (nq > d->q_cap || parts > d->p_cap) {
    int qc = d->q_cap > nq ? d->q_cap : nq;
    int pc = d->p_cap > parts ? d->p_cap : parts;
    LM_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d->q_keys);
    cudaFree(d->q_descs);
    cudaFree(d->part_d);
    cudaFree(d->part_i);
    cudaFree(d->top_d);
    cudaFree(d->top_i);
    cudaFree(d->tau);
    cudaFree(d->cand);
    cudaFree(d->out_id);
    cudaFree(d->out_dist);
    cudaFree(d->out_shift);
    d->q_cap = d->p_cap = 0;
    LM_CUDA(cudaMalloc(&d->q_keys, (size_t)qc * SC_NR * sizeof(float)));
    LM_CUDA(cudaMalloc(&d->q_descs, (size_t)qc * SC_CELLS * sizeof(float)));
    LM_CUDA(cudaMalloc(&d->part_d, (size_t)pc * qc * SC_K * sizeof(float)));
    LM_CUDA(cudaMalloc(&d->part_i, (size_t)pc * qc * SC_K * sizeof(int)));
    LM_CUDA(cudaMalloc(&d->top_d, (size_t)qc * SC_K * sizeof(float)));
    LM_CUDA(cudaMalloc(&d->top_i, (size_t)qc * SC_K * sizeof(int)));
    LM_CUDA(cudaMalloc(&d->tau, (size_t)qc * sizeof(float)));
    LM_CUDA(cudaMalloc(&d->cand, (size_t)qc * SC_K * sizeof(ScCand)));
    LM_CUDA(cudaMalloc(&d->out_id, (size_t)qc * sizeof(int)));
    LM_CUDA(cudaMalloc(&d->out_dist, (size_t)qc * sizeof(double)));
    LM_CUDA(cudaMalloc(&d->out_shift, (size_t)qc * sizeof(int)));
    d->q_cap = qc;
    d->p_cap = pc;
  }
  return LMSF_OK;
}

// descriptor + ring key of one device cloud into device buffers
static int make_dev(Ctx* c, ScDb* d, const float4* d_pts, int n, float* d_desc, float* d_key) {
  if (!d->bins) LM_CUDA(cudaMalloc(&d->bins, SC_CELLS * sizeof(int)));
  LM_LAUNCH(c, k_sc_clear, div_up(SC_CELLS, 256), 256, 0, d->bins);
  if (n > 0) {
    int grid = div_up(n, 256 * 4);
    if (grid > 148 * 4) grid = 148 * 4;
    LM_LAUNCH(c, k_sc_bin, grid, 256, 0, d_pts, n, d->bins);
  }
  LM_LAUNCH(c, k_sc_finish, 1, 256, 0, d->bins, d_desc, d_key);
  return LMSF_OK;
}

static int stage_cloud(Ctx* c, ScDb* d, const float* xyzi, int n) {
  if (n > d->cloud_cap) {
    LM_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d->cloud);
    d->cloud = nullptr;
    d->cloud_cap = 0;
    int cap = n < 1024 ? 1024 : n;
    LM_CUDA(cudaMalloc(&d->cloud, (size_t)cap * sizeof(float4)));
    d->cloud_cap = cap;
  }
  if (n > 0) LM_CUDA(cudaMemcpyAsync(d->cloud, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
  return LMSF_OK;
}

// one exact scan of keys[0, limit) (optionally thresholded) + merge
static int scan_once(Ctx* c, ScDb* d, const float* d_q_keys, int nq, int limit, const float* tau, float* tau_out) {
  int parts = scan_parts(nq, limit);
  int per = div_up(limit, parts);
  per = div_up(per, KCHUNK) * KCHUNK;
  parts = div_up(limit, per);
  if (parts < 1) parts = 1;
  dim3 grid(div_up(nq, KQ), parts);
  LM_LAUNCH(c, k_sc_scan, grid, KQ * KW, 0, d->keys, limit, per, d_q_keys, nq, tau, d->part_d, d->part_i);
  LM_LAUNCH(c, k_sc_merge_parts, div_up(nq, 4), 128, 0, d->part_d, d->part_i, parts, nq, d->top_d, d->top_i, tau_out);
  return LMSF_OK;
}

// ring-key 10-NN of nq device queries among keys[0, limit): results in d->top_d / d->top_i.  Large databases are
// scanned twice: a 4096-key sample first, whose 10th distance bounds the 10th distance over the whole prefix, then
// everything with every partial list starting at that bound.
// The sample shrinks with the prefix (an eighth of it, 512 .. 4096 keys): a shard of 12 500 keys scanned without a
// bound spent 124 us, mostly in sorted inserts — with a 1 536-key sample first the two passes together take a third.
static int knn_dev(Ctx* c, ScDb* d, const float* d_q_keys, int nq, int limit) {
  LM_TRY(ensure_query(c, d, nq, scan_parts(nq, limit)));
  if (limit <= 0) {  // nothing of this shard is searchable: every slot empty
    LM_LAUNCH(c, k_sc_merge_parts, div_up(nq, 4), 128, 0, d->part_d, d->part_i, 0, nq, d->top_d, d->top_i,
              (float*)nullptr);
    return LMSF_OK;
  }
  const float* tau = nullptr;
  if (limit >= 4096) {
    int sample = limit / 8;
    sample = sample < 512 ? 512 : (sample > 4096 ? 4096 : sample);
    sample = div_up(sample, KCHUNK) * KCHUNK;
    LM_TRY(scan_once(c, d, d_q_keys, nq, sample, nullptr, d->tau));
    tau = d->tau;
  }
  return scan_once(c, d, d_q_keys, nq, limit, tau, nullptr);
}

// candidates of nq device queries against this shard -> d_cand[nq][10]
static int search_shard_dev(Ctx* c, ScDb* d, const float* d_q_keys, const float* d_q_descs, int nq, int limit,
                            int id_base, ScCand* d_cand) {
  LM_TRY(knn_dev(c, d, d_q_keys, nq, limit));
  dim3 grid(SC_K, nq);
  LM_LAUNCH(c, k_sc_dist_cand, grid, 64, 0, d_q_descs, d->descs, d->top_d, d->top_i, id_base, d_cand);
  return LMSF_OK;
}

}  // namespace lm

using namespace lm;

#define ENTER(c)                      \
  if (!(c)) return LMSF_ERR_INVALID;  \
  if (cudaSetDevice((c)->device) != cudaSuccess) return LMSF_ERR_NO_DEVICE

static int sync(Ctx* c) {
  LM_CUDA(cudaStreamSynchronize(c->stream));
  return LMSF_OK;
}

extern "C" {

int lmsf_sc_make(lmsf_ctx* c, const float* xyzi, int n, float* desc1200, float* key20) {
  ENTER(c);
  if (n < 0 || (n > 0 && !xyzi) || !desc1200 || !key20) return LMSF_ERR_INVALID;
  ScDb* d = scdb_of(c);
  LM_TRY(ensure_query(c, d, 1, 1));
  LM_TRY(stage_cloud(c, d, xyzi, n));
  LM_TRY(make_dev(c, d, d->cloud, n, d->q_descs, d->q_keys));
  LM_CUDA(cudaMemcpyAsync(desc1200, d->q_descs, SC_CELLS * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(key20, d->q_keys, SC_NR * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  return sync(c);
}

int lmsf_scdb_reserve(lmsf_ctx* c, int capacity) {
  ENTER(c);
  if (capacity < 0) return LMSF_ERR_INVALID;
  return ensure_db(c, scdb_of(c), capacity);
}

int lmsf_scdb_clear(lmsf_ctx* c) {
  ENTER(c);
  scdb_of(c)->n = 0;
  return LMSF_OK;
}

int lmsf_scdb_size(lmsf_ctx* c, int* n) {
  ENTER(c);
  if (!n) return LMSF_ERR_INVALID;
  *n = scdb_of(c)->n;
  return LMSF_OK;
}

int lmsf_scdb_add(lmsf_ctx* c, const float* descs, const float* keys, int n) {
  ENTER(c);
  if (n < 0 || (n > 0 && (!descs || !keys))) return LMSF_ERR_INVALID;
  ScDb* d = scdb_of(c);
  LM_TRY(ensure_db(c, d, d->n + n));
  if (n == 0) return LMSF_OK;
  LM_CUDA(cudaMemcpyAsync(d->keys + (size_t)d->n * SC_NR, keys, (size_t)n * SC_NR * sizeof(float),
                          cudaMemcpyHostToDevice, c->stream));
  LM_CUDA(cudaMemcpyAsync(d->descs + (size_t)d->n * SC_CELLS, descs, (size_t)n * SC_CELLS * sizeof(float),
                          cudaMemcpyHostToDevice, c->stream));
  LM_TRY(sync(c));
  d->n += n;
  return LMSF_OK;
}

int lmsf_scdb_add_cloud(lmsf_ctx* c, const float* xyzi, int n, int* id_out) {
  ENTER(c);
  if (n < 0 || (n > 0 && !xyzi)) return LMSF_ERR_INVALID;
  ScDb* d = scdb_of(c);
  LM_TRY(ensure_db(c, d, d->n + 1));
  LM_TRY(stage_cloud(c, d, xyzi, n));
  LM_TRY(make_dev(c, d, d->cloud, n, d->descs + (size_t)d->n * SC_CELLS, d->keys + (size_t)d->n * SC_NR));
  LM_TRY(sync(c));  // the caller's cloud may be released after the call
  if (id_out) *id_out = d->n;
  d->n += 1;
  return LMSF_OK;
}

int lmsf_scdb_get(lmsf_ctx* c, int id, float* desc1200, float* key20) {
  ENTER(c);
  ScDb* d = scdb_of(c);
  if (id < 0 || id >= d->n) return LMSF_ERR_INVALID;
  if (desc1200)
    LM_CUDA(cudaMemcpyAsync(desc1200, d->descs + (size_t)id * SC_CELLS, SC_CELLS * sizeof(float),
                            cudaMemcpyDeviceToHost, c->stream));
  if (key20)
    LM_CUDA(cudaMemcpyAsync(key20, d->keys + (size_t)id * SC_NR, SC_NR * sizeof(float), cudaMemcpyDeviceToHost,
                            c->stream));
  return sync(c);
}

int lmsf_scdb_knn(lmsf_ctx* c, const float* q_keys, int nq, int limit, int32_t* idx10, float* d10) {
  ENTER(c);
  ScDb* d = scdb_of(c);
  if (nq < 0 || limit < 0 || limit > d->n || (nq > 0 && (!q_keys || !idx10 || !d10))) return LMSF_ERR_INVALID;
  if (nq == 0) return LMSF_OK;
  LM_TRY(ensure_query(c, d, nq, scan_parts(nq, limit)));
  LM_CUDA(cudaMemcpyAsync(d->q_keys, q_keys, (size_t)nq * SC_NR * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  LM_TRY(knn_dev(c, d, d->q_keys, nq, limit));
  LM_CUDA(cudaMemcpyAsync(idx10, d->top_i, (size_t)nq * SC_K * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(d10, d->top_d, (size_t)nq * SC_K * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  return sync(c);
}

int lmsf_sc_distance(lmsf_ctx* c, const float* desc_a, const float* desc_b, int n_pairs, double* dist, int32_t* shift) {
  ENTER(c);
  if (n_pairs < 0 || (n_pairs > 0 && (!desc_a || !desc_b || !dist || !shift))) return LMSF_ERR_INVALID;
  if (n_pairs == 0) return LMSF_OK;
  float *da = nullptr, *db = nullptr;
  double* dd = nullptr;
  int* ds = nullptr;
  size_t bytes = (size_t)n_pairs * SC_CELLS * sizeof(float);
  int rc = LMSF_OK;
  if (cudaMalloc(&da, bytes) != cudaSuccess || cudaMalloc(&db, bytes) != cudaSuccess ||
      cudaMalloc(&dd, (size_t)n_pairs * 8) != cudaSuccess || cudaMalloc(&ds, (size_t)n_pairs * 4) != cudaSuccess) {
    c->last_error = "cudaMalloc (lmsf_sc_distance)";
    rc = LMSF_ERR_CUDA;
  } else {
    cudaMemcpyAsync(da, desc_a, bytes, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(db, desc_b, bytes, cudaMemcpyHostToDevice, c->stream);
    LM_LAUNCH(c, k_sc_dist_pairs, n_pairs, 64, 0, da, db, dd, ds);
    cudaMemcpyAsync(dist, dd, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(shift, ds, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
      c->last_error = cudaGetErrorString(cudaGetLastError());
      rc = LMSF_ERR_CUDA;
    }
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dd);
  cudaFree(ds);
  return rc;
}

int lmsf_scdb_search(lmsf_ctx* c, const float* q_keys, const float* q_descs, int nq, int limit, double thresh,
                     int32_t* loop_id, double* loop_dist, int32_t* loop_shift) {
  ENTER(c);
  ScDb* d = scdb_of(c);
  if (nq < 0 || limit < 1 || limit > d->n || (nq > 0 && (!q_keys || !q_descs || !loop_id || !loop_dist || !loop_shift)))
    return LMSF_ERR_INVALID;
  if (nq == 0) return LMSF_OK;
  LM_TRY(ensure_query(c, d, nq, scan_parts(nq, limit)));
  LM_CUDA(cudaMemcpyAsync(d->q_keys, q_keys, (size_t)nq * SC_NR * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  LM_CUDA(cudaMemcpyAsync(d->q_descs, q_descs, (size_t)nq * SC_CELLS * sizeof(float), cudaMemcpyHostToDevice,
                          c->stream));
  LM_TRY(search_shard_dev(c, d, d->q_keys, d->q_descs, nq, limit, 0, d->cand));
  LM_LAUNCH(c, k_sc_pick, div_up(nq, 128), 128, 0, d->cand, 1, nq, thresh, d->out_id, d->out_dist, d->out_shift);
  LM_CUDA(cudaMemcpyAsync(loop_id, d->out_id, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(loop_dist, d->out_dist, (size_t)nq * 8, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(loop_shift, d->out_shift, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
  return sync(c);
}

int lmsf_scdb_search_shard_dev(lmsf_ctx* c, const float* d_q_keys, const float* d_q_descs, int nq, int limit_local,
                               int id_base, void* d_cand) {
  ENTER(c);
  ScDb* d = scdb_of(c);
  if (nq < 0 || limit_local < 0 || limit_local > d->n || (nq > 0 && (!d_q_keys || !d_q_descs || !d_cand)))
    return LMSF_ERR_INVALID;
  if (nq == 0) return LMSF_OK;
  return search_shard_dev(c, d, d_q_keys, d_q_descs, nq, limit_local, id_base, (ScCand*)d_cand);
}

int lmsf_scdb_keys_shard_dev(lmsf_ctx* c, const float* d_q_keys, int nq, int limit_local, int id_base, void* d_cand) {
  ENTER(c);
  ScDb* d = scdb_of(c);
  if (nq < 0 || limit_local < 0 || limit_local > d->n || (nq > 0 && (!d_q_keys || !d_cand))) return LMSF_ERR_INVALID;
  if (nq == 0) return LMSF_OK;
  LM_TRY(knn_dev(c, d, d_q_keys, nq, limit_local));
  LM_LAUNCH(c, k_sc_keys_to_cand, div_up(nq * SC_K, 256), 256, 0, d->top_d, d->top_i, nq * SC_K, id_base, (ScCand*)d_cand);
  return LMSF_OK;
}

int lmsf_scdb_score_owned_dev(lmsf_ctx* c, const void* d_cand_all, int n_ranks, int nq, const float* d_q_descs,
                              int id_base, int n_local, void* d_scored) {
  ENTER(c);
  ScDb* d = scdb_of(c);
  if (nq < 0 || n_ranks < 1 || n_local < 0 || n_local > d->n || (nq > 0 && (!d_cand_all || !d_q_descs || !d_scored)))
    return LMSF_ERR_INVALID;
  if (nq == 0) return LMSF_OK;
  LM_TRY(ensure_query(c, d, nq, 1));
  LM_LAUNCH(c, k_sc_select_global, div_up(nq, 128), 128, 0, (const ScCand*)d_cand_all, n_ranks, nq, d->cand);
  dim3 grid(SC_K, nq);
  LM_LAUNCH(c, k_sc_dist_owned, grid, 64, 0, d_q_descs, d->descs, d->cand, id_base, n_local, (ScCand*)d_scored);
  return LMSF_OK;
}

int lmsf_scdb_pick_dev(lmsf_ctx* c, const void* d_cand_all, int n_ranks, int nq, double thresh, int32_t* d_loop_id,
                       double* d_loop_dist, int32_t* d_loop_shift) {
  ENTER(c);
  if (nq < 0 || n_ranks < 1 || (nq > 0 && (!d_cand_all || !d_loop_id || !d_loop_dist || !d_loop_shift)))
    return LMSF_ERR_INVALID;
  if (nq == 0) return LMSF_OK;
  LM_LAUNCH(c, k_sc_pick, div_up(nq, 128), 128, 0, (const ScCand*)d_cand_all, n_ranks, nq, thresh, d_loop_id,
            d_loop_dist, d_loop_shift);
  return LMSF_OK;
}

int lmsf_sc_tree_limit(int n_keyframes) {
  if (n_keyframes < 51) return 0;
  int s = ((n_keyframes - 1) / 10) * 10 + 1;  // size at the last rebuild: (size - 1) % 10 == 0
  return s - 50;
}

}  // extern "C"
