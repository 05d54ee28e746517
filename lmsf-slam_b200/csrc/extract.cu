// extract.cu — LOAM per-ring curvature and edge/surf feature selection on the GPU.
//
// Replaces LOAMFeatureProcessorBase::Process
// (Algorithm/PointClouds/processing/FeatureExtract/LOAMFeatureProcessor_base.hpp:59-126):
//   k_ring_classify / k_ring_scan / k_ring_scatter  = splitScan            (:290-343)
//   k_sector_sort                                   = curvature + std::sort (:97-118, :152)
//   k_ring_pick                                     = checkBadEdgePoint     (:216-282)
//                                                     + greedy edge pick    (:159-195)
//   k_feat_scatter                                  = edge / surf clouds    (:170-180, :197-206)
// Arithmetic follows the reference expression by expression (float stays float,
// double stays double; the library is built with -fmad=false because the
// reference is built without FMA).  The greedy pick keeps the reference's
// sequential semantics: `disable_point` is shared by the six sectors of a ring
// (:75-78), so one block walks a ring's sectors in order.
#include "common.cuh"
#include "fmath.cuh"

namespace lm {

static constexpr int RING_BLOCK = 256;
static constexpr int MAX_RINGS = 64;
static constexpr int SORT_SMEM = 2048;   // (key,val) pairs a sector can sort in shared memory
static constexpr int RING_SMEM = 16384;  // ring length whose flags fit shared memory

struct ExtractParams {
  int n_scans;
  float min_range, max_range, edge_thresh;
  int remove_bad;
};

// splitScan :297-340
__device__ __forceinline__ int ring_of(float4 p, const ExtractParams& prm) {
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return -1;
  float s = p.x * p.x + p.y * p.y;
  // sqrt(float) is std::sqrt(float) in the reference's translation unit (`using namespace std;`, utility.hpp:51):
  // a correctly rounded float square root, widened
  double distance = (double)__fsqrt_rn(s);
  if (distance > (double)prm.max_range || distance < (double)prm.min_range) return -1;
  double angle = atan((double)p.z / distance) * 180 / 3.14159265358979323846;
  int id;
  if (prm.n_scans == 16) {
    id = int((angle + 15) / 2 + 0.5);
    if (id > 15 || id < 0) return -1;
  } else if (prm.n_scans == 32) {
    id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
    if (id > 31 || id < 0) return -1;
  } else {
    if (angle >= -8.83)
      id = int((2 - angle) * 3.0 + 0.5);
    else
      id = 32 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return -1;
  }
  return id;
}

// ---- RotaryLidarPreProcess<PointXYZI>::Process (Preprocess/RotaryLidar_preprocessing.hpp:31-104), fused into the ring
// pass.  The reference walks the points in firing order with one bit of state (half_passed); that bit turns true at the
// FIRST point whose branch-A angle is more than pi past the start, so the walk is: (1) first / last point, (2) a minimum
// over the points of "would flip", (3) every point on its own.  float / double mixing as the reference's expressions
// promote it: comparisons and the +- 2 pi in double, the angles stored back to float, rel_time in float.
// Points with a non-finite coordinate are passed over: the node's removeNaNFromPointCloud drops them first
// (MultiLidarSLAM_node.cpp:126-133).
struct RotaryState {
  int first, last, half, pad;
};
#define ROT_PI 3.14159265358979323846
__device__ __forceinline__ bool rot_valid(float4 p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }
// findStartEndAngle :80-94
__device__ __forceinline__ void rot_start_end(float4 p0, float4 pl, float& so, float& eo) {
  so = -atan2f_fdlibm(p0.y, p0.x);
  eo = (float)((double)(-atan2f_fdlibm(pl.y, pl.x)) + 2 * ROT_PI);
  if ((double)(eo - so) > 3 * ROT_PI)
    eo = (float)((double)eo - 2 * ROT_PI);
  else if ((double)(eo - so) < ROT_PI)
    eo = (float)((double)eo + 2 * ROT_PI);
}
// :42-55, the branch taken while half_passed is false; flips: this point turns it true
__device__ __forceinline__ float rot_first_half(float4 p, float so, bool& flips) {
  float ori = -atan2f_fdlibm(p.y, p.x);
  if ((double)ori < (double)so - ROT_PI / 2)
    ori = (float)((double)ori + 2 * ROT_PI);
  else if ((double)ori > (double)so + ROT_PI * 3 / 2)
    ori = (float)((double)ori - 2 * ROT_PI);
  flips = (double)(ori - so) > ROT_PI;
  return ori;
}
// :57-67
__device__ __forceinline__ float rot_second_half(float4 p, float eo) {
  float ori = -atan2f_fdlibm(p.y, p.x);
  ori = (float)((double)ori + 2 * ROT_PI);
  if ((double)ori < (double)eo - ROT_PI * 3 / 2)
    ori = (float)((double)ori + 2 * ROT_PI);
  else if ((double)ori > (double)eo + ROT_PI / 2)
    ori = (float)((double)ori - 2 * ROT_PI);
  return ori;
}
// relative time of point i (:68), given the sweep's state
__device__ __forceinline__ float rot_rel_time(const float4* __restrict__ in, float4 p, int i, const RotaryState& rs,
                                              float period) {
  float so, eo;
  rot_start_end(in[rs.first], in[rs.last], so, eo);
  bool flips;
  const float ori = (i <= rs.half) ? rot_first_half(p, so, flips) : rot_second_half(p, eo);
  return (ori - so) / (eo - so) * period;
}

__global__ void k_rotary_init(RotaryState* rs) {
  if (threadIdx.x == 0) {
    rs->first = 0x7fffffff;
    rs->last = -1;
    rs->half = 0x7fffffff;
    rs->pad = 0;
  }
}
__device__ __forceinline__ void rot_note_bounds(bool valid, int i, RotaryState* rs) {
  const unsigned m = __ballot_sync(0xffffffffu, valid);
  if (m && (threadIdx.x & 31) == 0) {  // lanes are consecutive points: the warp's first and last finite point
    const int base = i;                // lane 0's point
    atomicMin(&rs->first, base + __ffs(m) - 1);
    atomicMax(&rs->last, base + 31 - __clz(m));
  }
}
// standalone bounds pass (lmsf_rotary_preprocess); the extraction does it inside k_ring_classify
__global__ void __launch_bounds__(256) k_rotary_bounds(const float4* __restrict__ in, int n, RotaryState* rs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  rot_note_bounds(i < n && rot_valid(in[i]), i, rs);
}
__global__ void __launch_bounds__(256) k_rotary_half(const float4* __restrict__ in, int n, RotaryState* rs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int first = rs->first, last = rs->last;
  bool flips = false;
  if (i < n && last >= 0) {
    const float4 p = in[i];
    if (rot_valid(p)) {
      float so, eo;
      rot_start_end(in[first], in[last], so, eo);
      rot_first_half(p, so, flips);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, flips);
  if (m && (threadIdx.x & 31) == 0) atomicMin(&rs->half, i + __ffs(m) - 1);
}
__global__ void __launch_bounds__(256) k_rotary_apply(float4* __restrict__ pts, int n, const RotaryState* __restrict__ rs,
                                                      float period) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RotaryState st = *rs;
  float4 p = pts[i];
  if (!rot_valid(p)) return;
  p.w = rot_rel_time(pts, p, i, st, period);
  pts[i] = p;
}

__global__ void __launch_bounds__(RING_BLOCK) k_ring_classify(const float4* __restrict__ in, int n, ExtractParams prm,
                                                              int nblk, int* __restrict__ ring_id,
                                                              int* __restrict__ blk_cnt, RotaryState* rotary) {
  __shared__ int hist[MAX_RINGS];
  if (threadIdx.x < MAX_RINGS) hist[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * RING_BLOCK + threadIdx.x;
  if (rotary) rot_note_bounds(i < n && rot_valid(in[i]), i, rotary);  // first / last point of the NaN-free sweep
  if (i < n) {
    int id = ring_of(in[i], prm);
    ring_id[i] = id;
    if (id >= 0) atomicAdd(&hist[id], 1);
  }
  __syncthreads();
  if (threadIdx.x < MAX_RINGS) blk_cnt[threadIdx.x * nblk + blockIdx.x] = hist[threadIdx.x];
}

// per ring: exclusive scan of the per-block counts; then ring offsets
__global__ void __launch_bounds__(1024) k_ring_scan(int* __restrict__ blk_cnt, int nblk, int* __restrict__ ring_cnt,
                                                    int* __restrict__ ring_off, int* __restrict__ counts) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < MAX_RINGS; r += 32) {
    int carry = 0;
    for (int b0 = 0; b0 < nblk; b0 += 32) {
      int b = b0 + lane;
      int v = (b < nblk) ? blk_cnt[r * nblk + b] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (b < nblk) blk_cnt[r * nblk + b] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) ring_cnt[r] = carry;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < MAX_RINGS; ++r) {
      ring_off[r] = acc;
      acc += ring_cnt[r];
    }
    ring_off[MAX_RINGS] = acc;
    counts[2] = acc;
  }
}

// order-preserving scatter of the points into their rings
__global__ void __launch_bounds__(RING_BLOCK) k_ring_scatter(const float4* __restrict__ in, int n, int nblk,
                                                             const int* __restrict__ ring_id,
                                                             const int* __restrict__ blk_off,
                                                             const int* __restrict__ ring_off,
                                                             float4* __restrict__ ring_pts, int* __restrict__ ring_src,
                                                             const RotaryState* __restrict__ rotary, float period) {
  __shared__ int wcnt[RING_BLOCK / 32][MAX_RINGS];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < (RING_BLOCK / 32) * MAX_RINGS; k += RING_BLOCK) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  int i = blockIdx.x * RING_BLOCK + threadIdx.x;
  int id = (i < n) ? ring_id[i] : -1;
  unsigned peers = __match_any_sync(0xffffffffu, id);
  int rank = __popc(peers & ((1u << lane) - 1u));
  if (id >= 0 && rank == 0) wcnt[warp][id] = __popc(peers);
  __syncthreads();
  if (id >= 0) {
    int base = 0;
    for (int w = 0; w < warp; ++w) base += wcnt[w][id];
    int pos = ring_off[id] + blk_off[id * nblk + blockIdx.x] + base + rank;
    float4 p = in[i];
    if (rotary) p.w = rot_rel_time(in, p, i, *rotary, period);  // RotaryLidarPreProcess: intensity := relative time
    ring_pts[pos] = p;
    ring_src[pos] = i;
  }
}

// sector geometry shared by the three per-sector kernels (:71-90)
struct Sector {
  int P, off, s, n;
  bool live;
};
__device__ __forceinline__ Sector sector_of(const int* ring_cnt, const int* ring_off, int r, int k) {
  Sector sc;
  sc.P = ring_cnt[r];
  sc.off = ring_off[r];
  sc.live = !(sc.P < 20 || sc.P - 10 < 6);
  int len = (sc.P - 10) / 6;
  sc.s = 5 + len * k;
  int e = (k == 5) ? sc.P - 6 : sc.s + len - 1;
  sc.n = sc.live ? (e - sc.s + 1) : 0;
  return sc;
}

// curvature (:97-118): eleven-term float sum, left to right, then double squares
__device__ __forceinline__ double curvature_at(const float4* __restrict__ p, int j) {
  float4 a0 = p[j - 5], a1 = p[j - 4], a2 = p[j - 3], a3 = p[j - 2], a4 = p[j - 1], c = p[j];
  float4 b1 = p[j + 1], b2 = p[j + 2], b3 = p[j + 3], b4 = p[j + 4], b5 = p[j + 5];
  float sx = a0.x + a1.x + a2.x + a3.x + a4.x - 10 * c.x + b1.x + b2.x + b3.x + b4.x + b5.x;
  float sy = a0.y + a1.y + a2.y + a3.y + a4.y - 10 * c.y + b1.y + b2.y + b3.y + b4.y + b5.y;
  float sz = a0.z + a1.z + a2.z + a3.z + a4.z - 10 * c.z + b1.z + b2.z + b3.z + b4.z + b5.z;
  double dx = sx, dy = sy, dz = sz;
  return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ bool kv_less(double ka, int va, double kb, int vb) {
  return ka < kb || (ka == kb && va < vb);
}

// one block per (sector, ring): curvature, then a bitonic sort ascending by (curvature, id)
__global__ void __launch_bounds__(256) k_sector_sort(const float4* __restrict__ ring_pts,
                                                     const int* __restrict__ ring_cnt,
                                                     const int* __restrict__ ring_off, double* __restrict__ curv,
                                                     int* __restrict__ sorted, double* __restrict__ g_key,
                                                     int* __restrict__ g_val) {
  __shared__ double s_key[SORT_SMEM];
  __shared__ int s_val[SORT_SMEM];
  Sector sc = sector_of(ring_cnt, ring_off, blockIdx.y, blockIdx.x);
  if (!sc.live) return;
  int npad = 1;
  while (npad < sc.n) npad <<= 1;
  double* key = s_key;
  int* val = s_val;
  if (npad > SORT_SMEM) {  // sector too long for shared memory: same sort in global scratch
    key = g_key + 2 * (size_t)(sc.off + sc.s);
    val = g_val + 2 * (size_t)(sc.off + sc.s);
  }
  const float4* rp = ring_pts + sc.off;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    if (i < sc.n) {
      int j = sc.s + i;
      double cv = curvature_at(rp, j);
      curv[sc.off + j] = cv;
      key[i] = cv;
      val[i] = j;
    } else {
      key[i] = __longlong_as_double(0x7ff0000000000000LL);  // +inf pads the tail
      val[i] = 0x7fffffff;
    }
  }
  __syncthreads();
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        int lo = ((t / stride) * (stride << 1)) + (t % stride);
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        double ka = key[lo], kb = key[hi];
        int va = val[lo], vb = val[hi];
        bool swap = up ? kv_less(kb, vb, ka, va) : kv_less(ka, va, kb, vb);
        if (swap) {
          key[lo] = kb;
          key[hi] = ka;
          val[lo] = vb;
          val[hi] = va;
        }
      }
      __syncthreads();
    }
  }
  int* out = sorted + sc.off + (sc.s - 5);
  for (int i = threadIdx.x; i < sc.n; i += blockDim.x) out[i] = val[i];
}

// checkBadEdgePoint class of position j (:221-279), ignoring the skip logic:
// 0 nothing, 1 azimuth gap (disable j-5..j+5, skip 4), 2 occlusion with the
// nearer point first (disable j+1..j+5, skip 4), 3 occlusion with the nearer
// point second (disable j-5..j).
__device__ __forceinline__ int bad_class(const float4* __restrict__ p, int j) {
  float4 a = p[j], b = p[j + 1];
  // atan2 of two floats = std::atan2(float, float) = the C library's atan2f (:223-224); fmath.cuh returns its bits
  double a0 = (double)atan2f_fdlibm(a.x, a.y);
  double a1 = (double)atan2f_fdlibm(b.x, b.y);
  double da = fabs(a0 - a1);
  const double PI = 3.14159265358979323846;
  if (da > PI) da = PI * 2 - da;
  if (da > 0.0175) return 1;
  float s0 = a.x * a.x + a.y * a.y + a.z * a.z;
  float s1 = b.x * b.x + b.y * b.y + b.z * b.z;
  double d0 = (double)__fsqrt_rn(s0), d1 = (double)__fsqrt_rn(s1);  // std::sqrt(float) (:247-252)
  double ang = (d0 < d1) ? atan2(d0 * da, d1 - d0) : atan2(d1 * da, d0 - d1);
  if (ang <= 0.17) return (d0 < d1) ? 2 : 3;
  return 0;
}

// one block per ring: bad-point marks, then the greedy edge pick over the ring's six sectors
__global__ void __launch_bounds__(256) k_ring_pick(const float4* __restrict__ ring_pts,
                                                   const int* __restrict__ ring_cnt,
                                                   const int* __restrict__ ring_off, const double* __restrict__ curv,
                                                   const int* __restrict__ sorted, ExtractParams prm,
                                                   uint8_t* __restrict__ g_flag, uint8_t* __restrict__ g_type,
                                                   uint8_t* __restrict__ is_edge, int* __restrict__ edge_ids,
                                                   int* __restrict__ sec_cnt) {
  __shared__ uint8_t s_flag[RING_SMEM];
  __shared__ uint8_t s_type[RING_SMEM];
  const int r = blockIdx.x;
  const int P = ring_cnt[r], off = ring_off[r];
  if (P < 20 || P - 10 < 6) {
    if (threadIdx.x < 6) sec_cnt[r * 6 + threadIdx.x] = 0;
    return;
  }
  uint8_t* dis = (P <= RING_SMEM) ? s_flag : g_flag + off;
  uint8_t* typ = (P <= RING_SMEM) ? s_type : g_type + off;
  const float4* rp = ring_pts + off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    dis[j] = 0;
    typ[j] = 0;
    is_edge[off + j] = 0;
  }
  __syncthreads();
  if (prm.remove_bad) {
    for (int j = 5 + threadIdx.x; j < P - 6; j += blockDim.x) typ[j] = (uint8_t)bad_class(rp, j);
    __syncthreads();
    // the reference loop skips four positions after a class 1 / class 2 hit (:243,:270):
    // positions inside a taken skip are never examined, so their class is cleared.
    if (warp == 0) {
      int j = 5;
      while (j < P - 6) {
        int jj = j + lane;
        int t = (jj < P - 6) ? typ[jj] : 0;
        unsigned jump = __ballot_sync(0xffffffffu, t == 1 || t == 2);
        if (jump == 0) {
          j += 32;
        } else {
          int f = __ffs(jump) - 1;
          int hit = j + f;
          if (lane >= 1 && lane <= 4 && hit + lane < P - 6) typ[hit + lane] = 0;
          __syncwarp();
          j = hit + 5;
        }
      }
    }
    __syncthreads();
    for (int j = 5 + threadIdx.x; j < P - 6; j += blockDim.x) {
      int t = typ[j];
      if (t == 1) {
        for (int m = -5; m <= 5; ++m) dis[j + m] = 1;
      } else if (t == 2) {
        for (int m = 1; m <= 5; ++m) dis[j + m] = 1;
      } else if (t == 3) {
        for (int m = 0; m <= 5; ++m) dis[j - m] = 1;
      }
    }
    __syncthreads();
  }
  if (warp != 0) return;
  // featureExtractionFromSector (:145-195), sectors in order, one warp
  const double thresh = (double)prm.edge_thresh;
  const int len = (P - 10) / 6;
  for (int k = 0; k < 6; ++k) {
    int s = 5 + len * k;
    int e = (k == 5) ? P - 6 : s + len - 1;
    int n = e - s + 1;
    const int* ord = sorted + off + (s - 5);
    int picked = 0;
    bool done = false;
    for (int top = n - 1; top >= 0 && !done; top -= 32) {
      int ii = top - lane;
      bool valid = ii >= 0;
      int ind = valid ? ord[ii] : 0;
      double cv = valid ? curv[off + ind] : 0.0;
      unsigned pending = __ballot_sync(0xffffffffu, valid);
      while (pending) {
        bool en = ((pending >> lane) & 1u) && dis[ind] == 0;
        unsigned enb = __ballot_sync(0xffffffffu, en);
        if (enb == 0) break;
        int l = __ffs(enb) - 1;
        double cvl = __shfl_sync(0xffffffffu, cv, l);
        int indl = __shfl_sync(0xffffffffu, ind, l);
        if (cvl <= thresh) {
          done = true;
          break;
        }
        picked++;
        if (picked > 20) {
          done = true;
          break;
        }
        if (lane == 0) {
          edge_ids[(r * 6 + k) * 20 + picked - 1] = indl;
          is_edge[off + indl] = 1;
        }
        if (lane < 10) {
          int m = (lane < 5) ? (lane + 1) : -(lane - 4);
          int nn = indl + m;
          nn = nn >= P ? P - 1 : (nn < 0 ? 0 : nn);
          dis[nn] = 1;
        }
        __syncwarp();
        pending &= ~((2u << l) - 1u);
      }
    }
    if (lane == 0) sec_cnt[r * 6 + k] = picked > 20 ? 20 : picked;
    __syncwarp();
  }
}

// one block per (sector, ring): write the sector's edges and surfs at their global offsets
__global__ void __launch_bounds__(256) k_feat_scatter(const float4* __restrict__ ring_pts,
                                                      const int* __restrict__ ring_src,
                                                      const int* __restrict__ ring_cnt,
                                                      const int* __restrict__ ring_off, const int* __restrict__ sorted,
                                                      const uint8_t* __restrict__ is_edge,
                                                      const int* __restrict__ edge_ids, const int* __restrict__ sec_cnt,
                                                      int n_scans, float4* __restrict__ feat,
                                                      uint8_t* __restrict__ label, int* __restrict__ counts,
                                                      int* __restrict__ surf_rank, int* __restrict__ perm) {
  __shared__ int red[4][8];
  __shared__ int tot[4];
  __shared__ int wsum[8];
  const int r = blockIdx.y, k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // prefix of edge / surf counts over the sectors before (r,k), and the totals
  int me = r * 6 + k;
  int e_before = 0, s_before = 0, e_total = 0, s_total = 0;
  for (int t = threadIdx.x; t < n_scans * 6; t += blockDim.x) {
    Sector o = sector_of(ring_cnt, ring_off, t / 6, t % 6);
    int ec = o.live ? sec_cnt[t] : 0;
    int scn = o.n - ec;
    e_total += ec;
    s_total += scn;
    if (t < me) {
      e_before += ec;
      s_before += scn;
    }
  }
  int v[4] = {e_before, s_before, e_total, s_total};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int x = v[q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (lane == 0) red[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int x = 0;
    for (int w = 0; w < 8; ++w) x += red[threadIdx.x][w];
    tot[threadIdx.x] = x;
  }
  __syncthreads();
  e_before = tot[0];
  s_before = tot[1];
  e_total = tot[2];
  s_total = tot[3];
  if (me == 0 && threadIdx.x == 0) {
    counts[0] = e_total;
    counts[1] = s_total;
  }
  Sector sc = sector_of(ring_cnt, ring_off, r, k);
  if (!sc.live) return;
  const float4* rp = ring_pts + sc.off;
  const int* src = ring_src + sc.off;
  int ec = sec_cnt[me];
  if ((int)threadIdx.x < ec) {
    int id = edge_ids[me * 20 + threadIdx.x];
    feat[e_before + threadIdx.x] = rp[id];
    label[src[id]] = 1;
    perm[e_before + threadIdx.x] = e_before + threadIdx.x;  // edges keep their order in the processing permutation
  }
  // surfs: every id of the ascending order that is not an edge (:197-206)
  const int* ord = sorted + sc.off + (sc.s - 5);
  float4* surf = feat + e_total + s_before;
  int run = 0;
  for (int base = 0; base < sc.n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int id = (i < sc.n) ? ord[i] : 0;
    int f = (i < sc.n && is_edge[sc.off + id] == 0) ? 1 : 0;
    int inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int wbase = 0, all = 0;
    for (int w = 0; w < 8; ++w) {
      int x = wsum[w];
      if (w < warp) wbase += x;
      all += x;
    }
    if (f) {
      surf[run + wbase + inc - 1] = rp[id];
      label[src[id]] = 2;
      surf_rank[sc.off + id] = e_total + s_before + run + wbase + inc - 1;
    }
    run += all;
    __syncthreads();
  }
  // Processing permutation of the registration (a locality order, any permutation is correct): the surfs of this
  // sector in firing order instead of curvature order, so that the 32 queries of a warp are neighbours along the
  // ring.  perm[position] = feature index; the sector owns the same position range as in the feature cloud.
  run = 0;
  for (int base = 0; base < sc.n; base += blockDim.x) {
    int i = base + threadIdx.x;
    int id = sc.s + i;
    int f = (i < sc.n && is_edge[sc.off + id] == 0) ? 1 : 0;
    int inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int wbase = 0, all = 0;
    for (int w = 0; w < 8; ++w) {
      int x = wsum[w];
      if (w < warp) wbase += x;
      all += x;
    }
    if (f) perm[e_total + s_before + run + wbase + inc - 1] = surf_rank[sc.off + id];
    run += all;
    __syncthreads();
  }
}

// ---------------------------------------------------------------- host side
int extract_alloc(Ctx* c) {
  ExtractBufs& x = c->ex;
  x.cap = c->prm.max_points;
  x.nblk_cap = div_up(x.cap, RING_BLOCK);
  size_t cap = (size_t)x.cap;
  LM_CUDA(cudaMalloc(&x.ring_id, cap * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.rotary, 4 * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.blk_cnt, (size_t)MAX_RINGS * x.nblk_cap * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.ring_cnt, MAX_RINGS * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.ring_off, (MAX_RINGS + 1) * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.ring_pts, (cap + 16) * sizeof(float4)));
  LM_CUDA(cudaMalloc(&x.ring_src, cap * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.curv, cap * sizeof(double)));
  LM_CUDA(cudaMalloc(&x.sorted, cap * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.sort_key, (2 * cap + 64) * sizeof(double)));
  LM_CUDA(cudaMalloc(&x.sort_val, (2 * cap + 64) * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.flag, cap));
  LM_CUDA(cudaMalloc(&x.btype, cap));
  LM_CUDA(cudaMalloc(&x.is_edge, cap));
  LM_CUDA(cudaMalloc(&x.edge_ids, MAX_RINGS * 6 * 20 * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.sec_cnt, MAX_RINGS * 6 * sizeof(int)));
  LM_CUDA(cudaMalloc(&x.label, cap));
  LM_CUDA(cudaMalloc(&x.surf_rank, cap * sizeof(int)));
  return LMSF_OK;
}

void extract_free(Ctx* c) {
  ExtractBufs& x = c->ex;
  cudaFree(x.ring_id);
  cudaFree(x.rotary);
  cudaFree(x.blk_cnt);
  cudaFree(x.ring_cnt);
  cudaFree(x.ring_off);
  cudaFree(x.ring_pts);
  cudaFree(x.ring_src);
  cudaFree(x.curv);
  cudaFree(x.sorted);
  cudaFree(x.sort_key);
  cudaFree(x.sort_val);
  cudaFree(x.flag);
  cudaFree(x.btype);
  cudaFree(x.is_edge);
  cudaFree(x.edge_ids);
  cudaFree(x.sec_cnt);
  cudaFree(x.label);
  cudaFree(x.surf_rank);
  x = ExtractBufs();
}

int extract_run(Ctx* c, const float4* d_in, int n, cudaStream_t st, float4* feat_out, int* counts_out,
                int* perm_out) {
  ExtractBufs& x = c->ex;
  if (n > x.cap) return LMSF_ERR_CAPACITY;
  StageScope scope(c, LMSF_STAGE_EXTRACT, st);
  ExtractParams prm{c->prm.n_scans, c->prm.min_range, c->prm.max_range, c->prm.edge_thresh, c->prm.remove_bad_points};
  const int R = c->prm.n_scans;
  int nblk = div_up(n > 0 ? n : 1, RING_BLOCK);
  LM_CUDA(cudaMemsetAsync(x.label, 0, n > 0 ? n : 1, st));
  // rotary_scan_period > 0: removeNaN + RotaryLidarPreProcess ride along (bounds in the classify pass, one small kernel
  // for the half-sweep point, the relative time written while the points move into ring order)
  const float period = c->prm.rotary_scan_period;
  RotaryState* rot = (period > 0.f && n > 0) ? reinterpret_cast<RotaryState*>(x.rotary) : nullptr;
  if (rot) LM_LAUNCH_ON(c, st, k_rotary_init, 1, 32, 0, rot);
  LM_LAUNCH_ON(c, st, k_ring_classify, nblk, RING_BLOCK, 0, d_in, n, prm, nblk, x.ring_id, x.blk_cnt, rot);
  if (rot) LM_LAUNCH_ON(c, st, k_rotary_half, div_up(n, 256), 256, 0, d_in, n, rot);
  LM_LAUNCH_ON(c, st, k_ring_scan, 1, 1024, 0, x.blk_cnt, nblk, x.ring_cnt, x.ring_off, counts_out);
  LM_LAUNCH_ON(c, st, k_ring_scatter, nblk, RING_BLOCK, 0, d_in, n, nblk, x.ring_id, x.blk_cnt, x.ring_off, x.ring_pts,
               x.ring_src, rot, period);
  LM_LAUNCH_ON(c, st, k_sector_sort, dim3(6, R), 256, 0, x.ring_pts, x.ring_cnt, x.ring_off, x.curv, x.sorted,
               x.sort_key, x.sort_val);
  LM_LAUNCH_ON(c, st, k_ring_pick, R, 256, 0, x.ring_pts, x.ring_cnt, x.ring_off, x.curv, x.sorted, prm, x.flag,
               x.btype, x.is_edge, x.edge_ids, x.sec_cnt);
  LM_LAUNCH_ON(c, st, k_feat_scatter, dim3(6, R), 256, 0, x.ring_pts, x.ring_src, x.ring_cnt, x.ring_off, x.sorted,
               x.is_edge, x.edge_ids, x.sec_cnt, R, feat_out, x.label, counts_out, x.surf_rank, perm_out);
  LM_CUDA(cudaGetLastError());
  return LMSF_OK;
}

int rotary_apply(Ctx* c, float4* d_pts, int n, float scan_period, cudaStream_t st) {
  if (n <= 0) return LMSF_OK;
  RotaryState* rot = reinterpret_cast<RotaryState*>(c->ex.rotary);
  LM_LAUNCH_ON(c, st, k_rotary_init, 1, 32, 0, rot);
  LM_LAUNCH_ON(c, st, k_rotary_bounds, div_up(n, 256), 256, 0, d_pts, n, rot);
  LM_LAUNCH_ON(c, st, k_rotary_half, div_up(n, 256), 256, 0, d_pts, n, rot);
  LM_LAUNCH_ON(c, st, k_rotary_apply, div_up(n, 256), 256, 0, d_pts, n, rot, scan_period);
  LM_CUDA(cudaGetLastError());
  return LMSF_OK;
}

}  // namespace lm
