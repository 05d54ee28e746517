// knn.cuh — device-side exact 5-NN over the local-map grid (three nested cell levels).
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(point, 5, ...) as called by
// EdgeFeatureMatch::Match (registration/FeatureMatch/EdgeFeatureMatch.hpp:38) and
// SurfFeatureMatch::Match (surfFeatureMatch.hpp:37).  Both callers reject the
// query unless the 5th squared distance is < search_thresh_ = 1.0
// (FeatureMatchBase.hpp:29), so only neighbours with d2 < 1.0 are ever needed:
// the 27 coarse (1 m) cells around the query hold all of them.
//
// Levels: L0 = 1 m cells in a hash table; every L0 cell has a 64-bit occupancy
// mask of its 4x4x4 L1 cells (0.25 m); every occupied L1 cell has a 64-bit mask
// of its 4x4x4 L2 cells (0.0625 m).  Points are grouped by (L0, L1, L2) so each
// cell at each level is one contiguous range.  Search:
//   seed  (outer iterations after the first) the previous iteration's five neighbours give an
//         upper bound of the 5th distance before anything is scanned,
//   A     otherwise the 27 L2 cells around the query give it (dense regions end here: exact as
//         soon as the 5th distance is < 0.0625^2),
//   ball  with a bound < 0.25^2 the search restarts as a row sweep over exactly the L2 cells the
//         ball of that radius touches, nearest rows first; insertions are rare because the bound
//         is already tight,
//   B, C  without one (sparse surroundings) the 27 L1 cells, then the 27 L0 cells, are swept
//         cell by cell, skipping every cell whose box is farther than the current 5th distance.
// Cell sizes are powers of two, so cell indices and cell bounds are exact in
// fp32 and the box distance — computed with the same rounding sequence as a
// point distance — never exceeds the distance of a point inside the box:
// pruning is exact, ties included.
//
// Distances are FLANN's L2_Simple: ((dx*dx)+dy*dy)+dz*dz in fp32.  Results are
// ascending by (distance, original index): ties are resolved by index, which
// FLANN resolves by traversal order ("identical except at exact ties").
#pragma once
#include "common.cuh"

namespace lm {

// the 27 neighbour offsets, nearest first (centre, 6 faces, 12 edges, 8 corners): the 5th-distance bound
// tightens on the first cells and prunes most of the later ones
__device__ static const signed char kNear27[27][3] = {
    {0, 0, 0},   {-1, 0, 0},  {1, 0, 0},   {0, -1, 0},  {0, 1, 0},   {0, 0, -1},  {0, 0, 1},
    {-1, -1, 0}, {1, -1, 0},  {-1, 1, 0},  {1, 1, 0},   {-1, 0, -1}, {1, 0, -1},  {-1, 0, 1},
    {1, 0, 1},   {0, -1, -1}, {0, 1, -1},  {0, -1, 1},  {0, 1, 1},   {-1, -1, -1}, {1, -1, -1},
    {-1, 1, -1}, {1, 1, -1},  {-1, -1, 1}, {1, -1, 1},  {-1, 1, 1},  {1, 1, 1}};

struct MapView {
  const float4* sorted;             // cell-sorted points, .w = original index bits
  const CellRec* table;             // hash table of L0 cells
  const L1Rec* l1;                  // per occupied L1 cell: occupancy of its L2 cells, index of its first L2 cell
  const int* l2_start;              // start of every occupied L2 cell in `sorted` (+ one sentinel per L0 cell)
  const MapDev* dev;
};

struct Top5 {
  float d[5];
  int id[5];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      d[k] = 1.0f;  // search_thresh_: only neighbours with d2 < 1.0 can enter
      id[k] = -1;
    }
  }
  // five points are known to exist with d2 <= bound: accept d2 <= bound, ties included
  __device__ __forceinline__ void reset_inclusive(float bound) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      d[k] = bound;
      id[k] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ void add(float dd, int ii) {
    if (dd < d[4] || (dd == d[4] && ii < id[4])) {
      d[4] = dd;
      id[4] = ii;
#pragma unroll
      for (int k = 4; k > 0; --k) {
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
  __device__ __forceinline__ bool full() const { return id[4] >= 0 && id[4] != 0x7fffffff; }
};

__device__ __forceinline__ unsigned hash_cell(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (unsigned)k;
}

__device__ __forceinline__ unsigned long long pack_cell(const MapDev& md, int cx, int cy, int cz) {
  return ((((unsigned long long)(unsigned)cz << md.bits[1]) | (unsigned long long)(unsigned)cy) << md.bits[0]) |
         (unsigned long long)(unsigned)cx;
}

struct KnnStats {  // debug counters (debug_stats.cu): per-query work and the pass the search ended in
  int cand, boxes, lookups, level;
};
#ifdef LMSF_KNN_STATS
#define KSTAT(x) x
#define KS_DECL , KnnStats& ks
#define KS_PASS , ks
#else
#define KSTAT(x)
#define KS_DECL
#define KS_PASS
#endif

// one-entry cache in front of the hash probe: neighbouring cells mostly share their L0 cell
struct CellCursor {
  unsigned long long key;
  const CellRec* rec;
};

__device__ __forceinline__ const CellRec* find_cell(const MapView& mv, const MapDev& md, CellCursor& cur, int cx,
                                                    int cy, int cz KS_DECL) {
  KSTAT(ks.lookups++;)
  if (cx < 0 || cy < 0 || cz < 0 || cx >= md.dim[0] || cy >= md.dim[1] || cz >= md.dim[2]) return nullptr;
  unsigned long long key = pack_cell(md, cx, cy, cz);
  if (key == cur.key) return cur.rec;
  unsigned h = hash_cell(key) & md.table_mask;
  const CellRec* r = nullptr;
  for (unsigned probes = 0; probes <= md.table_mask; ++probes) {  // bounded: a full table must not hang the search
    unsigned long long k = mv.table[h].key;
    if (k == key) {
      r = &mv.table[h];
      break;
    }
    if (k == ~0ull) break;
    h = (h + 1) & md.table_mask;
  }
  cur.key = key;
  cur.rec = r;
  return r;
}

// `seeded`: the five slots were preloaded with real map points (the previous iteration's neighbours at their new
// distances); those points are met again by the sweep and must not enter twice.  A candidate can only pass the
// acceptance test of add() as a duplicate of slots 0..3 (slot 4 itself fails it), so four compares settle it.
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz,
                                           Top5& nb, bool seeded KS_DECL) {
  KSTAT(ks.cand += e - s;)
  if (s >= e) return;
  float4 nxt = __ldg(&pts[s]);  // the next candidate is in flight while the current one is judged
  for (int p = s; p < e; ++p) {
    const float4 m = nxt;
    if (p + 1 < e) nxt = __ldg(&pts[p + 1]);
    float dx = m.x - qx, dy = m.y - qy, dz = m.z - qz;
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    const int ii = __float_as_int(m.w);
    if (r < nb.d[4] || (r == nb.d[4] && ii < nb.id[4])) {
      if (seeded && (ii == nb.id[0] || ii == nb.id[1] || ii == nb.id[2] || ii == nb.id[3])) continue;
      nb.add(r, ii);
    }
  }
}

// squared distance from q to the axis-aligned cell [c*s, (c+1)*s)^3, same rounding sequence as scan_range
__device__ __forceinline__ float axis_gap(float q, int c, float s) {
  float lo = (float)c * s;
  float hi = lo + s;
  return (q < lo) ? (lo - q) : ((q > hi) ? (q - hi) : 0.0f);
}
__device__ __forceinline__ float box_d2(float qx, float qy, float qz, int cx, int cy, int cz, float s) {
  float dx = axis_gap(qx, cx, s), dy = axis_gap(qy, cy, s), dz = axis_gap(qz, cz, s);
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

// one 16-byte load per visited L1 cell
__device__ __forceinline__ L1Rec ldg_l1(const MapView& mv, int l1) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(mv.l1 + l1));
  L1Rec r;
  r.mask = ((unsigned long long)v.y << 32) | (unsigned long long)v.x;
  r.first = (int)v.z;
  r.pad = 0;
  return r;
}

// scan the L2 cells of one L1 cell (absolute L1 coords ax1..az1) that survive the box test
__device__ __forceinline__ void sweep_l1_cell(const MapView& mv, int l1, int ax1, int ay1, int az1, float qx, float qy,
                                              float qz, Top5& nb, bool seeded KS_DECL) {
  const L1Rec lr = ldg_l1(mv, l1);
  unsigned long long m2 = lr.mask;
  int base = lr.first;
  int rank = 0;
  while (m2) {
    int f2 = __ffsll((long long)m2) - 1;
    m2 &= m2 - 1;
    int ax2 = (ax1 << 2) | (f2 & 3), ay2 = (ay1 << 2) | ((f2 >> 2) & 3), az2 = (az1 << 2) | (f2 >> 4);
    KSTAT(ks.boxes++;)
    if (!(box_d2(qx, qy, qz, ax2, ay2, az2, 0.0625f) > nb.d[4])) {
      int s = mv.l2_start[base + rank];
      int e = mv.l2_start[base + rank + 1];
      scan_range(mv.sorted, s, e, qx, qy, qz, nb, seeded KS_PASS);
    }
    ++rank;
  }
}

// Row sweep at L2 resolution over the cells that intersect the ball of squared radius nb.d[4] (the
// running 5th distance, which only shrinks) around q, limited to +-R cells.  Each (z,y) row is a run
// of x cells that crosses at most R/2+2 L1 cells; inside one L1 cell a run is one contiguous range.
__device__ __forceinline__ void sweep_ball(const MapView& mv, const MapDev& md, CellCursor& cur, float qx, float qy,
                                           float qz, int ax, int ay, int az, int R, Top5& nb, bool seeded KS_DECL) {
  const int ox = md.min_c[0], oy = md.min_c[1], oz = md.min_c[2];
  for (int kz = 0; kz <= 2 * R; ++kz) {
    int dz = (kz + 1) >> 1;  // 0, +1, -1, +2, -2 ...: nearest layers first so the bound shrinks early
    if (!(kz & 1)) dz = -dz;
    int z = az + dz;
    float gz = axis_gap(qz, z, 0.0625f);
    float gz2 = gz * gz;
    if (gz2 > nb.d[4]) continue;
    for (int ky = 0; ky <= 2 * R; ++ky) {
      int dy = (ky + 1) >> 1;
      if (!(ky & 1)) dy = -dy;
      int y = ay + dy;
      float gy = axis_gap(qy, y, 0.0625f);
      float g2 = gy * gy + gz2;  // lower bound (same rounding order as a point distance's y,z terms added first)
      KSTAT(ks.boxes++;)
      if (g2 > nb.d[4]) continue;
      // x extent of the ball in this row (slightly widened; a superset is always correct)
      float rx = sqrtf(fmaxf(nb.d[4] - g2, 0.0f)) * 1.0001f + 1.0e-6f;
      int x0 = max((int)floorf((qx - rx) * 16.0f), ax - R);
      int x1 = min((int)floorf((qx + rx) * 16.0f), ax + R);
      int row2 = ((z & 3) << 4) | ((y & 3) << 2);
      int row1 = (((z >> 2) & 3) << 4) | (((y >> 2) & 3) << 2);
      int la = x0 >> 2, lb = x1 >> 2;
      for (int lx = la; lx <= lb; ++lx) {
        const CellRec* rec = find_cell(mv, md, cur, (lx >> 2) - ox, (y >> 4) - oy, (z >> 4) - oz KS_PASS);
        if (!rec) continue;
        int f1 = row1 | (lx & 3);
        unsigned long long m1 = rec->mask;
        if (!((m1 >> f1) & 1ull)) continue;
        int l1 = rec->fine_base + __popcll(m1 & ((1ull << f1) - 1ull));
        const L1Rec lr = ldg_l1(mv, l1);
        unsigned long long m2 = lr.mask;
        int lo = (lx == la) ? (x0 & 3) : 0;
        int hi = (lx == lb) ? (x1 & 3) : 3;
        int flo = row2 | lo, fhi = row2 | hi;
        unsigned long long below = (1ull << flo) - 1ull;
        unsigned long long sub = m2 & ((2ull << fhi) - 1ull) & ~below;
        if (!sub) continue;
        int b = lr.first + __popcll(m2 & below);
        scan_range(mv.sorted, mv.l2_start[b], mv.l2_start[b + __popcll(sub)], qx, qy, qz, nb, seeded KS_PASS);
      }
    }
  }
}

// exact 5-NN within squared radius 1.0; nb.id[k] = -1 for unfilled slots.
// seed (optional): five map indices believed to be close to q (the previous outer iteration's
// neighbours); they only provide the initial search radius, never the result.
__device__ __forceinline__ void knn5(const MapView& mv, float qx, float qy, float qz, Top5& nb,
                                     const float4* __restrict__ cat, const int* seed KS_DECL) {
  nb.reset();
  const MapDev md = *mv.dev;
  if (md.n <= 0) return;
  if (!(fabsf(qx) < 2.0e5f && fabsf(qy) < 2.0e5f && fabsf(qz) < 2.0e5f)) return;
  // absolute L2 cell coordinates of the query (x16 is exact in fp32)
  const int ax = (int)floorf(qx * 16.0f), ay = (int)floorf(qy * 16.0f), az = (int)floorf(qz * 16.0f);
  const int ox = md.min_c[0], oy = md.min_c[1], oz = md.min_c[2];
  CellCursor cur;
  cur.key = ~0ull;
  cur.rec = nullptr;
  float bound = 2.0f;  // > 1: no usable bound yet
  bool seeded = false;
  if (seed != nullptr && seed[4] >= 0) {
    // preload the five slots with the seeds at their distances from this query: five real points, so the running
    // 5th distance is a valid inclusive bound from the start and the sweep only inserts what beats them
    // (all lanes of a seeded launch do this in step: no divergence)
    float sd[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float4 m = __ldg(&cat[seed[k]]);
      float dx = m.x - qx, dy = m.y - qy, dz = m.z - qz;
      float r = dx * dx;
      r = r + dy * dy;
      r = r + dz * dz;
      sd[k] = r;
    }
    nb.reset_inclusive(__int_as_float(0x7f800000));
#pragma unroll
    for (int k = 0; k < 5; ++k) nb.add(sd[k], seed[k]);
    bound = nb.d[4];
    seeded = bound < 1.0f;  // a seed beyond the search radius cannot be a result: fall back to the unseeded search
    if (!seeded) nb.reset();
  }
  if (!(bound < 1.0f)) {
    // ---- A: the 27 L2 cells around the query, no prior bound
    sweep_ball(mv, md, cur, qx, qy, qz, ax, ay, az, 1, nb, false KS_PASS);
    KSTAT(ks.level = 1;)
    if (nb.full()) {
      if (nb.d[4] < 0.00390625f) return;  // 5th distance < one L2 cell: nothing outside the 27 cells can be closer
      bound = nb.d[4];
    }
  }
  if (bound < 0.0625f) {
    // ---- ball of known radius < 0.25 m: five points are known to lie within `bound` (inclusive)
    KSTAT(ks.level = 2;)
    if (!seeded) nb.reset_inclusive(bound);
    int R = (int)ceilf(sqrtf(bound) * 16.0f * 1.0001f) + 1;
    sweep_ball(mv, md, cur, qx, qy, qz, ax, ay, az, R > 5 ? 5 : R, nb, seeded KS_PASS);
    return;
  }
  // ---- sparse neighbourhood: restart, 27 L1 cells (B) then 27 L0 cells (C) with box pruning
  KSTAT(ks.level = 3;)
  if (!seeded) {
    if (bound < 1.0f)
      nb.reset_inclusive(bound);
    else
      nb.reset();
  }
  const int bx = ax >> 2, by = ay >> 2, bz = az >> 2;
  for (int k = 0; k < 27; ++k) {
    int x = bx + kNear27[k][0], y = by + kNear27[k][1], z = bz + kNear27[k][2];
    KSTAT(ks.boxes++;)
    if (box_d2(qx, qy, qz, x, y, z, 0.25f) > nb.d[4]) continue;
    const CellRec* rec = find_cell(mv, md, cur, (x >> 2) - ox, (y >> 2) - oy, (z >> 2) - oz KS_PASS);
    if (!rec) continue;
    int f1 = ((z & 3) << 4) | ((y & 3) << 2) | (x & 3);
    unsigned long long m1 = rec->mask;
    if (!((m1 >> f1) & 1ull)) continue;
    int l1 = rec->fine_base + __popcll(m1 & ((1ull << f1) - 1ull));
    sweep_l1_cell(mv, l1, x, y, z, qx, qy, qz, nb, seeded KS_PASS);
  }
  if (nb.full() && nb.d[4] < 0.0625f) return;
  KSTAT(ks.level = 4;)
  const int cx0 = bx >> 2, cy0 = by >> 2, cz0 = bz >> 2;
  for (int k = 0; k < 27; ++k) {
    int x = cx0 + kNear27[k][0], y = cy0 + kNear27[k][1], z = cz0 + kNear27[k][2];
    KSTAT(ks.boxes++;)
    if (box_d2(qx, qy, qz, x, y, z, 1.0f) > nb.d[4]) continue;
    const CellRec* rec = find_cell(mv, md, cur, x - ox, y - oy, z - oz KS_PASS);
    if (!rec) continue;
    unsigned long long m1 = rec->mask;
    int l1 = rec->fine_base;
    while (m1) {
      int f1 = __ffsll((long long)m1) - 1;
      m1 &= m1 - 1;
      int x1 = (x << 2) | (f1 & 3), y1 = (y << 2) | ((f1 >> 2) & 3), z1 = (z << 2) | (f1 >> 4);
      bool seen = (x1 >= bx - 1 && x1 <= bx + 1 && y1 >= by - 1 && y1 <= by + 1 && z1 >= bz - 1 && z1 <= bz + 1);
      KSTAT(ks.boxes++;)
      if (!seen && !(box_d2(qx, qy, qz, x1, y1, z1, 0.25f) > nb.d[4]))
        sweep_l1_cell(mv, l1, x1, y1, z1, qx, qy, qz, nb, seeded KS_PASS);
      ++l1;
    }
  }
}

}  // namespace lm
