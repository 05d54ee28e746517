// knn.cuh — device-side exact 5-NN over the local-map hash grid.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(point, 5, ...) as called by
// EdgeFeatureMatch::Match (registration/FeatureMatch/EdgeFeatureMatch.hpp:38) and
// SurfFeatureMatch::Match (surfFeatureMatch.hpp:37).  Both callers reject the
// query unless the 5th squared distance is < search_thresh_ = 1.0
// (FeatureMatchBase.hpp:29), so a grid of 1 m cells swept over the 27 cells
// around the query is exact for every query the matchers accept, and a query
// it cannot fill is one they reject.  A first pass over the 27 fine cells
// (0.25 m, a 4x4x4 occupancy mask per coarse cell) settles dense regions: it is
// exact whenever the 5th distance found is < 0.25^2, because cell indices are
// exact in fp32 for power-of-two cell sizes.
//
// Distances are FLANN's L2_Simple: ((dx*dx)+dy*dy)+dz*dz in fp32.  Results are
// ascending by (distance, original index): ties are resolved by index, which
// FLANN resolves by traversal order ("identical except at exact ties").
#pragma once
#include "common.cuh"

namespace lm {

struct MapView {
  const float4* sorted;    // cell-sorted points, .w = original index bits
  const CellRec* table;    // hash table of coarse cells
  const int* fine_start;   // start of every occupied fine cell in `sorted`, + sentinel
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
  __device__ __forceinline__ bool full() const { return id[4] >= 0; }
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

__device__ __forceinline__ const CellRec* find_cell(const MapView& mv, unsigned tmask, unsigned long long key) {
  unsigned h = hash_cell(key) & tmask;
  while (true) {
    unsigned long long k = mv.table[h].key;
    if (k == key) return &mv.table[h];
    if (k == ~0ull) return nullptr;
    h = (h + 1) & tmask;
  }
}

__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz,
                                           Top5& nb) {
  for (int p = s; p < e; ++p) {
    float4 m = __ldg(&pts[p]);
    float dx = m.x - qx, dy = m.y - qy, dz = m.z - qz;
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    nb.add(r, __float_as_int(m.w));
  }
}

// exact 5-NN within squared radius 1.0; nb.id[k] = -1 for unfilled slots
__device__ __forceinline__ void knn5(const MapView& mv, float qx, float qy, float qz, Top5& nb) {
  nb.reset();
  const MapDev md = *mv.dev;
  if (md.n <= 0) return;
  if (!(fabsf(qx) < 1.0e8f && fabsf(qy) < 1.0e8f && fabsf(qz) < 1.0e8f)) return;
  // fine coordinates relative to the grid origin (x4 is exact in fp32)
  const int gx = (int)floorf(qx * 4.0f) - 4 * md.min_c[0];
  const int gy = (int)floorf(qy * 4.0f) - 4 * md.min_c[1];
  const int gz = (int)floorf(qz * 4.0f) - 4 * md.min_c[2];
  const unsigned tmask = md.table_mask;
  // ---- pass 1: 3x3x3 fine cells; each (z,y) row is at most two runs of x sub-cells
  for (int dz = -1; dz <= 1; ++dz) {
    int z = gz + dz, cz = z >> 2;
    if (cz < 0 || cz >= md.dim[2]) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      int y = gy + dy, cy = y >> 2;
      if (cy < 0 || cy >= md.dim[1]) continue;
      int row = ((z & 3) << 4) | ((y & 3) << 2);
      int x0 = gx - 1, x1 = gx + 1;
      int ca = x0 >> 2, cb = x1 >> 2;
      for (int cx = ca; cx <= cb; ++cx) {
        if (cx < 0 || cx >= md.dim[0]) continue;
        int lo = (cx == ca) ? (x0 & 3) : 0;
        int hi = (cx == cb) ? (x1 & 3) : 3;
        const CellRec* rec = find_cell(mv, tmask, pack_cell(md, cx, cy, cz));
        if (!rec) continue;
        unsigned long long mask = rec->mask;
        int flo = row | lo, fhi = row | hi;
        unsigned long long below = (1ull << flo) - 1ull;
        unsigned long long upto = (2ull << fhi) - 1ull;
        unsigned long long sub = mask & upto & ~below;
        if (!sub) continue;
        int base = rec->fine_base + __popcll(mask & below);
        int s = mv.fine_start[base];
        int e = mv.fine_start[base + __popcll(sub)];
        scan_range(mv.sorted, s, e, qx, qy, qz, nb);
      }
    }
  }
  if (nb.full() && nb.d[4] < 0.0625f) return;
  // ---- pass 2: 3x3x3 coarse cells (exact for every query the matchers accept)
  nb.reset();
  const int cx0 = gx >> 2, cy0 = gy >> 2, cz0 = gz >> 2;
  for (int dz = -1; dz <= 1; ++dz) {
    int cz = cz0 + dz;
    if (cz < 0 || cz >= md.dim[2]) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      int cy = cy0 + dy;
      if (cy < 0 || cy >= md.dim[1]) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        int cx = cx0 + dx;
        if (cx < 0 || cx >= md.dim[0]) continue;
        const CellRec* rec = find_cell(mv, tmask, pack_cell(md, cx, cy, cz));
        if (!rec) continue;
        scan_range(mv.sorted, rec->start, rec->end, qx, qy, qz, nb);
      }
    }
  }
}

}  // namespace lm
