// knn.cuh — exact 5-NN over the local-map grid: one thread per query, in two phases per sweep.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(point, 5, ...) as called by
// EdgeFeatureMatch::Match (registration/FeatureMatch/EdgeFeatureMatch.hpp:38) and
// SurfFeatureMatch::Match (surfFeatureMatch.hpp:37).  Both callers reject the
// query unless the 5th squared distance is < search_thresh_ = 1.0
// (FeatureMatchBase.hpp:29), so only neighbours with d2 < 1.0 are ever needed.
//
// Index (mapindex.cu): L0 = 1 m cells in a hash table; every L0 cell has a 64-bit occupancy
// mask of its 4x4x4 L1 cells (0.25 m); every occupied L1 cell has a 64-bit mask
// of its 4x4x4 L2 cells (0.0625 m).  Points are grouped by (L0, L1, L2), x fastest, so any set of L2 cells that are
// consecutive among the OCCUPIED cells of one L1 cell is one contiguous range of `sorted`.
//
// Search (round 2).  Round 1 walked the grid row by row with nested loops per thread: 5-6 of 32 lanes active and a
// chain of ~75 dependent loads per query.  A first round-2 attempt (8 lanes per query, shared-memory task lists) reached
// 25 lanes per instruction but executed 20x the thread instructions.  This version keeps one thread per query and
// removes the divergence where it was made:
//   sweep    = all points in a box of L2 cells [lo, hi] minus an already searched box, in two phases:
//   phase A  cells -> segments.  The box is intersected with the occupancy masks (two ANDs with a box mask built by
//            bit arithmetic: no row or cell loops), the selected cells are cut into runs of consecutive occupied
//            cells, and every run becomes one (start, end) pair in a small per-lane list in shared memory.
//   phase B  segments -> neighbours.  ONE flat loop over the candidates of all the segments, the next point in flight
//            while the current one is judged; a candidate enters the sorted five (64-bit keys = distance bits | index)
//            only if it beats the current fifth.  The lanes of a warp differ in the trip count of this one loop only.
// With seeds (the previous outer iteration's neighbours) one sweep of the ball's bounding box is enough; without,
// the search adapts to the density around the query (the eight cells nearest to it, then a counted growth: see kq_knn5).
//
// Cell sizes are powers of two, so cell indices and cell bounds are exact in
// fp32 and a box is a superset of the ball it covers (edges computed exactly, in fp64).
// Distances are FLANN's L2_Simple: ((dx*dx)+dy*dy)+dz*dz in fp32.  Results are
// ascending by (distance, original index): ties are resolved by index, which
// FLANN resolves by traversal order ("identical except at exact ties").
//
// The search is plain per-thread code (HD): csrc/test_knn_model.cu runs the same source on the CPU against brute force.
#pragma once
#include <string.h>

#include "common.cuh"

namespace lm {

#if defined(__CUDA_ARCH__)
#define KG_POPC64(x) __popcll(x)
#define KG_POPC32(x) __popc(x)
#define KG_FFS32(x) (__ffs((int)(x)) - 1)
#define KG_LD(p) __ldg(p)
#else
#define KG_POPC64(x) __builtin_popcountll(x)
#define KG_POPC32(x) __builtin_popcount(x)
#define KG_FFS32(x) __builtin_ctz(x)
#define KG_LD(p) (*(p))
#endif

struct MapView {
  const float4* sorted;             // cell-sorted points, .w = original index bits
  const CellRec* table;             // hash table of L0 cells
  const L1Rec* l1;                  // per occupied L1 cell: occupancy of its L2 cells, index of its first L2 cell
  const int* l2_start;              // start of every occupied L2 cell in `sorted` (+ one sentinel per L0 cell)
  const MapDev* dev;
};

HD unsigned hash_cell(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (unsigned)k;
}

HD unsigned long long pack_cell(const MapDev& md, int cx, int cy, int cz) {
  return ((((unsigned long long)(unsigned)cz << md.bits[1]) | (unsigned long long)(unsigned)cy) << md.bits[0]) |
         (unsigned long long)(unsigned)cx;
}

// L0 cell at coordinates relative to the grid origin; nullptr when outside the grid or unoccupied.  The probe
// sequence is bounded: a completely full table must not hang the search.
HD const CellRec* find_cell(const MapView& mv, const MapDev& md, int cx, int cy, int cz) {
  if (cx < 0 || cy < 0 || cz < 0 || cx >= md.dim[0] || cy >= md.dim[1] || cz >= md.dim[2]) return nullptr;
  const unsigned long long key = pack_cell(md, cx, cy, cz);
  unsigned h = hash_cell(key) & md.table_mask;
  for (unsigned probes = 0; probes <= md.table_mask; ++probes) {
    const unsigned long long k = KG_LD(&mv.table[h].key);
    if (k == key) return &mv.table[h];
    if (k == ~0ull) return nullptr;
    h = (h + 1) & md.table_mask;
  }
  return nullptr;
}

// squared distance from q to the axis-aligned cell [c*s, (c+1)*s)^3, same rounding sequence as a point distance
HD float axis_gap(float q, int c, float s) {
  float lo = (float)c * s;
  float hi = lo + s;
  return (q < lo) ? (lo - q) : ((q > hi) ? (q - hi) : 0.0f);
}
HD float box_d2(float qx, float qy, float qz, int cx, int cy, int cz, float s) {
  float dx = axis_gap(qx, cx, s), dy = axis_gap(qy, cy, s), dz = axis_gap(qz, cz, s);
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

// one 16-byte load per visited L1 cell
HD L1Rec ld_l1(const MapView& mv, int l1) {
  const uint4 v = KG_LD(reinterpret_cast<const uint4*>(mv.l1 + l1));
  L1Rec r;
  r.mask = ((unsigned long long)v.y << 32) | (unsigned long long)v.x;
  r.first = (int)v.z;
  r.pad = 0;
  return r;
}

HD unsigned f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
HD float u2f(unsigned u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// a / b for 0 <= a < 4096 and the reciprocal of 1 <= b <= 128 computed once: (a + 0.5) / b is at least 0.5 / b away
// from an integer, far more than the rounding of the two float operations
HD int div_small(int a, float inv_b) { return (int)(((float)a + 0.5f) * inv_b); }

// (distance, index) as one ordered key: squared distances are >= +0, so their bit patterns order like the values
HD unsigned long long kg_key(float d2, int id) { return ((unsigned long long)f2u(d2) << 32) | (unsigned)id; }
HD float kg_key_d2(unsigned long long k) { return u2f((unsigned)(k >> 32)); }
HD int kg_key_id(unsigned long long k) { return (int)(unsigned)(k & 0xffffffffull); }

constexpr int KQ_SEG_CAP = 16;  // segments a lane collects before it scans them
constexpr int KQ_SHELL_MAX = 96;  // a grown box with more points than this is searched cell by cell, nearest first
constexpr unsigned long long KG_KEY_LT_1 = (0x3f800000ull << 32) - 1ull;  // largest key with d2 < 1.0f

// where a lane keeps its segment list: seg[(2 k) * stride], seg[(2 k + 1) * stride] = start, end of segment k
// (device: shared memory, stride 32, so that the lanes of a warp never share a bank)
struct KqList {
  int* seg;
  int stride;
};

#ifdef LMSF_KNN_STATS  // tuning builds: work counters (global atomics)
static __device__ unsigned long long g_knn_stat[16];
static unsigned long long g_knn_stat_host[16];  // the CPU model's copy
HD void kg_stat_add(int slot, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  if (v) atomicAdd(&g_knn_stat[slot], v);
#else
  g_knn_stat_host[slot] += v;
#endif
}
#define KG_STAT(slot, v) kg_stat_add(slot, (unsigned long long)(v))
#else
#define KG_STAT(slot, v)
#endif

#ifdef LMSF_KNN_CHECK  // debugging builds: index checks that record the first violation instead of faulting
static __device__ int g_knn_check[8];
HD bool kq_check_fail(int code, int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  if (atomicCAS(&g_knn_check[0], 0, code) == 0) {
    g_knn_check[1] = a;
    g_knn_check[2] = b;
    g_knn_check[3] = c;
    g_knn_check[4] = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  }
#endif
  return true;
}
#define KQ_CHECK(ok, code, a, b, c) \
  if (!(ok) && kq_check_fail(code, a, b, c)) return
#define KQ_CHECK_RET(ok, code, a, b, c, ret) \
  if (!(ok) && kq_check_fail(code, a, b, c)) return ret
#else
#define KQ_CHECK(ok, code, a, b, c)
#define KQ_CHECK_RET(ok, code, a, b, c, ret)
#endif

HD int kq_ffs64(unsigned long long v) {  // index of the lowest set bit; v != 0
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
HD unsigned long long kq_below(int b) { return (1ull << b) - 1ull; }  // bits [0, b), b < 64

// the five best keys so far, ascending; slots that hold no point yet hold the (exclusive) bound
struct KqTop {
  unsigned long long k[5];
  float d4;  // distance part of k[4]: a candidate farther than this cannot enter
  int n;     // points among them
};
HD void kq_top_reset(KqTop& t, unsigned long long exclusive_bound) {
#pragma unroll
  for (int i = 0; i < 5; ++i) t.k[i] = exclusive_bound;
  t.d4 = kg_key_d2(exclusive_bound);
  t.n = 0;
}
// key < t.k[4].  dedup: the sweep may meet points that are in the list already (sparse path only)
HD void kq_insert(KqTop& t, unsigned long long key, bool dedup) {
  if (dedup && (key == t.k[0] || key == t.k[1] || key == t.k[2] || key == t.k[3])) return;
  t.n = t.n < 5 ? t.n + 1 : 5;
  // the new key takes the place of the 5th and sinks; most keys stop after a step or two
  if (key < t.k[3]) {
    t.k[4] = t.k[3];
    if (key < t.k[2]) {
      t.k[3] = t.k[2];
      if (key < t.k[1]) {
        t.k[2] = t.k[1];
        if (key < t.k[0]) {
          t.k[1] = t.k[0];
          t.k[0] = key;
        } else {
          t.k[1] = key;
        }
      } else {
        t.k[2] = key;
      }
    } else {
      t.k[3] = key;
    }
  } else {
    t.k[4] = key;
  }
  t.d4 = kg_key_d2(t.k[4]);
}

// bits (x + 4 y + 16 z) of a 4x4x4 block whose coordinates lie in [xl, xh] x [yl, yh] x [zl, zh]; the ranges are
// clamped to the block, an empty range gives 0
HD unsigned long long kq_box_mask(int xl, int xh, int yl, int yh, int zl, int zh) {
  xl = xl < 0 ? 0 : xl, yl = yl < 0 ? 0 : yl, zl = zl < 0 ? 0 : zl;
  xh = xh > 3 ? 3 : xh, yh = yh > 3 ? 3 : yh, zh = zh > 3 ? 3 : zh;
  if (xl > xh || yl > yh || zl > zh) return 0ull;
  const unsigned xb = ((2u << xh) - 1u) & ~((1u << xl) - 1u);                 // x range, 4 bits
  const unsigned yb = ((2u << (4 * yh + 3)) - 1u) & ~((1u << (4 * yl)) - 1u);  // rows yl..yh of one layer, 16 bits
  const unsigned w = (xb * 0x11111111u) & (yb * 0x00010001u);                  // two layers
  // layers zl..zh: two per 32-bit half
  const unsigned zb = ((2u << zh) - 1u) & ~((1u << zl) - 1u);  // 4 bits
  const unsigned lo = w & ((zb & 1u ? 0x0000ffffu : 0u) | (zb & 2u ? 0xffff0000u : 0u));
  const unsigned hi = w & ((zb & 4u ? 0x0000ffffu : 0u) | (zb & 8u ? 0xffff0000u : 0u));
  return ((unsigned long long)hi << 32) | lo;
}

// phase B: one flat loop over the `total` candidates of segments [first, ...) of the list, KQ_DEPTH loads ahead of
// the judgement
HD void kq_scan(const MapView& mv, const KqList& li, int first, int total, float qx, float qy, float qz, KqTop& top,
                bool dedup) {
  if (total <= 0) return;
  int at = (2 * first) * li.stride;  // the fetch cursor: list position of the current segment, next point, its end
  int p = li.seg[at], e = li.seg[at + li.stride];
  int fetched = 0;
  float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0;
#define KQ_FETCH(dst)                                               \
  if (fetched < total) {                                            \
    if (p == e) {                                                   \
      at += 2 * li.stride;                                          \
      p = li.seg[at];                                               \
      e = li.seg[at + li.stride];                                   \
    }                                                               \
    KQ_CHECK(p >= 0 && p < e && e <= mv.dev->n, 2, p, e, fetched);  \
    dst = KG_LD(&mv.sorted[p]);                                     \
    ++p;                                                            \
    ++fetched;                                                      \
  }
#ifndef KQ_DEPTH
#define KQ_DEPTH 2  // candidate loads in flight per lane (3 and 4 measured: no gain, r2l)
#endif
  float4 n2 = n0, n3 = n0;
  KQ_FETCH(n0);
  KQ_FETCH(n1);
  if (KQ_DEPTH > 2) {
    KQ_FETCH(n2);
  }
  if (KQ_DEPTH > 3) {
    KQ_FETCH(n3);
  }
  for (int i = 0; i < total; ++i) {
    const float4 m = n0;
    n0 = n1;
    if (KQ_DEPTH > 2) n1 = n2;
    if (KQ_DEPTH > 3) n2 = n3;
    if (KQ_DEPTH == 2) {
      KQ_FETCH(n1);
    } else if (KQ_DEPTH == 3) {
      KQ_FETCH(n2);
    } else {
      KQ_FETCH(n3);
    }
    const float dx = m.x - qx, dy = m.y - qy, dz = m.z - qz;
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    KG_STAT(2, 1);
    if (r <= top.d4) {  // cheap rejection first; ties and order are settled on the full key
      const unsigned long long key = kg_key(r, (int)f2u(m.w));
      if (key < top.k[4]) kq_insert(top, key, dedup);
    }
  }
#undef KQ_FETCH
}

// what one sweep does with the cells of its box
// KQ_SCAN_FEW: count, and keep the segments as long as they fit the list and hold at most KQ_SHELL_MAX points (few = true
// on return: the caller scans them; otherwise the sweep was a count)
enum { KQ_SCAN = 0, KQ_COUNT = 1, KQ_NEAREST_L1 = 2, KQ_SCAN_FEW = 3 };

struct KqSweep {
  int lo[3], hi[3];    // box of L2 cells (inclusive)
  int xlo[3], xhi[3];  // cells searched before: left out (none when xlo[0] > xhi[0])
  float bound_d;       // L1 cells farther than this (squared) are skipped
  int mode;
  // KQ_NEAREST_L1: the occupied L1 cell of the box nearest to the query among those behind (after_d, after_id)
  float after_d;
  int after_id;
  // results
  int count;           // points in the swept cells
  bool few;            // KQ_SCAN_FEW: the segments of all `count` points are in the list
  float best_d;        // KQ_NEAREST_L1: squared distance, id and L2 origin of the cell found (best_id < 0: none)
  int best_id, best_o[3];
};

// Phase A of one sweep (see the header).  KQ_SCAN: segments to the list; scanning is left to the caller (kq_scan over
// the list) as long as the list does not overflow.  Returns the number of list entries that are waiting to be scanned.
HD int kq_collect(const MapView& mv, const MapDev& md, const KqList& li, float qx, float qy, float qz, KqSweep& sw,
                  KqTop& top, bool dedup, int& pending) {
  int nseg = 0;
  pending = 0;
  sw.count = 0;
  sw.few = true;
  sw.best_id = -1;
  sw.best_d = 3.0e38f;
  const bool excl = sw.xlo[0] <= sw.xhi[0];
  for (int cz = sw.lo[2] >> 4; cz <= (sw.hi[2] >> 4); ++cz)
    for (int cy = sw.lo[1] >> 4; cy <= (sw.hi[1] >> 4); ++cy)
      for (int cx = sw.lo[0] >> 4; cx <= (sw.hi[0] >> 4); ++cx) {
        KG_STAT(0, 1);
        const CellRec* rec = find_cell(mv, md, cx - md.min_c[0], cy - md.min_c[1], cz - md.min_c[2]);
        if (!rec) continue;
        const unsigned long long m1 = KG_LD(&rec->mask);
        const int fine_base = KG_LD(&rec->fine_base);
        // L1 cells of this L0 cell that the box touches, without those that lie inside the excluded box
        unsigned long long sel1 = m1 & kq_box_mask((sw.lo[0] >> 2) - (cx << 2), (sw.hi[0] >> 2) - (cx << 2),
                                                   (sw.lo[1] >> 2) - (cy << 2), (sw.hi[1] >> 2) - (cy << 2),
                                                   (sw.lo[2] >> 2) - (cz << 2), (sw.hi[2] >> 2) - (cz << 2));
        if (excl)
          sel1 &= ~kq_box_mask(((sw.xlo[0] + 3) >> 2) - (cx << 2), ((sw.xhi[0] + 1) >> 2) - 1 - (cx << 2),
                               ((sw.xlo[1] + 3) >> 2) - (cy << 2), ((sw.xhi[1] + 1) >> 2) - 1 - (cy << 2),
                               ((sw.xlo[2] + 3) >> 2) - (cz << 2), ((sw.xhi[2] + 1) >> 2) - 1 - (cz << 2));
        while (sel1) {
          const int f1 = kq_ffs64(sel1);
          sel1 &= sel1 - 1ull;
          // origin of the L1 cell in L2 coordinates
          const int ox = ((cx << 2) | (f1 & 3)) << 2, oy = ((cy << 2) | ((f1 >> 2) & 3)) << 2,
                    oz = ((cz << 2) | (f1 >> 4)) << 2;
          const float cd = box_d2(qx, qy, qz, ox >> 2, oy >> 2, oz >> 2, 0.25f);
          if (cd > sw.bound_d) continue;
          if (sw.mode == KQ_NEAREST_L1) {
            // (distance, id) orders the cells; id = position in a 16^3 block of L1 cells around the box corner
            const int id = ((((oz >> 2) - (sw.lo[2] >> 2)) & 15) << 8) | ((((oy >> 2) - (sw.lo[1] >> 2)) & 15) << 4) |
                           (((ox >> 2) - (sw.lo[0] >> 2)) & 15);
            const bool behind = cd > sw.after_d || (cd == sw.after_d && id > sw.after_id);
            const bool better = cd < sw.best_d || (cd == sw.best_d && id < sw.best_id);
            if (behind && better) {
              sw.best_d = cd;
              sw.best_id = id;
              sw.best_o[0] = ox, sw.best_o[1] = oy, sw.best_o[2] = oz;
            }
            continue;
          }
          KG_STAT(1, 1);
          KQ_CHECK_RET(fine_base >= 0 && fine_base <= md.n, 3, fine_base, f1, md.n, 0);
          const L1Rec lr = ld_l1(mv, fine_base + KG_POPC64(m1 & kq_below(f1)));
          unsigned long long sel2 = lr.mask & kq_box_mask(sw.lo[0] - ox, sw.hi[0] - ox, sw.lo[1] - oy, sw.hi[1] - oy,
                                                          sw.lo[2] - oz, sw.hi[2] - oz);
          if (excl)
            sel2 &= ~kq_box_mask(sw.xlo[0] - ox, sw.xhi[0] - ox, sw.xlo[1] - oy, sw.xhi[1] - oy, sw.xlo[2] - oz,
                                 sw.xhi[2] - oz);
          while (sel2) {
            // a run: selected cells with no occupied unselected cell between them are contiguous in `sorted`
            const int lb = kq_ffs64(sel2);
            const unsigned long long gap = lr.mask & ~sel2 & ~kq_below(lb);
            const unsigned long long run = gap ? (sel2 & kq_below(kq_ffs64(gap))) : sel2;
            sel2 &= ~run;
            const int b0 = lr.first + KG_POPC64(lr.mask & kq_below(lb));
            KQ_CHECK_RET(b0 >= 0 && b0 + KG_POPC64(run) <= 2 * md.n, 4, b0, lr.first, (int)KG_POPC64(run), 0);
            const int s = KG_LD(&mv.l2_start[b0]);
            const int e = KG_LD(&mv.l2_start[b0 + KG_POPC64(run)]);
            KQ_CHECK_RET(s >= 0 && s < e && e <= md.n, 5, s, e, b0, 0);
            sw.count += e - s;
            KG_STAT(3, 1);
            if (sw.mode == KQ_SCAN_FEW) {
              if (nseg == KQ_SEG_CAP || sw.count > KQ_SHELL_MAX) sw.few = false;
              if (!sw.few) continue;
            }
            if (sw.mode == KQ_SCAN || sw.mode == KQ_SCAN_FEW) {
              if (nseg == KQ_SEG_CAP) {  // rare: a long list is scanned in pieces
                kq_scan(mv, li, 0, pending, qx, qy, qz, top, dedup);
                nseg = 0;
                pending = 0;
              }
              li.seg[(2 * nseg) * li.stride] = s;
              li.seg[(2 * nseg + 1) * li.stride] = e;
              ++nseg;
              pending += e - s;
            }
          }
        }
      }
  return nseg;
}

// L2 cells of the bounding box of the ball of squared radius d2 around q: exact edges (fp64), radius slightly widened
HD void kq_ball_box(float qx, float qy, float qz, float d2, int lo[3], int hi[3]) {
  const double rad = (double)(sqrtf(d2) * 1.0001f + 1.0e-6f) * 16.0;
  const double c[3] = {(double)qx * 16.0, (double)qy * 16.0, (double)qz * 16.0};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    lo[k] = (int)floor(c[k] - rad);
    hi[k] = (int)floor(c[k] + rad);
  }
}

// everything outside the box [lo, hi] of L2 cells is at least sqrt(result) away from q (differences rounded like a
// point's coordinate differences: a comparison with a point's distance is safe under monotonic rounding)
HD float kq_gap2(const float q[3], const int lo[3], const int hi[3]) {
  float gap = 3.0e38f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float g0 = q[k] - (float)lo[k] * 0.0625f, g1 = (float)(hi[k] + 1) * 0.0625f - q[k];
    gap = fminf(gap, fminf(g0, g1));
  }
  return gap > 0.0f ? gap * gap : 0.0f;
}

enum {
  KQ_ST_START = 0,   // the 2 x 2 x 2 block of L2 cells nearest to the query
  KQ_ST_BALL = 1,    // the ball of the known bound, minus what was searched
  KQ_ST_GROW = 2,    // sparse: count a larger box
  KQ_ST_SHELL = 3,   // sparse: scan the box that holds five points (few points)
  KQ_ST_FIND = 4,    // sparse: the box holds many points — find the nearest L1 cell not scanned yet
  KQ_ST_CELL = 5,    // sparse: scan that L1 cell
  KQ_ST_LAST = 6,    // sparse: the ball of the bound, everything (duplicates filtered)
  KQ_ST_DONE = 7
};

// Exact 5-NN within squared radius 1.0 of one query.  seed = nullptr: no prior knowledge; otherwise the keys of five
// distinct map points (any order) believed to be near the query.  Returns n <= 5 and the n nearest keys,
// ascending, in top.k[0, n); n < 5: fewer than five points within the radius (the callers reject such a query).
//
// Every step is one sweep (phase A + phase B) at ONE call site, so that the lanes of a warp stay together whatever
// state their queries are in.  Without a seed the search adapts to the density around the query:
//   START  the 2 x 2 x 2 block of L2 cells nearest to the query is collected and scanned (every face of the block is at
//          least half a cell away from the query).  Five points in hand and nothing unsearched nearer than the 5th: done.
//          (Round 2 until r2t: the 27 cells around the own cell, or the own cell alone when it held five points — three
//          times the candidates, or a second sweep for most queries; the block is 12 % faster per pass.)
//   BALL   else what the ball of the 5th distance holds beyond the searched box settles the answer.
//   sparse (fewer than five points in the block): the box grows — points are only COUNTED — until it holds five.
//          Few points: they are scanned (SHELL) and BALL finishes.  Many (a dense surface has entered the box, which a
//          box does with a whole face): the L1 cells of the box are scanned nearest first (FIND, CELL) until five
//          points are known, and the ball of that bound is swept once more with duplicates filtered (LAST).
//
// DEFER (the registration kernel): the sparse cases that cost a single thread a long chain of dependent loads —
// no five points within 4 cells, or a dense surface inside the grown box — are not searched here: the function returns
// -1 and the caller hands the query to the warp-cooperative search (kw_knn5).
template <bool DEFER>
HD int kq_knn5(const MapView& mv, const KqList& li, float qx, float qy, float qz, const unsigned long long* seed,
               KqTop& top) {
  const MapDev md = *mv.dev;
  kq_top_reset(top, KG_KEY_LT_1 + 1ull);
  if (!(md.n > 0) || !(fabsf(qx) < 2.0e5f && fabsf(qy) < 2.0e5f && fabsf(qz) < 2.0e5f)) return 0;
  // absolute L2 cell coordinates of the query (x16 is exact in fp32)
  const int a[3] = {(int)floorf(qx * 16.0f), (int)floorf(qy * 16.0f), (int)floorf(qz * 16.0f)};
  const float q[3] = {qx, qy, qz};
  KqSweep sw;
  sw.xlo[0] = 1, sw.xhi[0] = 0;  // nothing searched yet
  sw.xlo[1] = sw.xlo[2] = sw.xhi[1] = sw.xhi[2] = 0;
  sw.bound_d = 1.0f;
  sw.mode = KQ_SCAN;
  sw.after_d = -1.0f;
  sw.after_id = -1;
  int state;
  bool dedup = false;
  int total = 0;  // points in the boxes swept so far
  int grow = 0;   // sparse: next box width to try
  int glo[3] = {0, 0, 0}, ghi[3] = {0, 0, 0};  // sparse: the box that holds five points
  // START's box: the 2 x 2 x 2 block of L2 cells nearest to the query — every face of it is at least half a cell
  // (31 mm) away, and it holds a third of the candidates of the 27 cells around the query's own cell
  int slo[3], shi[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int up = (q[k] * 16.0f - (float)a[k]) >= 0.5f ? 1 : 0;  // exact: x16 and the difference are exact in fp32
    slo[k] = a[k] - 1 + up, shi[k] = a[k] + up;
  }
  bool seeded = false;
  if (seed != nullptr) {
    // the five seeds ARE the list to beat: the sweep of their ball only inserts what is nearer than one of them (between
    // two outer iterations the pose moves by millimetres and hardly anything is); the sweep meets the seeds again
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (seed[k] < top.k[4]) kq_insert(top, seed[k], false);
    seeded = top.n == 5;  // a seed beyond the search radius is no bound
    if (!seeded) kq_top_reset(top, KG_KEY_LT_1 + 1ull);
  }
  if (seeded) {
    state = KQ_ST_BALL;
    dedup = true;
    sw.bound_d = top.d4;
    kq_ball_box(qx, qy, qz, sw.bound_d, sw.lo, sw.hi);
  } else {
    state = KQ_ST_START;
#pragma unroll
    for (int k = 0; k < 3; ++k) sw.lo[k] = slo[k], sw.hi[k] = shi[k];
  }
  while (state != KQ_ST_DONE) {
    KG_STAT(8 + state, 1);
    int pending = 0;
    const int nseg = kq_collect(mv, md, li, qx, qy, qz, sw, top, dedup, pending);
    if (sw.mode == KQ_SCAN || (sw.mode == KQ_SCAN_FEW && sw.few)) {
      if (nseg > 0) kq_scan(mv, li, 0, pending, qx, qy, qz, top, dedup);
      if (state == KQ_ST_START) {  // searched: START's box
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.xlo[k] = slo[k], sw.xhi[k] = shi[k];
        total = sw.count;
      }
    }
    // ---- transitions
    if (state == KQ_ST_BALL || state == KQ_ST_LAST) {
      state = KQ_ST_DONE;  // exhaustive
    } else if (state == KQ_ST_START || state == KQ_ST_SHELL) {
      if (state == KQ_ST_SHELL) {
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.xlo[k] = sw.lo[k], sw.xhi[k] = sw.hi[k];
      }
      if (top.n == 5) {
        if (top.d4 < kq_gap2(q, sw.xlo, sw.xhi)) {
          state = KQ_ST_DONE;  // nothing outside the searched box can be nearer than the 5th
        } else {
          state = KQ_ST_BALL;
          sw.mode = KQ_SCAN;
          sw.bound_d = top.d4;
          kq_ball_box(qx, qy, qz, top.d4, sw.lo, sw.hi);
        }
      } else if (state == KQ_ST_START) {
        state = KQ_ST_GROW;
        sw.mode = KQ_SCAN_FEW;  // the first grown box is scanned in the same sweep when it holds few points
        grow = md.grow0 > 2 ? (md.grow0 > 16 ? 16 : md.grow0) : 2;
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.lo[k] = a[k] - grow, sw.hi[k] = a[k] + grow;
      } else {
        // the shell was scanned and still fewer than five points are within the radius 1.0: cannot happen while the
        // counted points lie within it; a box that reaches beyond the radius can hold points that do not count
        state = KQ_ST_LAST;
        dedup = true;
        sw.mode = KQ_SCAN;
        sw.xlo[0] = 1, sw.xhi[0] = 0;
        sw.bound_d = 1.0f;
        kq_ball_box(qx, qy, qz, 1.0f, sw.lo, sw.hi);
      }
    } else if (state == KQ_ST_GROW) {
      total += sw.count;
      const bool scanned = sw.mode == KQ_SCAN_FEW && sw.few;
      sw.mode = KQ_COUNT;
      if (scanned) {  // the grown box is searched: it takes the place of START's block
#pragma unroll
        for (int k = 0; k < 3; ++k) slo[k] = sw.lo[k], shi[k] = sw.hi[k];
      }
      if (DEFER && ((total < 5 && grow >= 4) || total - top.n > KQ_SHELL_MAX)) return -1;
      if (scanned && top.n == 5) {
        // as after SHELL: the ball of the 5th distance beyond the searched box settles the answer
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.xlo[k] = slo[k], sw.xhi[k] = shi[k];
        if (top.d4 < kq_gap2(q, sw.xlo, sw.xhi)) {
          state = KQ_ST_DONE;
        } else {
          state = KQ_ST_BALL;
          sw.mode = KQ_SCAN;
          sw.bound_d = top.d4;
          kq_ball_box(qx, qy, qz, top.d4, sw.lo, sw.hi);
        }
      } else if (!scanned && (total >= 5 || grow >= 16)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) glo[k] = sw.lo[k], ghi[k] = sw.hi[k];
        // the counted box becomes the searched one after the next step; what START searched stays left out
        if (total - top.n <= KQ_SHELL_MAX) {
          state = KQ_ST_SHELL;
          sw.mode = KQ_SCAN;  // box = the counted box, minus START's
        } else {
          state = KQ_ST_FIND;
          sw.mode = KQ_NEAREST_L1;
          sw.xlo[0] = 1, sw.xhi[0] = 0;  // whole L1 cells from here on, duplicates filtered
          dedup = true;
          sw.after_d = -1.0f;
          sw.after_id = -1;
          if (DEFER) return -1;  // not reached (deferred above): lets the compiler drop the cell-by-cell states
        }
      } else {
        // next width: 2, 3, 4, 6, 8, 12, 16; the box counted so far is left out of the next count
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.xlo[k] = sw.lo[k], sw.xhi[k] = sw.hi[k];
        grow = grow < 4 ? grow + 1 : (grow < 6 ? 6 : (grow < 8 ? 8 : (grow < 12 ? 12 : 16)));
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.lo[k] = a[k] - grow, sw.hi[k] = a[k] + grow;
      }
      if (state == KQ_ST_SHELL) {
        // leave out what START searched (its box), not the boxes that were only counted
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.xlo[k] = slo[k], sw.xhi[k] = shi[k];
      }
    } else if (DEFER) {
      return -1;  // not reached
    } else if (state == KQ_ST_FIND) {
      if (sw.best_id < 0) {
        // no cell left in the box: sweep the search radius
        state = KQ_ST_LAST;
        sw.mode = KQ_SCAN;
        sw.bound_d = top.n == 5 ? top.d4 : 1.0f;
        kq_ball_box(qx, qy, qz, sw.bound_d, sw.lo, sw.hi);
      } else {
        state = KQ_ST_CELL;
        sw.mode = KQ_SCAN;
        sw.after_d = sw.best_d;
        sw.after_id = sw.best_id;
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.lo[k] = sw.best_o[k], sw.hi[k] = sw.best_o[k] + 3;
      }
    } else {  // KQ_ST_CELL
      if (top.n == 5) {
        state = KQ_ST_LAST;
        sw.mode = KQ_SCAN;
        sw.bound_d = top.d4;
        kq_ball_box(qx, qy, qz, top.d4, sw.lo, sw.hi);
      } else {
        state = KQ_ST_FIND;
        sw.mode = KQ_NEAREST_L1;
#pragma unroll
        for (int k = 0; k < 3; ++k) sw.lo[k] = glo[k], sw.hi[k] = ghi[k];
      }
    }
  }
  return top.n;
}
// The seeded search alone (outer iterations after the first, registration kernel): the five seeds are the list to beat and
// one sweep of the bounding box of their ball settles the answer — no START / GROW states, so the kernel that calls only
// this needs fewer registers.  Returns top.n == 5, or -1 when a seed lies beyond the search radius (no bound: the caller
// hands the query to the warp-cooperative search).
HD int kq_knn5_seeded(const MapView& mv, const KqList& li, float qx, float qy, float qz, const unsigned long long* seed,
                      KqTop& top) {
  const MapDev md = *mv.dev;
  kq_top_reset(top, KG_KEY_LT_1 + 1ull);
  if (!(md.n > 0) || !(fabsf(qx) < 2.0e5f && fabsf(qy) < 2.0e5f && fabsf(qz) < 2.0e5f)) return 0;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (seed[k] < top.k[4]) kq_insert(top, seed[k], false);
  if (top.n != 5) return -1;
  KqSweep sw;
  sw.xlo[0] = 1, sw.xhi[0] = 0;  // nothing left out
  sw.xlo[1] = sw.xlo[2] = sw.xhi[1] = sw.xhi[2] = 0;
  sw.bound_d = top.d4;
  sw.mode = KQ_SCAN;
  sw.after_d = -1.0f;
  sw.after_id = -1;
  kq_ball_box(qx, qy, qz, sw.bound_d, sw.lo, sw.hi);
  int pending = 0;
  const int nseg = kq_collect(mv, md, li, qx, qy, qz, sw, top, true, pending);
  if (nseg > 0) kq_scan(mv, li, 0, pending, qx, qy, qz, top, true);
  return top.n;
}

// ---- the sparse cases, searched by a whole warp -------------------------------------------------------------------
// One query per warp: best-first over the occupied L1 cells within the search radius.  Every load a single thread
// would chain (27 hash probes, one L1 record and two cell starts per L1 cell) is issued by its own lane.
//   1  lanes 0..26 probe the 27 L0 cells the ball of radius 1 can touch;
//   2  the occupied L1 cells of those cells become items in shared memory (prefix sum over the lanes);
//   3  every lane takes items: distance of the cell's box to the query, and — if that is inside the radius — the
//      cell's point range (L1 record, first and last L2 start);
//   4  until no unscanned item is nearer than the 5th neighbour so far: the nearest item (warp minimum) is scanned,
//      32 points per step; candidates that beat the 5th key are inserted by all lanes alike (every lane holds the same
//      five keys).
// Exact: a cell's box distance never exceeds the distance of a point inside it (knn.cuh header).
constexpr int KW_ITEM_CAP = 512;
struct KwScratch {  // per warp
  float cd[KW_ITEM_CAP];  // squared box distance of the item's L1 cell; KW_NONE: scanned or outside the radius
  int a[KW_ITEM_CAP];     // step 2: L1 record index; step 3 on: first point
  int b[KW_ITEM_CAP];     // step 2: packed cell coordinates; step 3 on: number of points
};
constexpr float KW_NONE = 3.0e38f;

#if defined(__CUDACC__)
struct Warp32 {  // executor of the device
  int lane;
  __device__ Warp32() : lane(threadIdx.x & 31) {}
  __device__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ int shfl_up(int v, int d) const { return __shfl_up_sync(0xffffffffu, v, d); }
  __device__ unsigned long long shfl64(unsigned long long v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ unsigned long long shfl_xor64(unsigned long long v, int d) const {
    return __shfl_xor_sync(0xffffffffu, v, d);
  }
  __device__ void sync() const { __syncwarp(); }
};
#endif

// all 32 lanes call with the same arguments; on return every lane holds the same `top`.  Returns top.n, or -1 when the
// surroundings hold more occupied L1 cells than the item list (the caller falls back to the per-thread search).
template <class X>
HD int kw_knn5(const X& x, const MapView& mv, KwScratch* s, float qx, float qy, float qz, KqTop& top) {
  const MapDev md = *mv.dev;
  kq_top_reset(top, KG_KEY_LT_1 + 1ull);
  if (!(md.n > 0) || !(fabsf(qx) < 2.0e5f && fabsf(qy) < 2.0e5f && fabsf(qz) < 2.0e5f)) return 0;
  const int c0[3] = {(int)floorf(qx), (int)floorf(qy), (int)floorf(qz)};
  // 1: one L0 cell per lane
  unsigned long long m1 = 0ull;
  int fine_base = 0;
  const int dx = x.lane % 3 - 1, dy = (x.lane / 3) % 3 - 1, dz = x.lane / 9 - 1;
  if (x.lane < 27) {
    const CellRec* rec =
        find_cell(mv, md, c0[0] + dx - md.min_c[0], c0[1] + dy - md.min_c[1], c0[2] + dz - md.min_c[2]);
    if (rec) {
      m1 = KG_LD(&rec->mask);
      fine_base = KG_LD(&rec->fine_base);
    }
  }
  // 2: items
  const int cnt = KG_POPC64(m1);
  int inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = x.shfl_up(inc, d);
    if (x.lane >= d) inc += t;
  }
  const int total = x.shfl(inc, 31);
  if (total > KW_ITEM_CAP) return -1;
  {
    int at = inc - cnt, r = 0;
    unsigned long long m = m1;
    while (m) {
      const int f1 = kq_ffs64(m);
      m &= m - 1ull;
      s->a[at] = fine_base + r;
      // L1 coordinates relative to the corner of the 3x3x3 block of L0 cells: 0..11 per axis
      s->b[at] = ((((dz + 1) << 2) | (f1 >> 4)) << 8) | ((((dy + 1) << 2) | ((f1 >> 2) & 3)) << 4) |
                 (((dx + 1) << 2) | (f1 & 3));
      ++at;
      ++r;
    }
  }
  x.sync();
  // 3: distance and point range of every item
  for (int k = x.lane; k < total; k += 32) {
    const int pk = s->b[k];
    const int lx = ((c0[0] - 1) << 2) + (pk & 15), ly = ((c0[1] - 1) << 2) + ((pk >> 4) & 15),
              lz = ((c0[2] - 1) << 2) + (pk >> 8);
    float cd = box_d2(qx, qy, qz, lx, ly, lz, 0.25f);
    int ps = 0, np = 0;
    if (cd < 1.0f) {
      const L1Rec lr = ld_l1(mv, s->a[k]);
      ps = KG_LD(&mv.l2_start[lr.first]);
      np = KG_LD(&mv.l2_start[lr.first + KG_POPC64(lr.mask)]) - ps;
    } else {
      cd = KW_NONE;
    }
    s->cd[k] = cd;
    s->a[k] = ps;
    s->b[k] = np;
  }
  x.sync();
  // 4: best first
  while (true) {
    unsigned long long best = ~0ull;  // (distance bits, item): box distances are >= 0
    for (int k = x.lane; k < total; k += 32) {
      const float cd = s->cd[k];
      if (cd != KW_NONE) {
        const unsigned long long key = kg_key(cd, k);
        best = key < best ? key : best;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = x.shfl_xor64(best, d);
      best = o < best ? o : best;
    }
    if (best == ~0ull) break;
    const float cd = kg_key_d2(best);
    if (top.n == 5 && cd > top.d4) break;  // no unscanned cell can hold a point that beats the 5th
    const int k = kg_key_id(best);
    const int ps = s->a[k], np = s->b[k];
    x.sync();  // every lane has read the item
    if (x.lane == 0) s->cd[k] = KW_NONE;
    for (int p0 = 0; p0 < np; p0 += 32) {
      const int p = p0 + x.lane;
      unsigned long long key = ~0ull;
      if (p < np) {
        const float4 m = KG_LD(&mv.sorted[ps + p]);
        const float ex = m.x - qx, ey = m.y - qy, ez = m.z - qz;
        float r = ex * ex;
        r = r + ey * ey;
        r = r + ez * ez;
        key = kg_key(r, (int)f2u(m.w));
      }
      unsigned pass = x.ballot(key < top.k[4]);
      while (pass) {  // the same sequence of inserts in every lane
        const int src = KG_FFS32(pass);
        pass &= pass - 1u;
        const unsigned long long kk = x.shfl64(key, src);
        if (kk < top.k[4]) kq_insert(top, kk, false);
      }
    }
    x.sync();
  }
  return top.n;
}

// five neighbours of one query as the fits read them (match.cu): ascending by (distance, index)
struct Top5 {
  float d[5];
  int id[5];
  HD bool full() const { return id[4] >= 0; }
};

// kernels that search: blocks of KG_BLOCK threads, one query per thread, one segment list per lane in shared memory
constexpr int KG_BLOCK = 128;
constexpr int KQ_SMEM_INTS = KG_BLOCK * 2 * KQ_SEG_CAP;  // per block
#if defined(__CUDACC__)
// the calling lane's list inside the block's array s_seg[KQ_SMEM_INTS]
__device__ __forceinline__ KqList kq_list(int* s_seg) {
  KqList li;
  li.seg = s_seg + (threadIdx.x >> 5) * (32 * 2 * KQ_SEG_CAP) + (threadIdx.x & 31);
  li.stride = 32;
  return li;
}
#endif

}  // namespace lm
