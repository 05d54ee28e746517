// knn.cuh — exact 5-NN over the local-map grid, searched COOPERATIVELY by a group of lanes per query.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch(point, 5, ...) as called by
// EdgeFeatureMatch::Match (registration/FeatureMatch/EdgeFeatureMatch.hpp:38) and
// SurfFeatureMatch::Match (surfFeatureMatch.hpp:37).  Both callers reject the
// query unless the 5th squared distance is < search_thresh_ = 1.0
// (FeatureMatchBase.hpp:29), so only neighbours with d2 < 1.0 are ever needed.
//
// Index (mapindex.cu): L0 = 1 m cells in a hash table; every L0 cell has a 64-bit occupancy
// mask of its 4x4x4 L1 cells (0.25 m); every occupied L1 cell has a 64-bit mask
// of its 4x4x4 L2 cells (0.0625 m).  Points are grouped by (L0, L1, L2), x fastest, so an x-run of
// L2 cells inside one L1 cell — and a whole L1 cell — is one contiguous range of `sorted`.
//
// Search (round 2; round 1 walked the grid with one thread per query: 5-6 of 32 lanes active, 19-31 % of the
// issue slots in a per-thread sorted insert).  A query is owned by a GROUP of G = 8 lanes, four queries per warp:
//   tasks       the cells of a box around the query are cut into tasks — (z,y) rows of L2 cells per L1 cell
//               ("rows" mode, boxes up to 11 cells wide) or whole L1 cells ("cells" mode, larger boxes) — dealt
//               round-robin to the lanes; a task ends in one contiguous segment [s, e) of `sorted` (three dependent
//               loads: hash probe, L1 record, two L2 starts).  With a bound on the 5th distance, rows outside the
//               ball are skipped and the x-run is clipped to the ball.  Non-empty segments are compacted into a
//               per-group list in shared memory.
//   candidates  the segments are flattened (prefix sums) and the group strides over the candidates with coalesced
//               16-byte loads: every lane computes one distance per step, whatever the cell occupancy is.
//   accept      a candidate whose key (distance bits, index) is <= the bound key is appended to the group's
//               accepted list (ballot + popc: no sorted insert, no divergence); when the list fills up it is
//               compressed to its five smallest keys and the bound tightens.
//   select      the five smallest keys of the list, ascending: rank counting for lists of up to G keys, five rounds
//               of group-min extraction otherwise.
// A query runs through a small state machine: with seeds (the previous outer iteration's neighbours) one bounded
// sweep is enough; without, the 27 L2 cells around the query give the bound (or the answer when the 5th distance
// is below one cell), then one bounded sweep; sparse surroundings escalate to the 27 L1 cells and to everything
// within 1 m.  All groups of a warp step through the machine in lockstep.
//
// Cell sizes are powers of two, so cell indices and cell bounds are exact in
// fp32 and the box distance — computed with the same rounding sequence as a
// point distance — never exceeds the distance of a point inside the box:
// pruning is exact, ties included.
//
// Distances are FLANN's L2_Simple: ((dx*dx)+dy*dy)+dz*dz in fp32.  Results are
// ascending by (distance, original index): ties are resolved by index, which
// FLANN resolves by traversal order ("identical except at exact ties").
//
// The search is written once, as templates over an executor X (lane id inside the group, ballot / shuffle /
// any): WarpGroup8 on the device, HostGroup1 (one lane) in csrc/test_knn_model.cu, which runs the same source on the
// CPU against brute force.
#pragma once
#include <string.h>

#include "common.cuh"

namespace lm {

#if defined(__CUDA_ARCH__)
#define KG_POPC64(x) __popcll(x)
#define KG_POPC32(x) __popc(x)
#define KG_LD(p) __ldg(p)
#else
#define KG_POPC64(x) __builtin_popcountll(x)
#define KG_POPC32(x) __builtin_popcount(x)
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

constexpr int KG_SEG_CAP = 64;   // segments a group collects before it scans them
constexpr int KG_ACC_CAP = 64;   // accepted keys a group collects before it compresses them to five
constexpr unsigned long long KG_KEY_LT_1 = (0x3f800000ull << 32) - 1ull;  // accepts exactly the keys with d2 < 1.0f

struct KnnScratch {  // per group (shared memory on the device)
  unsigned long long acc[KG_ACC_CAP];  // accepted keys; after a search: the result, ascending
  int seg_s[KG_SEG_CAP];               // first point of every collected segment
  int seg_pre[KG_SEG_CAP + 8];         // collected: lengths; flattened: [i] = offset of segment i, [nseg] = total
};

enum { KG_START = 0, KG_BALL = 1, KG_SPARSE1 = 2, KG_SPARSE2 = 3, KG_DONE = 4 };

#ifdef LMSF_KNN_STATS  // tuning builds: work counters (global atomics), see kg_stat
__device__ unsigned long long g_knn_stat[16];
HD void kg_stat_add(int slot, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  if (v) atomicAdd(&g_knn_stat[slot], v);
#endif
}
#define KG_STAT(slot, v) kg_stat_add(slot, (unsigned long long)(v))
#else
#define KG_STAT(slot, v)
#endif

// ---- executors -------------------------------------------------------------------------------------------------
// eight lanes per query, four queries per warp; every warp-wide primitive is executed by all 32 lanes
struct WarpGroup8 {
  static constexpr int G = 8;
  int l;            // lane inside the group
  unsigned gshift;  // bit position of the group inside a warp ballot
#if defined(__CUDA_ARCH__)
  HD WarpGroup8() {
    const int lane = threadIdx.x & 31;
    l = lane & 7;
    gshift = lane & 24;
  }
  HD unsigned ballot(bool p) const { return (__ballot_sync(0xffffffffu, p) >> gshift) & 0xffu; }
  HD bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }  // warp-wide: loop control
  HD int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src, 8); }
  HD int shfl_up(int v, int d) const { return __shfl_up_sync(0xffffffffu, v, d, 8); }
  HD unsigned long long shfl_xor64(unsigned long long v, int d) const { return __shfl_xor_sync(0xffffffffu, v, d, 8); }
  HD void sync() const { __syncwarp(); }
#else  // host pass of nvcc: never executed
  HD WarpGroup8() : l(0), gshift(0) {}
  HD unsigned ballot(bool) const { return 0u; }
  HD bool any(bool) const { return false; }
  HD int shfl(int v, int) const { return v; }
  HD int shfl_up(int v, int) const { return v; }
  HD unsigned long long shfl_xor64(unsigned long long v, int) const { return v; }
  HD void sync() const {}
#endif
};
struct HostGroup1 {  // the same source with one lane per query (CPU model, csrc/test_knn_model.cu)
  static constexpr int G = 1;
  int l = 0;
  HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
  HD bool any(bool p) const { return p; }
  HD int shfl(int v, int) const { return v; }
  HD int shfl_up(int v, int) const { return v; }
  HD unsigned long long shfl_xor64(unsigned long long v, int) const { return v; }
  HD void sync() const {}
};

// ---- select: the five smallest keys of acc[0, n), ascending, to acc[0, min(n, 5)); n <- min(n, 5) ---------------------
template <class X>
HD void kg_select5(const X& x, KnnScratch* s, int& n) {
  constexpr int G = X::G;
  const bool small = n <= G;
  // lists of up to G keys: one key per lane, rank = number of smaller keys (keys are unique: indices are)
  unsigned long long mine = ~0ull;
  int rank = 0;
  if (x.any(small && n > 0)) {
    if (small && x.l < n) mine = s->acc[x.l];
    for (int j = 0; x.any(small && j < n); ++j) {
      if (small && j < n) rank += (s->acc[j] < mine) ? 1 : 0;
    }
  }
  // longer lists: five rounds of "smallest key above the last one taken"
  unsigned long long res[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) res[r] = ~0ull;
  if (x.any(!small)) {
    unsigned long long last = 0ull;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      unsigned long long m = ~0ull;
      for (int e0 = 0; x.any(!small && e0 < n); e0 += G) {
        const int e = e0 + x.l;
        if (!small && e < n) {
          const unsigned long long k = s->acc[e];
          if ((r == 0 || k > last) && k < m) m = k;
        }
      }
#pragma unroll
      for (int d = G >> 1; d > 0; d >>= 1) {
        const unsigned long long o = x.shfl_xor64(m, d);
        m = o < m ? o : m;
      }
      res[r] = m;
      last = m;
    }
  }
  x.sync();  // every lane has read what it needs of acc[]
  if (small) {
    if (x.l < n && rank < 5) s->acc[rank] = mine;
  } else if (x.l == 0) {
#pragma unroll
    for (int r = 0; r < 5; ++r)
      if (r < n) s->acc[r] = res[r];
  }
  x.sync();
  if (n > 5) n = 5;
}

// ---- sweep ---------------------------------------------------------------------------------------------------------
// Per-group search state (every lane of a group holds the same values).
struct KgState {
  float qx, qy, qz;
  int lo[3], hi[3];              // box of absolute L2 cells to visit (inclusive)
  bool active;                   // this group takes part in the sweep
  bool rows;                     // tasks are (z,y) rows of L2 cells (else whole L1 cells)
  bool ball;                     // bound_d is valid: skip cells farther than it
  float bound_d;                 // squared radius for pruning (the distance part of bound_key, or 1.0)
  unsigned long long bound_key;  // accept keys <= bound_key
  int nacc;                      // keys in acc[]
};

// scan the collected segments of every group: flatten, stride over the candidates, accept, compress when full
template <class X>
HD void kg_scan_segments(const X& x, const MapView& mv, KnnScratch* s, KgState& st, int& nseg) {
  constexpr int G = X::G;
  // exclusive prefix of the segment lengths, in place
  int run = 0;
  for (int i0 = 0; x.any(i0 < nseg); i0 += G) {
    const int i = i0 + x.l;
    const int len = (i < nseg) ? s->seg_pre[i] : 0;
    int inc = len;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      const int t = x.shfl_up(inc, d);
      if (x.l >= d) inc += t;
    }
    if (i < nseg) s->seg_pre[i] = run + inc - len;
    run += x.shfl(inc, G - 1);
  }
  if (x.l == 0) s->seg_pre[nseg] = run;
  x.sync();
  const int total = run;
  KG_STAT(2, (x.l == 0) ? total : 0);
  int j = 0;
  for (int c0 = 0; x.any(c0 < total); c0 += G) {
    const int c = c0 + x.l;
    bool pass = false;
    unsigned long long key = 0ull;
    if (c < total) {
      while (c >= s->seg_pre[j + 1]) ++j;
      const float4 m = KG_LD(&mv.sorted[s->seg_s[j] + (c - s->seg_pre[j])]);
      const float dx = m.x - st.qx, dy = m.y - st.qy, dz = m.z - st.qz;
      float r = dx * dx;
      r = r + dy * dy;
      r = r + dz * dz;
      key = kg_key(r, (int)f2u(m.w));
      pass = key <= st.bound_key;
    }
    const unsigned gb = x.ballot(pass);
    if (pass) s->acc[st.nacc + KG_POPC32(gb & ((1u << x.l) - 1u))] = key;
    st.nacc += KG_POPC32(gb);
    if (x.any(st.nacc > KG_ACC_CAP - G)) {
      // a list about to overflow is cut to its five smallest keys; their largest is the new bound
      x.sync();
      const bool cut = st.nacc > KG_ACC_CAP - G;
      int n = cut ? st.nacc : 0;
      kg_select5(x, s, n);
      if (cut) {
        st.nacc = n;
        st.bound_key = s->acc[4];
        st.bound_d = kg_key_d2(st.bound_key);
        st.ball = true;
      }
      KG_STAT(5, (x.l == 0 && cut) ? 1 : 0);
    }
  }
  x.sync();
  nseg = 0;
}

// one sweep over the box of every active group
template <class X>
HD void kg_sweep(const X& x, const MapView& mv, const MapDev& md, KnnScratch* s, KgState& st) {
  constexpr int G = X::G;
  // L0 cells the box touches (at most 3 per axis)
  int c0[3], n0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c0[a] = st.lo[a] >> 4;
    n0[a] = (st.hi[a] >> 4) - c0[a] + 1;
  }
  const int n0_all = st.active ? n0[0] * n0[1] * n0[2] : 0;
  const float inv_n0xy = 1.0f / (float)(n0[0] * n0[1] > 0 ? n0[0] * n0[1] : 1);
  const float inv_n0x = 1.0f / (float)(n0[0] > 0 ? n0[0] : 1);
  int nseg = 0;
  int i0 = -1;             // current L0 cell of the box
  int ntasks = 0, t0 = 0;  // tasks of the current L0 cell, next task round
  // state of the current L0 cell
  unsigned long long m1 = 0ull;
  int fine_base = 0;
  int slo[3] = {0, 0, 0}, shi[3] = {0, 0, 0};
  int na = 1, nb = 1;                // rows: L1 parts per row, rows per layer; cells: L1 cells per row, per layer
  float inv_na = 1.0f, inv_nb = 1.0f, inv_nab = 1.0f;
  bool tasks_done = false;
  while (true) {
    if (!tasks_done) {
      if (!x.any(t0 < ntasks)) {
        // every group has dealt out the tasks of its current L0 cell: on to the next one
        ++i0;
        if (!x.any(i0 < n0_all)) {
          tasks_done = true;
        } else {
          ntasks = 0;
          t0 = 0;
          if (i0 < n0_all) {
            const int iz = div_small(i0, inv_n0xy);
            const int rem = i0 - iz * (n0[0] * n0[1]);
            const int iy = div_small(rem, inv_n0x);
            const int cx = c0[0] + (rem - iy * n0[0]), cy = c0[1] + iy, cz = c0[2] + iz;
            bool ok = true;
            if (st.ball) ok = !(box_d2(st.qx, st.qy, st.qz, cx, cy, cz, 1.0f) > st.bound_d);
            const CellRec* rec = ok ? find_cell(mv, md, cx - md.min_c[0], cy - md.min_c[1], cz - md.min_c[2]) : nullptr;
            KG_STAT(0, (x.l == 0) ? 1 : 0);
            if (rec) {
              m1 = KG_LD(&rec->mask);
              fine_base = KG_LD(&rec->fine_base);
              const int cc[3] = {cx, cy, cz};
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                slo[a] = st.lo[a] > (cc[a] << 4) ? st.lo[a] : (cc[a] << 4);
                shi[a] = st.hi[a] < (cc[a] << 4) + 15 ? st.hi[a] : (cc[a] << 4) + 15;
              }
              if (st.rows) {
                na = (shi[0] >> 2) - (slo[0] >> 2) + 1;  // L1 cells the x-run crosses
                nb = shi[1] - slo[1] + 1;                // rows per z layer
                ntasks = na * nb * (shi[2] - slo[2] + 1);
              } else {
                na = (shi[0] >> 2) - (slo[0] >> 2) + 1;
                nb = (shi[1] >> 2) - (slo[1] >> 2) + 1;
                ntasks = na * nb * ((shi[2] >> 2) - (slo[2] >> 2) + 1);
              }
              inv_na = 1.0f / (float)na;
              inv_nb = 1.0f / (float)nb;
              inv_nab = 1.0f / (float)(na * nb);
            }
          }
        }
      } else {
        // one round of tasks: lane l takes task t0 + l of its group
        const int t = t0 + x.l;
        t0 += G;
        int seg_s = 0, seg_n = 0;
        if (t < ntasks) {
          KG_STAT(1, 1);
          if (st.rows) {
            const int r = div_small(t, inv_na), p = t - r * na;
            const int zz = div_small(r, inv_nb);
            const int y = slo[1] + (r - zz * nb), z = slo[2] + zz;
            int xa = slo[0], xb = shi[0];
            bool ok = true;
            if (st.ball) {
              const float gz = axis_gap(st.qz, z, 0.0625f), gy = axis_gap(st.qy, y, 0.0625f);
              const float g2 = gy * gy + gz * gz;  // lower bound: the y and z terms of a point distance
              ok = !(g2 > st.bound_d);
              if (ok) {
                // x extent of the ball in this row (slightly widened; a superset is always correct)
                const float rem = st.bound_d - g2;
                const float rx = sqrtf(rem > 0.0f ? rem : 0.0f) * 1.0001f + 1.0e-6f;
                const int bx0 = (int)floorf((st.qx - rx) * 16.0f), bx1 = (int)floorf((st.qx + rx) * 16.0f);
                xa = bx0 > xa ? bx0 : xa;
                xb = bx1 < xb ? bx1 : xb;
              }
            }
            const int lx = (slo[0] >> 2) + p;  // absolute L1 x of this part of the row
            const int xlo = xa > (lx << 2) ? xa : (lx << 2);
            const int xhi = xb < (lx << 2) + 3 ? xb : (lx << 2) + 3;
            ok = ok && xlo <= xhi;
            const int f1 = (((z >> 2) & 3) << 4) | (((y >> 2) & 3) << 2) | (lx & 3);
            ok = ok && ((m1 >> f1) & 1ull);
            if (ok) {
              const L1Rec lr = ld_l1(mv, fine_base + KG_POPC64(m1 & ((1ull << f1) - 1ull)));
              const int row2 = ((z & 3) << 4) | ((y & 3) << 2);
              const int flo = row2 | (xlo & 3), fhi = row2 | (xhi & 3);
              const unsigned long long below = (1ull << flo) - 1ull;
              const unsigned long long sub = lr.mask & ((2ull << fhi) - 1ull) & ~below;
              if (sub) {
                const int b = lr.first + KG_POPC64(lr.mask & below);
                seg_s = KG_LD(&mv.l2_start[b]);
                seg_n = KG_LD(&mv.l2_start[b + KG_POPC64(sub)]) - seg_s;
              }
            }
          } else {
            const int iz = div_small(t, inv_nab);
            const int rem = t - iz * (na * nb);
            const int iy = div_small(rem, inv_na);
            const int lx = (slo[0] >> 2) + (rem - iy * na), ly = (slo[1] >> 2) + iy, lz = (slo[2] >> 2) + iz;
            const int f1 = ((lz & 3) << 4) | ((ly & 3) << 2) | (lx & 3);
            bool ok = (m1 >> f1) & 1ull;
            if (ok && st.ball) ok = !(box_d2(st.qx, st.qy, st.qz, lx, ly, lz, 0.25f) > st.bound_d);
            if (ok) {
              const L1Rec lr = ld_l1(mv, fine_base + KG_POPC64(m1 & ((1ull << f1) - 1ull)));
              seg_s = KG_LD(&mv.l2_start[lr.first]);
              seg_n = KG_LD(&mv.l2_start[lr.first + KG_POPC64(lr.mask)]) - seg_s;
            }
          }
        }
        const bool has = seg_n > 0;
        const unsigned gb = x.ballot(has);
        if (has) {
          const int at = nseg + KG_POPC32(gb & ((1u << x.l) - 1u));
          s->seg_s[at] = seg_s;
          s->seg_pre[at] = seg_n;
        }
        nseg += KG_POPC32(gb);
      }
    }
    if (tasks_done || x.any(nseg > KG_SEG_CAP - G)) {
      x.sync();
      kg_scan_segments(x, mv, s, st, nseg);
      if (tasks_done) break;
    }
  }
}

// ---- the search ----------------------------------------------------------------------------------------------------
// Exact 5-NN within squared radius 1.0 of one query per group.  seed_key = 0: no prior knowledge; otherwise an
// inclusive upper bound of the 5th key (five map points are known to have keys <= seed_key).  On return acc[0, n)
// holds the n <= 5 nearest keys, ascending; n < 5: fewer than five points within the radius (callers reject).
template <class X>
HD int kg_knn5(const X& x, const MapView& mv, KnnScratch* s, float qx, float qy, float qz, bool active,
               unsigned long long seed_key) {
  const MapDev md = *mv.dev;
  KgState st;
  st.qx = qx;
  st.qy = qy;
  st.qz = qz;
  st.nacc = 0;
  if (!(md.n > 0) || !(fabsf(qx) < 2.0e5f && fabsf(qy) < 2.0e5f && fabsf(qz) < 2.0e5f)) active = false;
  // absolute L2 cell coordinates of the query (x16 is exact in fp32)
  const int a[3] = {active ? (int)floorf(qx * 16.0f) : 0, active ? (int)floorf(qy * 16.0f) : 0,
                    active ? (int)floorf(qz * 16.0f) : 0};
  const float q[3] = {qx, qy, qz};
#pragma unroll
  for (int k = 0; k < 3; ++k) st.lo[k] = st.hi[k] = 0;
  int state = active ? KG_START : KG_DONE;
  st.bound_key = KG_KEY_LT_1;
  st.bound_d = 1.0f;
  if (active && seed_key != 0ull && seed_key <= KG_KEY_LT_1) {  // a seed beyond the search radius is no bound
    state = KG_BALL;
    st.bound_key = seed_key;
    st.bound_d = kg_key_d2(seed_key);
  }
  int n_res = 0;
  while (x.any(state != KG_DONE)) {
    st.active = state != KG_DONE;
    st.nacc = 0;
    if (state == KG_START) {
      // the 27 L2 cells around the query, no prior bound
      st.rows = true;
      st.ball = false;
      st.bound_key = KG_KEY_LT_1;
      st.bound_d = 1.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        st.lo[k] = a[k] - 1;
        st.hi[k] = a[k] + 1;
      }
    } else if (state == KG_BALL) {
      // five points are known within bound_key (inclusive): the cells the ball touches, rows while it is small
      st.ball = true;
      const float rad = sqrtf(st.bound_d) * 1.0001f + 1.0e-6f;
      int w = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        st.lo[k] = (int)floorf((q[k] - rad) * 16.0f);
        st.hi[k] = (int)floorf((q[k] + rad) * 16.0f);
        w = st.hi[k] - st.lo[k] > w ? st.hi[k] - st.lo[k] : w;
      }
      st.rows = w <= 10;
    } else if (state == KG_SPARSE1) {
      // fewer than five points in the 27 L2 cells: the 27 L1 cells, whole cells
      st.rows = false;
      st.ball = false;
      st.bound_key = KG_KEY_LT_1;
      st.bound_d = 1.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        st.lo[k] = ((a[k] >> 2) - 1) << 2;
        st.hi[k] = (((a[k] >> 2) + 1) << 2) + 3;
      }
    } else if (state == KG_SPARSE2) {
      // everything within the search radius
      st.rows = false;
      st.ball = true;
      st.bound_key = KG_KEY_LT_1;
      st.bound_d = 1.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        st.lo[k] = (int)floorf((q[k] - 1.0f) * 16.0f);
        st.hi[k] = (int)floorf((q[k] + 1.0f) * 16.0f);
      }
    }
    KG_STAT(8 + (state < 4 ? state : 0), (x.l == 0 && state != KG_DONE) ? 1 : 0);
    kg_sweep(x, mv, md, s, st);
    int n = st.active ? st.nacc : 0;
    kg_select5(x, s, n);
    if (st.active) {
      const bool full = n == 5;
      const unsigned long long k5 = full ? s->acc[4] : 0ull;
      const float d5 = kg_key_d2(k5);
      n_res = n;
      if (state == KG_START) {
        if (!full) {
          state = KG_SPARSE1;
        } else if (d5 < 0.00390625f) {
          state = KG_DONE;  // 5th distance < one L2 cell: nothing outside the 27 cells can be closer
        } else {
          state = KG_BALL;
          st.bound_key = k5;
          st.bound_d = d5;
        }
      } else if (state == KG_SPARSE1) {
        if (!full) {
          state = KG_SPARSE2;
        } else if (d5 < 0.0625f) {
          state = KG_DONE;  // 5th distance < one L1 cell: the 27 L1 cells hold every closer point
        } else {
          state = KG_BALL;
          st.bound_key = k5;
          st.bound_d = d5;
        }
      } else {
        state = KG_DONE;  // KG_BALL and KG_SPARSE2 are exhaustive
      }
    }
    x.sync();
  }
  return n_res;
}

// five neighbours of one query as the fits read them (match.cu): ascending by (distance, index)
struct Top5 {
  float d[5];
  int id[5];
  HD bool full() const { return id[4] >= 0; }
};

// kernels that search: blocks of KG_BLOCK threads = KG_QPB queries in flight, one KnnScratch per group
constexpr int KG_BLOCK = 128;
constexpr int KG_QPB = KG_BLOCK / WarpGroup8::G;

}  // namespace lm
