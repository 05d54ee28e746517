// mapindex.cu — build of the local-map kNN index (uniform hash grid, cell-sorted points).
//
// Replaces pcl::KdTreeFLANN::setInputCloud behind FeatureMatch::SetSearchTarget
// (registration/FeatureMatch/FeatureMatchBase.hpp:40-44), which the tracker triggers
// on every keyframe through SetInputSource (LidarTracker/LidarTrackerLocalMap.hpp:229).
// Build = bbox (only for maps set through lmsf_map_set) -> three-level counting sort: points grouped by 1 m cell
// through the hash table, every cell's segment sorted by its 6-bit 0.25 m | 6-bit 0.0625 m sub-cell in shared
// memory, which also yields the L1 masks, the L2 cell starts and the hash-table record of the cell.
#include <stdlib.h>

#include "common.cuh"
#include "knn.cuh"

namespace lm {

// defined in voxel.cu
__global__ void k_bbox_init(unsigned* bbox);
__global__ void k_bbox(const float4* __restrict__ in, int n, unsigned* __restrict__ bbox);

// d_cnt layout: [0] occupied L2 cells (+ one sentinel per L0 cell)  [1] occupied L1 cells  [2] occupied L0 cells
//               [3] hash-table overflow flag  [4] point cursor (segment allocation)
enum { CNT_L2 = 0, CNT_L1 = 1, CNT_L0 = 2, CNT_FAIL = 3, CNT_PTS = 4, CNT_CELLS = 5, CNT_WORDS = 8 };  // (CNT_PTS, CNT_CELLS): one 8-byte word

// The index is built by a counting sort in three levels instead of a global radix sort: the order of the L0 cells
// in memory is irrelevant (they are found through the hash table) and so is the order of the points inside one
// L2 cell (every search result is ordered by (distance, original index)), so it is enough to (1) count the points
// of every L0 cell while inserting the cells into the hash table, (2) give every L0 cell a contiguous segment,
// (3) move the points into their segment, (4) counting-sort every segment by its 12-bit (L1, L2) sub-cell in
// shared memory, emitting the L1 masks and L2 cell starts on the way.  Six launches, no library calls, no host
// round trip.

// (1) one thread per map point: find-or-insert its L0 cell, take a rank inside it.  The table arrives filled with
// 0xff: key = ~0 (empty), end = -1 so that the first atomicAdd returns rank 0 (count = end + 1).
__global__ void __launch_bounds__(256) k_cell_count(const float4* __restrict__ in, int n,
                                                    const MapDev* __restrict__ dev, CellRec* __restrict__ table,
                                                    unsigned tmask, int2* __restrict__ slot_rank,
                                                    int* __restrict__ d_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const MapDev md = *dev;
  float4 p = (i < n) ? in[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  int slot = -1, rank = 0;
  if (i < n && isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    // absolute L2 (0.0625 m) cell coordinates; x16 is exact in fp32
    int ax = (int)floorf(p.x * 16.0f), ay = (int)floorf(p.y * 16.0f), az = (int)floorf(p.z * 16.0f);
    unsigned long long ck = pack_cell(md, (ax >> 4) - md.min_c[0], (ay >> 4) - md.min_c[1], (az >> 4) - md.min_c[2]);
    unsigned h = hash_cell(ck) & tmask;
    unsigned probes = 0;
    while (true) {
      unsigned long long k = table[h].key;
      if (k == ~0ull) k = atomicCAS(&table[h].key, ~0ull, ck);
      if (k == ~0ull || k == ck) {
        slot = (int)h;
        break;
      }
      h = (h + 1) & tmask;
      if (++probes > tmask) {  // table full: flagged, reported as LMSF_ERR_CAPACITY with the next read-back
        atomicExch(&d_cnt[CNT_FAIL], 1);
        break;
      }
    }
  }
  {
    // warp-aggregated count: the lanes that hit the same cell take consecutive ranks from one atomic
    // (all 32 lanes take part in the match; the lanes without a cell form a group that does nothing)
    const unsigned peers = __match_any_sync(0xffffffffu, slot);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (slot >= 0 && lane == leader) base = atomicAdd(&table[slot].end, __popc(peers)) + 1;
    base = __shfl_sync(0xffffffffu, base, leader);
    rank = base + __popc(peers & ((1u << lane) - 1u));
  }
  if (i < n) slot_rank[i] = make_int2(slot, rank);
}

// (2) one thread per table slot: occupied cells get their segment [start, end) and an entry in the cell list
__global__ void __launch_bounds__(256) k_cell_alloc(CellRec* __restrict__ table, unsigned slots,
                                                    int* __restrict__ cell_list, int* __restrict__ d_cnt) {
  unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = h < slots && table[h].key != ~0ull;
  int cnt = live ? table[h].end + 1 : 0;
  // warp-aggregated allocation: one atomic per warp for the points, one for the list
  int lane = threadIdx.x & 31;
  int inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  unsigned lv = __ballot_sync(0xffffffffu, live);
  int tot = __shfl_sync(0xffffffffu, inc, 31);
  // ONE 64-bit atomic hands out the point range and the list positions together (points in the low word = CNT_PTS,
  // cells in the high word = CNT_CELLS): list position and segment start then grow in the same order, which is what
  // lets k_cell_sort place a cell's records at offsets derived from (start, list position) without a counter
  int pbase = 0, cbase = 0;
  if (lane == 0 && lv) {
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(&d_cnt[CNT_PTS]),
                                             ((unsigned long long)__popc(lv) << 32) | (unsigned long long)(unsigned)tot);
    pbase = (int)(unsigned)(old & 0xffffffffull);
    cbase = (int)(unsigned)(old >> 32);
  }
  pbase = __shfl_sync(0xffffffffu, pbase, 0);
  cbase = __shfl_sync(0xffffffffu, cbase, 0);
  if (live) {
    int start = pbase + inc - cnt;
    table[h].start = start;
    table[h].end = start + cnt;
    cell_list[cbase + __popc(lv & ((1u << lane) - 1u))] = (int)h;
  }
}

// (3) one thread per map point: into its cell's segment, .w = original index
__global__ void __launch_bounds__(256) k_cell_scatter(const float4* __restrict__ in, int n,
                                                      const int2* __restrict__ slot_rank,
                                                      const CellRec* __restrict__ table,
                                                      float4* __restrict__ grouped) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int2 sr = slot_rank[i];
  if (sr.x < 0) return;
  float4 p = in[i];
  p.w = __int_as_float(i);
  grouped[table[sr.x].start + sr.y] = p;
}

__device__ __forceinline__ unsigned sub_cell(float4 p) {
  int ax = (int)floorf(p.x * 16.0f), ay = (int)floorf(p.y * 16.0f), az = (int)floorf(p.z * 16.0f);
  unsigned f1 = (((az >> 2) & 3) << 4) | (((ay >> 2) & 3) << 2) | ((ax >> 2) & 3);
  unsigned f2 = ((az & 3) << 4) | ((ay & 3) << 2) | (ax & 3);
  return (f1 << 6) | f2;
}

// (4) one block per occupied L0 cell: counting sort of its segment by the 12-bit (L1, L2) sub-cell.  Thread t owns
// the CS_BINS bins [CS_BINS t, CS_BINS (t + 1)) = 1 / CS_LPL of L1 cell t / CS_LPL; the L1 masks, the L2 cell starts (plus one sentinel
// per L0 cell, so that "start of the next L2 cell" is always the end of the previous one) and the cell record are
// written from the same histogram.
#ifndef CS_THREADS_N
#define CS_THREADS_N 512  // measured: 256 -> 726, 512 -> 713, 1024 -> 743 us per sweep (128: 754)
#endif
constexpr int CS_THREADS = CS_THREADS_N;
constexpr int CS_BINS = 4096 / CS_THREADS;   // bins per thread
constexpr int CS_LPL = 64 / CS_BINS;         // lanes per L1 cell
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_L1W = 32 / CS_LPL;          // L1 cells per warp
constexpr int CS_HIST_WORDS = 4096 + 4096 / 32;
__device__ __forceinline__ int cs_bin(int b) { return b + (b >> 5); }
#ifndef CS_KEEP_N
#define CS_KEEP_N 2  // 2: no spills at two resident blocks of 512 threads (3: 12 bytes, 4: 96 bytes)
#endif
constexpr int CS_KEEP = CS_KEEP_N;           // points per thread loaded together / kept in registers between the passes
__global__ void __launch_bounds__(CS_THREADS, 2) k_cell_sort(const int* __restrict__ cell_list,
                                                          CellRec* __restrict__ table,
                                                          const float4* __restrict__ grouped,
                                                          float4* __restrict__ sorted,
                                                          L1Rec* __restrict__ l1, int* __restrict__ l2_start,
                                                          int* __restrict__ d_cnt) {
  // bin b lives at word b + b / 32: a thread's CS_BINS consecutive bins (stride CS_BINS words between lanes: an
  // 8-way bank conflict on every one of the 2 x CS_BINS scan accesses of a cell) then fall into different banks for all
  // 32 lanes; ncu had the kernel stalled on the shared-memory queue (mio_throttle) first, barriers second
  __shared__ int hist[CS_HIST_WORDS];
  __shared__ int warp_pts[CS_WARPS], warp_n2[CS_WARPS], warp_n1[CS_WARPS];
  __shared__ int s_tot1, s_tot2;
  __shared__ unsigned long long s_m1[CS_WARPS];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n_cells = d_cnt[CNT_CELLS];  // final: written by k_cell_alloc, the previous launch on this stream
  for (int ci = blockIdx.x; ci < n_cells; ci += gridDim.x) {
  const int h = cell_list[ci];
  const int s = table[h].start, e = table[h].end;
  __syncthreads();  // the previous cell's scatter is done with the bins
  for (int i = t; i < CS_HIST_WORDS; i += CS_THREADS) hist[i] = 0;
  // The first CS_KEEP points of every thread are loaded together (one round trip instead of one per point: in a dense
  // cell the chain of dependent loads was the kernel) and stay in registers for the scatter pass; cells with more
  // than CS_KEEP * CS_THREADS points re-read the rest, again CS_KEEP loads at a time.
  float4 keep[CS_KEEP];
#pragma unroll
  for (int k = 0; k < CS_KEEP; ++k) {
    const int i = s + t + k * CS_THREADS;
    if (i < e) keep[k] = grouped[i];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < CS_KEEP; ++k)
    if (s + t + k * CS_THREADS < e) atomicAdd(&hist[cs_bin((int)sub_cell(keep[k]))], 1);
  for (int i0 = s + t + CS_KEEP * CS_THREADS; i0 < e; i0 += CS_KEEP * CS_THREADS) {
    float4 q[CS_KEEP];
#pragma unroll
    for (int k = 0; k < CS_KEEP; ++k)
      if (i0 + k * CS_THREADS < e) q[k] = grouped[i0 + k * CS_THREADS];
#pragma unroll
    for (int k = 0; k < CS_KEEP; ++k)
      if (i0 + k * CS_THREADS < e) atomicAdd(&hist[cs_bin((int)sub_cell(q[k]))], 1);
  }
  __syncthreads();
  // thread-local: points, occupied L2 cells and their mask
  int cnt[CS_BINS];
  int pts = 0, n2 = 0;
  unsigned mloc = 0;
#pragma unroll
  for (int k = 0; k < CS_BINS; ++k) {
    cnt[k] = hist[cs_bin(t * CS_BINS + k)];
    pts += cnt[k];
    if (cnt[k]) {
      ++n2;
      mloc |= 1u << k;
    }
  }
  // the CS_LPL threads of one L1 cell are consecutive lanes
  unsigned long long m2 = (unsigned long long)mloc << (CS_BINS * (lane & (CS_LPL - 1)));
#pragma unroll
  for (int d = 1; d < CS_LPL; d <<= 1) m2 |= __shfl_xor_sync(0xffffffffu, m2, d);
  const bool l1_head = (lane & (CS_LPL - 1)) == 0 && m2 != 0ull;
  int n1 = l1_head ? 1 : 0;
  // block-wide exclusive scans of (pts, n2, n1)
  int ip = pts, i2 = n2, i1 = n1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int a = __shfl_up_sync(0xffffffffu, ip, d), b = __shfl_up_sync(0xffffffffu, i2, d),
        cc = __shfl_up_sync(0xffffffffu, i1, d);
    if (lane >= d) {
      ip += a;
      i2 += b;
      i1 += cc;
    }
  }
  if (lane == 31) {
    warp_pts[warp] = ip;
    warp_n2[warp] = i2;
    warp_n1[warp] = i1;
  }
  __syncthreads();
  // offsets of the warps: one warp scans the CS_WARPS totals and reserves the cell's records (every thread walking all
  // the totals was a quarter of the kernel's instructions: 23.2 M -> 17.0 M warp instructions per surf build; the launch
  // time did not follow — 67-70 us either way, r4j: the kernel is bound by the barrier / round-trip chain of a cell)
  if (warp == 0) {
    int vp = lane < CS_WARPS ? warp_pts[lane] : 0, v2 = lane < CS_WARPS ? warp_n2[lane] : 0,
        v1 = lane < CS_WARPS ? warp_n1[lane] : 0;
    int sp = vp, s2 = v2, s1 = v1;
#pragma unroll
    for (int d = 1; d < CS_WARPS; d <<= 1) {
      int a = __shfl_up_sync(0xffffffffu, sp, d), b = __shfl_up_sync(0xffffffffu, s2, d),
          cc = __shfl_up_sync(0xffffffffu, s1, d);
      if (lane >= d) {
        sp += a;
        s2 += b;
        s1 += cc;
      }
    }
    if (lane < CS_WARPS) {  // exclusive
      warp_pts[lane] = sp - vp;
      warp_n2[lane] = s2 - v2;
      warp_n1[lane] = s1 - v1;
    }
    if (lane == CS_WARPS - 1) {
      s_tot1 = s1;
      s_tot2 = s2;
      atomicAdd(&d_cnt[CNT_L1], s1);  // statistics only (MapDev::n_fine): no value comes back, nobody waits
    }
  }
  __syncthreads();
  const int tot2 = s_tot2, tot1 = s_tot1;
  const int ex_p = warp_pts[warp] + ip - pts, ex_2 = warp_n2[warp] + i2 - n2, ex_1 = warp_n1[warp] + i1 - n1;
  // A cell with k points has at most k occupied L1 cells and at most k occupied L2 cells, and the cells' segments are
  // disjoint and in list order (k_cell_alloc): its L1 records go to l1[s, ...) and its L2 starts + sentinel to
  // l2_start[s + ci, e + ci] — no counter, no global round trip in the middle of a cell
  const int base1 = s, base2 = s + ci;
  // L2 cell starts of this thread's occupied bins; the bins become running cursors for the scatter
  {
    int run = s + ex_p, r2 = base2 + ex_2;
#pragma unroll
    for (int k = 0; k < CS_BINS; ++k) {
      hist[cs_bin(t * CS_BINS + k)] = run;
      if (cnt[k]) l2_start[r2++] = run;
      run += cnt[k];
    }
  }
  if (l1_head) {
    L1Rec r;
    r.mask = m2;
    r.first = base2 + ex_2;
    r.pad = 0;
    l1[base1 + ex_1] = r;
  }
  // occupancy of the 64 L1 cells: L1 cell j = lanes j CS_LPL .. of warp j / CS_L1W
  unsigned heads = __ballot_sync(0xffffffffu, l1_head);
  unsigned long long m1w = 0ull;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < CS_L1W; ++j)
      if ((heads >> (CS_LPL * j)) & 1u) m1w |= 1ull << (warp * CS_L1W + j);
  }
  if (lane == 0) s_m1[warp] = m1w;
  __syncthreads();
  if (t == 0) {
    unsigned long long m1 = 0ull;
#pragma unroll
    for (int w = 0; w < CS_WARPS; ++w) m1 |= s_m1[w];
    l2_start[base2 + tot2] = e;  // sentinel of this L0 cell
    table[h].mask = m1;
    table[h].fine_base = base1;
    table[h].pad = 0;
  }
#pragma unroll
  for (int k = 0; k < CS_KEEP; ++k)
    if (s + t + k * CS_THREADS < e) sorted[atomicAdd(&hist[cs_bin((int)sub_cell(keep[k]))], 1)] = keep[k];
  for (int i0 = s + t + CS_KEEP * CS_THREADS; i0 < e; i0 += CS_KEEP * CS_THREADS) {
    float4 q[CS_KEEP];
#pragma unroll
    for (int k = 0; k < CS_KEEP; ++k)
      if (i0 + k * CS_THREADS < e) q[k] = grouped[i0 + k * CS_THREADS];
#pragma unroll
    for (int k = 0; k < CS_KEEP; ++k)
      if (i0 + k * CS_THREADS < e) sorted[atomicAdd(&hist[cs_bin((int)sub_cell(q[k]))], 1)] = q[k];
  }
  }  // cells of this block
}

__global__ void k_map_finish(MapDev* __restrict__ dev, int* __restrict__ d_cnt) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    dev->n_fine = d_cnt[CNT_L1];
    d_cnt[CNT_L0] = d_cnt[CNT_CELLS];  // the host reads the table occupancy from words [0, 4)
  }
}

int map_alloc(Ctx* c, MapIndex& m, int cap) {
  m.cap = cap;
  size_t n = (size_t)cap;
  LM_CUDA(cudaMalloc(&m.win, n * sizeof(float4)));
  LM_CUDA(cudaMalloc(&m.win_alt, n * sizeof(float4)));
  LM_CUDA(cudaMalloc(&m.vox, n * sizeof(float4)));
  m.cat = m.win;
  LM_CUDA(cudaMalloc(&m.sorted, n * sizeof(float4)));
  LM_CUDA(cudaMalloc(&m.grouped, n * sizeof(float4)));
  LM_CUDA(cudaMalloc(&m.slot_rank, n * sizeof(int2)));
  LM_CUDA(cudaMalloc(&m.cell_list, n * 4));
  LM_CUDA(cudaMalloc(&m.l2_start, (2 * n + 1) * 4));  // one entry per occupied L2 cell + one sentinel per L0 cell
  LM_CUDA(cudaMalloc(&m.l1, n * sizeof(L1Rec)));
  LM_CUDA(cudaMalloc(&m.d_cnt, CNT_WORDS * sizeof(int)));
  LM_CUDA(cudaMalloc(&m.d_box, 8 * sizeof(unsigned)));
  LM_CUDA(cudaMallocHost(&m.h_cnt, 64 * sizeof(int)));
  memset(m.h_cnt, 0, 64 * sizeof(int));
  m.fixed = false;
  m.n_cells_seen = 0;
  LM_CUDA(cudaMalloc(&m.dev, sizeof(MapDev)));
  LM_CUDA(cudaMemset(m.dev, 0, sizeof(MapDev)));
  m.table_cap = 1u << 18;
  LM_CUDA(cudaMalloc(&m.table, (size_t)m.table_cap * sizeof(CellRec)));
  m.ready = false;
  m.n_host = 0;
  m.frame_n.clear();
  return LMSF_OK;
}

void map_free(MapIndex& m) {
  cudaFree(m.win);
  cudaFree(m.win_alt);
  cudaFree(m.vox);
  cudaFree(m.sorted);
  cudaFree(m.grouped);
  cudaFree(m.slot_rank);
  cudaFree(m.cell_list);
  cudaFree(m.l2_start);
  cudaFree(m.l1);
  cudaFree(m.d_cnt);
  cudaFree(m.d_box);
  cudaFreeHost(m.h_cnt);
  cudaFree(m.dev);
  cudaFree(m.table);
  m = MapIndex();
}

static inline float ord2f_host(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &v, 4);
  return f;
}

// grid frozen for tracker maps: 10 + 10 + 8 bits of 1 m cells around the given sensor position
static constexpr int FIX_BITS[3] = {10, 10, 8};

bool map_freeze_grid(MapIndex& m, const double p[3]) {
  for (int a = 0; a < 3; ++a)
    if (!(fabs(p[a]) < 5.0e4)) return false;
  MapDev md;
  memset(&md, 0, sizeof md);
  for (int a = 0; a < 3; ++a) {
    md.bits[a] = FIX_BITS[a];
    md.dim[a] = 1 << FIX_BITS[a];
    md.min_c[a] = (int)floor(p[a]) - (md.dim[a] >> 1);
  }
  m.host = md;
  m.fixed = true;
  return true;
}

bool map_grid_covers(const MapIndex& m, const double p[3], double reach) {
  if (!m.fixed) return false;
  for (int a = 0; a < 3; ++a) {
    double lo = floor(p[a] - reach) - 1.0, hi = floor(p[a] + reach) + 1.0;
    if (!(lo >= (double)m.host.min_c[a] && hi < (double)m.host.min_c[a] + (double)m.host.dim[a])) return false;
  }
  return true;
}

int wait_map(Ctx* c) {
  if (c->map_pending) {
    LM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_map_done, 0));
    c->map_pending = false;
  }
  return LMSF_OK;
}

// main stream is about to overwrite the current feature slot: the map stream must have consumed it
int wait_feat(Ctx* c) {
  Ctx::FeatSlot& sl = c->slot[c->slot_cur];
  if (sl.freed_pending) {
    LM_CUDA(cudaStreamWaitEvent(c->stream, sl.freed, 0));
    sl.freed_pending = false;
  }
  return LMSF_OK;
}

int map_build(Ctx* c, MapIndex& m, int n, cudaStream_t st, bool fixed_grid) {
  if (n > m.cap) return LMSF_ERR_CAPACITY;
  StageScope scope(c, LMSF_STAGE_MAP, st);
  m.n_host = n;
  m.ready = false;
  if (n == 0) return LMSF_OK;
  const int nb = div_up(n, 256);
  MapDev md;
  int total_bits = 12;
  unsigned slots = 64;
  if (fixed_grid) {
    // frozen origin and key width: nothing to read back; the caller has checked that the points fit
    md = m.host;
    md.n = n;
    total_bits += md.bits[0] + md.bits[1] + md.bits[2];
    long long want = 4LL * ((long long)m.n_cells_seen + 65536);
    while ((long long)slots < want) slots <<= 1;
  } else {
    LM_LAUNCH_ON(c, st, k_bbox_init, 1, 32, 0, m.d_box);
    LM_LAUNCH_ON(c, st, k_bbox, nb < 592 ? nb : 592, 256, 0, m.cat, n, m.d_box);
    unsigned* hb = (unsigned*)m.h_cnt + 8;
    LM_CUDA(cudaMemcpyAsync(hb, m.d_box, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    LM_CUDA(cudaStreamSynchronize(st));
    memset(&md, 0, sizeof md);
    md.n = (int)hb[6];
    if (md.n == 0) return LMSF_OK;
    long long cells = 1;
    bool inside = m.fixed;  // a frozen grid is kept when the cloud still fits it
    int mn_c[3], dim[3];
    for (int a = 0; a < 3; ++a) {
      float mn = ord2f_host(hb[a]), mx = ord2f_host(hb[3 + a]);
      if (!(fabsf(mn) < 1.0e5f && fabsf(mx) < 1.0e5f)) return LMSF_ERR_INVALID;  // beyond the exact-cell-bound range
      mn_c[a] = (int)floorf(mn);
      dim[a] = (int)floorf(mx) - mn_c[a] + 1;
      if (m.fixed && !(mn_c[a] >= m.host.min_c[a] && mn_c[a] + dim[a] <= m.host.min_c[a] + m.host.dim[a])) inside = false;
    }
    if (inside) {
      int keep_n = md.n;
      md = m.host;
      md.n = keep_n;
    } else {
      m.fixed = false;
      for (int a = 0; a < 3; ++a) {
        md.min_c[a] = mn_c[a];
        md.dim[a] = dim[a];
        int b = 0;
        while ((1LL << b) < (long long)dim[a]) ++b;
        md.bits[a] = b;
      }
    }
    for (int a = 0; a < 3; ++a) {
      total_bits += md.bits[a];
      cells *= md.dim[a];
    }
    // one slot per possible coarse cell, at load factor <= 1/2
    long long want = (cells < (long long)md.n ? cells : (long long)md.n) * 2;
    while ((long long)slots < want) slots <<= 1;
  }
  if (slots > m.table_cap) {
    LM_CUDA(cudaFree(m.table));  // synchronises the device: rare (the table only grows)
    m.table = nullptr;
    LM_CUDA(cudaMalloc(&m.table, (size_t)slots * sizeof(CellRec)));
    m.table_cap = slots;
  }
  md.table_mask = slots - 1;
  md.n_fine = 0;
  md.grow0 = m.grow0;
  m.host = md;
  // the descriptor is passed through pinned memory owned by the map: the copy is asynchronous on `st`
  MapDev* hd = (MapDev*)(m.h_cnt + 32);
  *hd = md;
  LM_CUDA(cudaMemcpyAsync(m.dev, hd, sizeof(MapDev), cudaMemcpyHostToDevice, st));
  LM_CUDA(cudaMemsetAsync(m.table, 0xff, (size_t)slots * sizeof(CellRec), st));
  LM_CUDA(cudaMemsetAsync(m.d_cnt, 0, CNT_WORDS * sizeof(int), st));
  LM_LAUNCH_ON(c, st, k_cell_count, nb, 256, 0, m.cat, n, m.dev, m.table, md.table_mask, m.slot_rank, m.d_cnt);
  LM_LAUNCH_ON(c, st, k_cell_alloc, div_up((int)slots, 256), 256, 0, m.table, slots, m.cell_list, m.d_cnt);
  LM_LAUNCH_ON(c, st, k_cell_scatter, nb, 256, 0, m.cat, n, m.slot_rank, m.table, m.grouped);
  // persistent blocks stride over the occupied L0 cells (their number is only known on the device)
  static const int cs_per_sm = [] {  // tuning: LMSF_CS_GRID = blocks per SM
    const char* v = getenv("LMSF_CS_GRID");
    const int x = v ? atoi(v) : 0;
    return x > 0 ? x : (CS_THREADS >= 512 ? 4 : 6);
  }();
  LM_LAUNCH_ON(c, st, k_cell_sort, 148 * cs_per_sm, CS_THREADS, 0, m.cell_list, m.table, m.grouped, m.sorted, m.l1,
               m.l2_start, m.d_cnt);
  LM_LAUNCH_ON(c, st, k_map_finish, 1, 32, 0, m.dev, m.d_cnt);
  // occupancy of the table (and the insert-failed flag) travel back with the next pose read-back
  LM_CUDA(cudaMemcpyAsync(m.h_cnt, m.d_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaGetLastError());
  m.ready = true;
  return LMSF_OK;
}

}  // namespace lm
