// mapindex.cu — build of the local-map kNN index (uniform hash grid, cell-sorted points).
//
// Replaces pcl::KdTreeFLANN::setInputCloud behind FeatureMatch::SetSearchTarget
// (registration/FeatureMatch/FeatureMatchBase.hpp:40-44), which the tracker triggers
// on every keyframe through SetInputSource (LidarTracker/LidarTrackerLocalMap.hpp:229).
// Build = bbox -> 64-bit cell keys (1 m cell | 6-bit 0.25 m sub-cell | 6-bit 0.0625 m sub-cell)
// -> radix sort -> gather -> L2 cell starts -> L1 cell masks -> one hash-table record per 1 m cell.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"
#include "knn.cuh"

namespace lm {

// defined in voxel.cu
__global__ void k_bbox_init(unsigned* bbox);
__global__ void k_bbox(const float4* __restrict__ in, int n, unsigned* __restrict__ bbox);
__global__ void k_head_flags(const unsigned long long* __restrict__ keys, int n, uint8_t* __restrict__ flags);

__global__ void __launch_bounds__(256) k_map_keys(const float4* __restrict__ in, int n,
                                                  const MapDev* __restrict__ dev,
                                                  unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const MapDev md = *dev;
  float4 p = in[i];
  unsigned long long key = ~0ull;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    // absolute L2 (0.0625 m) cell coordinates; x16 is exact in fp32
    int ax = (int)floorf(p.x * 16.0f), ay = (int)floorf(p.y * 16.0f), az = (int)floorf(p.z * 16.0f);
    unsigned long long ck = pack_cell(md, (ax >> 4) - md.min_c[0], (ay >> 4) - md.min_c[1], (az >> 4) - md.min_c[2]);
    unsigned f1 = (((az >> 2) & 3) << 4) | (((ay >> 2) & 3) << 2) | ((ax >> 2) & 3);
    unsigned f2 = ((az & 3) << 4) | ((ay & 3) << 2) | (ax & 3);
    key = (ck << 12) | ((unsigned long long)f1 << 6) | (unsigned long long)f2;
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void __launch_bounds__(256) k_map_gather(const float4* __restrict__ in, const int* __restrict__ vals, int n,
                                                    float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int src = vals[i];
  float4 p = in[src];
  p.w = __int_as_float(src);
  sorted[i] = p;
}

// flags over the occupied L2 cells: 1 where a new L1 cell starts (0 beyond the n2 live entries)
__global__ void __launch_bounds__(256) k_l1_flags(const unsigned long long* __restrict__ keys,
                                                  const int* __restrict__ l2_start, const int* __restrict__ d_n2,
                                                  int n, uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int n2 = *d_n2;
  uint8_t f = 0;
  if (i < n2) f = (i == 0 || (keys[l2_start[i]] >> 6) != (keys[l2_start[i - 1]] >> 6)) ? 1 : 0;
  flags[i] = f;
}

// one thread per occupied L1 cell: its L2 occupancy mask; the first L1 cell of an L0 cell also
// assembles that cell's record and inserts it into the hash table
__global__ void __launch_bounds__(256) k_cell_insert(const unsigned long long* __restrict__ keys,
                                                     int* __restrict__ l2_start, int* __restrict__ l1_first,
                                                     unsigned long long* __restrict__ l1_mask,
                                                     int* __restrict__ d_cnt, int n_pts,
                                                     MapDev* __restrict__ dev, CellRec* __restrict__ table,
                                                     unsigned tmask) {
  const int n2 = d_cnt[0], n1 = d_cnt[1];
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {
    dev->n_fine = n1;
  }
  if (j >= n1) return;
  const int a = l1_first[j];
  const int b = (j + 1 < n1) ? l1_first[j + 1] : n2;
  unsigned long long m2 = 0ull;
  for (int i = a; i < b; ++i) m2 |= 1ull << (unsigned)(keys[l2_start[i]] & 63ull);
  l1_mask[j] = m2;
  const unsigned long long kj = keys[l2_start[a]];
  const unsigned long long ck = kj >> 12;
  if (j > 0 && (keys[l2_start[l1_first[j - 1]]] >> 12) == ck) return;
  unsigned long long m1 = 0ull;
  int last = j;
  for (int g = j; g < n1; ++g) {
    unsigned long long k = keys[l2_start[l1_first[g]]];
    if ((k >> 12) != ck) break;
    m1 |= 1ull << (unsigned)((k >> 6) & 63ull);
    last = g;
  }
  int end = (last + 1 < n1) ? l2_start[l1_first[last + 1]] : n_pts;
  unsigned h = hash_cell(ck) & tmask;
  unsigned probes = 0;
  while (true) {
    unsigned long long prev = atomicCAS(&table[h].key, ~0ull, ck);
    if (prev == ~0ull) break;
    h = (h + 1) & tmask;
    if (++probes > tmask) {  // table full: flagged, reported as LMSF_ERR_CAPACITY with the next read-back
      atomicExch(&d_cnt[3], 1);
      return;
    }
  }
  atomicAdd(&d_cnt[2], 1);
  table[h].mask = m1;
  table[h].start = l2_start[a];
  table[h].end = end;
  table[h].fine_base = j;
  table[h].pad = 0;
}

// sentinels behind the live entries (written after every reader of the previous values is done)
__global__ void k_sentinels(int* __restrict__ l2_start, int* __restrict__ l1_first, const int* __restrict__ d_cnt,
                            int n_pts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    l2_start[d_cnt[0]] = n_pts;
    l1_first[d_cnt[1]] = d_cnt[0];
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
  LM_CUDA(cudaMalloc(&m.keys, n * 8));
  LM_CUDA(cudaMalloc(&m.keys_alt, n * 8));
  LM_CUDA(cudaMalloc(&m.vals, n * 4));
  LM_CUDA(cudaMalloc(&m.vals_alt, n * 4));
  LM_CUDA(cudaMalloc(&m.l2_start, (n + 1) * 4));
  LM_CUDA(cudaMalloc(&m.l1_first, (n + 1) * 4));
  LM_CUDA(cudaMalloc(&m.l1_mask, n * 8));
  LM_CUDA(cudaMalloc(&m.d_cnt, 4 * sizeof(int)));
  LM_CUDA(cudaMalloc(&m.d_box, 8 * sizeof(unsigned)));
  LM_CUDA(cudaMallocHost(&m.h_cnt, 64 * sizeof(int)));
  memset(m.h_cnt, 0, 64 * sizeof(int));
  m.fixed = false;
  m.n_cells_seen = 0;
  LM_CUDA(cudaMalloc(&m.flags, n));
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
  cudaFree(m.keys);
  cudaFree(m.keys_alt);
  cudaFree(m.vals);
  cudaFree(m.vals_alt);
  cudaFree(m.l2_start);
  cudaFree(m.l1_first);
  cudaFree(m.l1_mask);
  cudaFree(m.d_cnt);
  cudaFree(m.d_box);
  cudaFreeHost(m.h_cnt);
  cudaFree(m.flags);
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

int wait_feat(Ctx* c) {
  if (c->feat_pending) {
    LM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_feat_free, 0));
    c->feat_pending = false;
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
  void* tmp_buf = (st == c->stream) ? c->cub_tmp : c->cub_tmp_map;
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
  m.host = md;
  // the descriptor is passed through pinned memory owned by the map: the copy is asynchronous on `st`
  MapDev* hd = (MapDev*)(m.h_cnt + 32);
  *hd = md;
  LM_CUDA(cudaMemcpyAsync(m.dev, hd, sizeof(MapDev), cudaMemcpyHostToDevice, st));
  LM_CUDA(cudaMemsetAsync(m.table, 0xff, (size_t)slots * sizeof(CellRec), st));
  LM_CUDA(cudaMemsetAsync(m.d_cnt, 0, 4 * sizeof(int), st));
  LM_LAUNCH_ON(c, st, k_map_keys, nb, 256, 0, m.cat, n, m.dev, m.keys, m.vals);
  int end_bit = (md.n < n) ? 64 : total_bits;
  size_t tmp = c->cub_tmp_bytes;
  LM_CUDA(cub::DeviceRadixSort::SortPairs(tmp_buf, tmp, m.keys, m.keys_alt, m.vals, m.vals_alt, n, 0, end_bit, st));
  c->launches++;
  const int nf = md.n;
  LM_LAUNCH_ON(c, st, k_map_gather, div_up(nf, 256), 256, 0, m.cat, m.vals_alt, nf, m.sorted);
  LM_LAUNCH_ON(c, st, k_head_flags, div_up(nf, 256), 256, 0, m.keys_alt, nf, m.flags);
  tmp = c->cub_tmp_bytes;
  LM_CUDA(cub::DeviceSelect::Flagged(tmp_buf, tmp, thrust::counting_iterator<int>(0), m.flags, m.l2_start, m.d_cnt + 0,
                                     nf, st));
  c->launches++;
  LM_LAUNCH_ON(c, st, k_l1_flags, div_up(nf, 256), 256, 0, m.keys_alt, m.l2_start, m.d_cnt + 0, nf, m.flags);
  tmp = c->cub_tmp_bytes;
  LM_CUDA(cub::DeviceSelect::Flagged(tmp_buf, tmp, thrust::counting_iterator<int>(0), m.flags, m.l1_first, m.d_cnt + 1,
                                     nf, st));
  c->launches++;
  LM_LAUNCH_ON(c, st, k_cell_insert, div_up(nf, 256), 256, 0, m.keys_alt, m.l2_start, m.l1_first, m.l1_mask, m.d_cnt,
               nf, m.dev, m.table, md.table_mask);
  LM_LAUNCH_ON(c, st, k_sentinels, 1, 32, 0, m.l2_start, m.l1_first, m.d_cnt, nf);
  // occupancy of the table (and the insert-failed flag) travel back with the next pose read-back
  LM_CUDA(cudaMemcpyAsync(m.h_cnt, m.d_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaGetLastError());
  m.ready = true;
  return LMSF_OK;
}

}  // namespace lm
