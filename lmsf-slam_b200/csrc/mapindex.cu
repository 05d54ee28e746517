// mapindex.cu — build of the local-map kNN index (uniform hash grid, cell-sorted points).
//
// Replaces pcl::KdTreeFLANN::setInputCloud behind FeatureMatch::SetSearchTarget
// (registration/FeatureMatch/FeatureMatchBase.hpp:40-44), which the tracker triggers
// on every keyframe through SetInputSource (LidarTracker/LidarTrackerLocalMap.hpp:229).
// Build = bbox -> 64-bit cell keys (coarse 1 m cell | 6-bit 0.25 m sub-cell) -> radix
// sort -> gather -> fine-cell starts -> one hash-table record per coarse cell.
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
    int gx = (int)floorf(p.x * 4.0f) - 4 * md.min_c[0];
    int gy = (int)floorf(p.y * 4.0f) - 4 * md.min_c[1];
    int gz = (int)floorf(p.z * 4.0f) - 4 * md.min_c[2];
    unsigned long long ck = pack_cell(md, gx >> 2, gy >> 2, gz >> 2);
    key = (ck << 6) | (unsigned long long)(((gz & 3) << 4) | ((gy & 3) << 2) | (gx & 3));
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

// one thread per occupied fine cell; the first fine cell of a coarse cell inserts the record
__global__ void __launch_bounds__(256) k_cell_insert(const unsigned long long* __restrict__ keys,
                                                     int* __restrict__ fine_start, const int* __restrict__ d_nfine,
                                                     int n_pts, MapDev* __restrict__ dev, CellRec* __restrict__ table,
                                                     unsigned tmask) {
  const int nf = *d_nfine;
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0) {
    fine_start[nf] = n_pts;
    dev->n_fine = nf;
  }
  if (f >= nf) return;
  int pos = fine_start[f];
  unsigned long long ck = keys[pos] >> 6;
  if (f > 0 && (keys[fine_start[f - 1]] >> 6) == ck) return;
  unsigned long long mask = 0ull;
  int last = f;
  for (int g = f; g < nf; ++g) {
    unsigned long long k = keys[fine_start[g]];
    if ((k >> 6) != ck) break;
    mask |= 1ull << (unsigned)(k & 63ull);
    last = g;
  }
  int end = (last + 1 < nf) ? fine_start[last + 1] : n_pts;
  unsigned h = hash_cell(ck) & tmask;
  while (true) {
    unsigned long long prev = atomicCAS(&table[h].key, ~0ull, ck);
    if (prev == ~0ull) break;
    h = (h + 1) & tmask;
  }
  table[h].mask = mask;
  table[h].start = pos;
  table[h].end = end;
  table[h].fine_base = f;
  table[h].pad = 0;
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
  LM_CUDA(cudaMalloc(&m.fine_start, (n + 1) * 4));
  LM_CUDA(cudaMalloc(&m.flags, n));
  LM_CUDA(cudaMalloc(&m.dev, sizeof(MapDev)));
  LM_CUDA(cudaMemset(m.dev, 0, sizeof(MapDev)));
  m.table_cap = 1u << 16;
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
  cudaFree(m.fine_start);
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

int map_build(Ctx* c, MapIndex& m, int n) {
  if (n > m.cap) return LMSF_ERR_CAPACITY;
  StageScope scope(c, LMSF_STAGE_MAP);
  m.n_host = n;
  m.ready = false;
  if (n == 0) return LMSF_OK;
  const int nb = div_up(n, 256);
  LM_LAUNCH(c, k_bbox_init, 1, 32, 0, c->d_bbox);
  LM_LAUNCH(c, k_bbox, nb < 592 ? nb : 592, 256, 0, m.cat, n, c->d_bbox);
  unsigned* hb = (unsigned*)c->h_ints;
  LM_CUDA(cudaMemcpyAsync(hb, c->d_bbox, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  MapDev md;
  memset(&md, 0, sizeof md);
  md.n = (int)hb[6];
  if (md.n == 0) return LMSF_OK;
  long long cells = 1;
  int total_bits = 6;
  for (int a = 0; a < 3; ++a) {
    float mn = ord2f_host(hb[a]), mx = ord2f_host(hb[3 + a]);
    if (!(fabsf(mn) < 1.0e8f && fabsf(mx) < 1.0e8f)) return LMSF_ERR_INVALID;  // coordinates beyond the grid's int range
    md.min_c[a] = (int)floorf(mn);
    md.dim[a] = (int)floorf(mx) - md.min_c[a] + 1;
    int b = 0;
    while ((1LL << b) < (long long)md.dim[a]) ++b;
    md.bits[a] = b;
    total_bits += b;
    cells *= md.dim[a];
  }
  // one slot per possible coarse cell, at load factor <= 1/2
  long long want = (cells < (long long)md.n ? cells : (long long)md.n) * 2;
  unsigned slots = 64;
  while ((long long)slots < want) slots <<= 1;
  if (slots > m.table_cap) {
    LM_CUDA(cudaFree(m.table));
    m.table = nullptr;
    LM_CUDA(cudaMalloc(&m.table, (size_t)slots * sizeof(CellRec)));
    m.table_cap = slots;
  }
  md.table_mask = slots - 1;
  md.n_fine = 0;
  m.host = md;
  LM_CUDA(cudaMemcpyAsync(m.dev, &m.host, sizeof(MapDev), cudaMemcpyHostToDevice, c->stream));
  LM_CUDA(cudaMemsetAsync(m.table, 0xff, (size_t)slots * sizeof(CellRec), c->stream));
  LM_LAUNCH(c, k_map_keys, nb, 256, 0, m.cat, n, m.dev, m.keys, m.vals);
  int end_bit = (md.n < n) ? 64 : total_bits;
  size_t tmp = c->cub_tmp_bytes;
  LM_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, m.keys, m.keys_alt, m.vals, m.vals_alt, n, 0, end_bit,
                                          c->stream));
  c->launches++;
  const int nf = md.n;
  LM_LAUNCH(c, k_map_gather, div_up(nf, 256), 256, 0, m.cat, m.vals_alt, nf, m.sorted);
  LM_LAUNCH(c, k_head_flags, div_up(nf, 256), 256, 0, m.keys_alt, nf, m.flags);
  tmp = c->cub_tmp_bytes;
  int* d_nsel = (int*)c->d_bbox + 7;
  LM_CUDA(cub::DeviceSelect::Flagged(c->cub_tmp, tmp, thrust::counting_iterator<int>(0), m.flags, m.fine_start, d_nsel,
                                     nf, c->stream));
  c->launches++;
  LM_LAUNCH(c, k_cell_insert, div_up(nf, 256), 256, 0, m.keys_alt, m.fine_start, d_nsel, nf, m.dev, m.table,
            md.table_mask);
  LM_CUDA(cudaGetLastError());
  m.ready = true;
  return LMSF_OK;
}

}  // namespace lm
