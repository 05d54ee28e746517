// voxel.cu — voxel-grid downsampling: 64-bit voxel keys -> radix sort -> segmented centroid.
//
// Replaces the pcl::VoxelGrid behind FilterBase::Filter
// (Algorithm/PointClouds/processing/Filter/filter_base.hpp:34-45, voxel_grid.hpp:25-29,
// factory/processing/pointcloud/filter/filter_factory.hpp:36-41).  PCL semantics kept:
// fp32 index math ijk = int(floor(p*inv_leaf) - float(min_b)), the int32 overflow
// guard (output = input), one output per occupied voxel in ascending linear-index
// order, mean of all four fields.  The packed key (k | j | i, k major) orders
// voxels exactly like PCL's i + j*dx + k*dx*dy.  Within a voxel the fp32 sums run
// in ascending point index (stable radix sort), one warp per voxel.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"

namespace lm {

struct VoxelParams {
  float inv;
  int minb[3];
  int div[3];
  int bits[3];
  int overflow;
  int n_finite;
  int n_vox;
  int pad;
};

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_bbox_init(unsigned* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  if (threadIdx.x >= 3 && threadIdx.x < 6) bbox[threadIdx.x] = 0u;
  if (threadIdx.x == 6) bbox[6] = 0u;
}

// getMinMax3D over the finite points + their count
__global__ void __launch_bounds__(256) k_bbox(const float4* __restrict__ in, int n, unsigned* __restrict__ bbox) {
  unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  unsigned cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = in[i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      unsigned ux = f2ord(p.x), uy = f2ord(p.y), uz = f2ord(p.z);
      mn[0] = min(mn[0], ux);
      mn[1] = min(mn[1], uy);
      mn[2] = min(mn[2], uz);
      mx[0] = max(mx[0], ux);
      mx[1] = max(mx[1], uy);
      mx[2] = max(mx[2], uz);
      ++cnt;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
      mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  // one set of atomics per block, not per warp: 592 x 8 warps x 7 atomics on seven addresses serialised in L2 and made
  // this the slowest kernel of the voxel path (29 us for 16 MB, profiles/r2n_voxel_ncu_summary.txt)
  __shared__ unsigned s_v[8][7];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      s_v[warp][a] = mn[a];
      s_v[warp][3 + a] = mx[a];
    }
    s_v[warp][6] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const int a = threadIdx.x;
    unsigned v = s_v[0][a];
    for (int w = 1; w < 8; ++w) v = a < 3 ? min(v, s_v[w][a]) : (a < 6 ? max(v, s_v[w][a]) : v + s_v[w][a]);
    if (a < 3)
      atomicMin(&bbox[a], v);
    else if (a < 6)
      atomicMax(&bbox[a], v);
    else
      atomicAdd(&bbox[6], v);
  }
}

__device__ __forceinline__ int bits_for(int div) {
  int b = 0;
  while ((1 << b) < div && b < 31) ++b;
  return b;
}

__global__ void k_voxel_params(const unsigned* __restrict__ bbox, float leaf, VoxelParams* __restrict__ vp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  VoxelParams p;
  p.inv = 1.0f / leaf;
  p.n_finite = (int)bbox[6];
  p.overflow = 0;
  p.n_vox = 0;
  p.pad = 0;
  if (p.n_finite == 0) {
    for (int a = 0; a < 3; ++a) p.minb[a] = p.div[a] = p.bits[a] = 0;
    *vp = p;
    return;
  }
  long long d[3];
  for (int a = 0; a < 3; ++a) {
    float mn = ord2f(bbox[a]), mx = ord2f(bbox[3 + a]);
    d[a] = (long long)((mx - mn) * p.inv) + 1;
    p.minb[a] = (int)floorf(mn * p.inv);
    int maxb = (int)floorf(mx * p.inv);
    p.div[a] = maxb - p.minb[a] + 1;
    p.bits[a] = bits_for(p.div[a]);
  }
  // "Leaf size is too small for the input dataset. Integer indices would overflow."
  if (d[0] * d[1] * d[2] > 2147483647LL) p.overflow = 1;
  *vp = p;
}

__global__ void __launch_bounds__(256) k_voxel_keys(const float4* __restrict__ in, int n,
                                                    const VoxelParams* __restrict__ vp,
                                                    unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float inv = vp->inv;
  float4 p = in[i];
  unsigned long long key = ~0ull;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    int i0 = (int)(floorf(p.x * inv) - (float)vp->minb[0]);
    int i1 = (int)(floorf(p.y * inv) - (float)vp->minb[1]);
    int i2 = (int)(floorf(p.z * inv) - (float)vp->minb[2]);
    key = ((((unsigned long long)(unsigned)i2 << vp->bits[1]) | (unsigned long long)(unsigned)i1) << vp->bits[0]) |
          (unsigned long long)(unsigned)i0;
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void __launch_bounds__(256) k_head_flags(const unsigned long long* __restrict__ keys, int n,
                                                    uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one warp per voxel: fp32 sums in ascending point index, then / float(count)
__global__ void __launch_bounds__(256) k_voxel_centroid(const float4* __restrict__ in, const int* __restrict__ vals,
                                                        const int* __restrict__ heads, int n_vox, int n_finite,
                                                        float4* __restrict__ out, int* __restrict__ member) {
  int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (v >= n_vox) return;
  int start = heads[v];
  int end = (v + 1 < n_vox) ? heads[v + 1] : n_finite;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (int base = start; base < end; base += 32) {
    int i = base + lane;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < end) {
      int src = vals[i];
      p = in[src];
      if (member) member[src] = v;
    }
    int cnt = min(32, end - base);
    for (int l = 0; l < cnt; ++l) {
      sx += __shfl_sync(0xffffffffu, p.x, l);
      sy += __shfl_sync(0xffffffffu, p.y, l);
      sz += __shfl_sync(0xffffffffu, p.z, l);
      si += __shfl_sync(0xffffffffu, p.w, l);
    }
  }
  if (lane == 0) {
    float c = (float)(end - start);
    out[v] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
}

__global__ void __launch_bounds__(256) k_fill_member(int* __restrict__ member, int n, int mode) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) member[i] = mode ? i : -1;
}

// ---- PointCloudCommonProcess's order-preserving point filters (row f4) ---------------------------------------
// mode 0: pcl::removeNaNFromPointCloud — keep points whose x, y and z are finite
// mode 1: DistanceFilter::Filter (Filter/distance_filter.hpp:24-44) — d = Vector3f norm (fp32 (x x + y y) + z z,
//         fp32 sqrt) widened to double, keep near < d < far against the float thresholds widened to double
__global__ void __launch_bounds__(256) k_keep_flags(const float4* __restrict__ in, int n, int mode, float near_t,
                                                    float far_t, uint8_t* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = in[i];
  bool keep;
  if (mode == 0) {
    keep = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
  } else {
    float s = p.x * p.x + p.y * p.y;
    s = s + p.z * p.z;
    double d = (double)sqrtf(s);
    keep = d > (double)near_t && d < (double)far_t;
  }
  flags[i] = keep ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_gather_points(const float4* __restrict__ in, const int* __restrict__ sel,
                                                       int n, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[sel[i]];
}

// stable compaction of d_in[0..n) by the predicate `mode` into d_out; *n_out read back (one sync)
int filter_run(Ctx* c, const float4* d_in, int n, int mode, float near_t, float far_t, float4* d_out, int* n_out) {
  if (n > c->vox_cap) return LMSF_ERR_CAPACITY;
  *n_out = 0;
  if (n == 0) return LMSF_OK;
  const int nb = div_up(n, 256);
  LM_LAUNCH(c, k_keep_flags, nb, 256, 0, d_in, n, mode, near_t, far_t, c->v_flags);
  size_t tmp = c->cub_tmp_bytes;
  int* d_nsel = (int*)c->d_bbox + 7;
  LM_CUDA(cub::DeviceSelect::Flagged(c->cub_tmp, tmp, thrust::counting_iterator<int>(0), c->v_flags, c->v_heads, d_nsel,
                                     n, c->stream));
  c->launches++;
  LM_CUDA(cudaMemcpyAsync(c->h_ints, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  const int m = c->h_ints[0];
  if (m > 0) LM_LAUNCH(c, k_gather_points, div_up(m, 256), 256, 0, d_in, c->v_heads, m, d_out);
  LM_CUDA(cudaGetLastError());
  *n_out = m;
  return LMSF_OK;
}

int voxel_alloc(Ctx* c) {
  int cap = c->prm.max_points > c->prm.max_map_points ? c->prm.max_points : c->prm.max_map_points;
  c->vox_cap = cap;
  size_t n = (size_t)cap;
  LM_CUDA(cudaMalloc(&c->v_keys, n * 8));
  LM_CUDA(cudaMalloc(&c->v_keys_alt, n * 8));
  LM_CUDA(cudaMalloc(&c->v_vals, n * 4));
  LM_CUDA(cudaMalloc(&c->v_vals_alt, n * 4));
  LM_CUDA(cudaMalloc(&c->v_heads, (n + 1) * 4));
  LM_CUDA(cudaMalloc(&c->v_flags, n));
  LM_CUDA(cudaMalloc(&c->v_in, n * sizeof(float4)));
  LM_CUDA(cudaMalloc(&c->v_out, n * sizeof(float4)));
  LM_CUDA(cudaMalloc(&c->v_member, n * 4));
  LM_CUDA(cudaMalloc(&c->v_params, sizeof(VoxelParams)));
  LM_CUDA(cudaMalloc(&c->d_bbox, 8 * sizeof(unsigned)));
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, cap, 0, 64, c->stream);
  cub::DeviceSelect::Flagged(nullptr, b2, thrust::counting_iterator<int>(0), (uint8_t*)nullptr, (int*)nullptr,
                             (int*)nullptr, cap, c->stream);
  c->cub_tmp_bytes = (b1 > b2 ? b1 : b2) + 256;
  LM_CUDA(cudaMalloc(&c->cub_tmp, c->cub_tmp_bytes));
  return LMSF_OK;
}

void voxel_free(Ctx* c) {
  cudaFree(c->v_keys);
  cudaFree(c->v_keys_alt);
  cudaFree(c->v_vals);
  cudaFree(c->v_vals_alt);
  cudaFree(c->v_heads);
  cudaFree(c->v_flags);
  cudaFree(c->v_in);
  cudaFree(c->v_out);
  cudaFree(c->v_member);
  cudaFree(c->v_params);
  cudaFree(c->d_bbox);
  cudaFree(c->cub_tmp);
}

int voxel_run(Ctx* c, const float4* d_in, int n, float leaf, float4* d_out, int* n_out, int* d_member) {
  if (n > c->vox_cap) return LMSF_ERR_CAPACITY;
  StageScope scope(c, LMSF_STAGE_VOXEL);
  *n_out = 0;
  if (n == 0) return LMSF_OK;
  VoxelParams* vp = (VoxelParams*)c->v_params;
  VoxelParams* hp = (VoxelParams*)c->h_ints;  // pinned staging, 64 ints
  const int nb = div_up(n, 256);
  LM_LAUNCH(c, k_bbox_init, 1, 32, 0, c->d_bbox);
  LM_LAUNCH(c, k_bbox, nb < 592 ? nb : 592, 256, 0, d_in, n, c->d_bbox);
  LM_LAUNCH(c, k_voxel_params, 1, 32, 0, c->d_bbox, leaf, vp);
  LM_CUDA(cudaMemcpyAsync(hp, vp, sizeof(VoxelParams), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  VoxelParams P = *hp;
  if (P.n_finite == 0) {
    if (d_member) LM_LAUNCH(c, k_fill_member, nb, 256, 0, d_member, n, 0);
    return LMSF_OK;
  }
  if (P.overflow) {
    LM_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    if (d_member) LM_LAUNCH(c, k_fill_member, nb, 256, 0, d_member, n, 1);
    *n_out = n;
    return LMSF_OK;
  }
  if (d_member && P.n_finite < n) LM_LAUNCH(c, k_fill_member, nb, 256, 0, d_member, n, 0);
  LM_LAUNCH(c, k_voxel_keys, nb, 256, 0, d_in, n, vp, c->v_keys, c->v_vals);
  int end_bit = (P.n_finite < n) ? 64 : (P.bits[0] + P.bits[1] + P.bits[2]);
  if (end_bit < 1) end_bit = 1;
  size_t tmp = c->cub_tmp_bytes;
  LM_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, c->v_keys, c->v_keys_alt, c->v_vals, c->v_vals_alt, n, 0,
                                          end_bit, c->stream));
  c->launches++;
  const int nf = P.n_finite;
  LM_LAUNCH(c, k_head_flags, div_up(nf, 256), 256, 0, c->v_keys_alt, nf, c->v_flags);
  tmp = c->cub_tmp_bytes;
  int* d_nsel = (int*)c->d_bbox + 7;
  LM_CUDA(cub::DeviceSelect::Flagged(c->cub_tmp, tmp, thrust::counting_iterator<int>(0), c->v_flags, c->v_heads,
                                     d_nsel, nf, c->stream));
  c->launches++;
  LM_CUDA(cudaMemcpyAsync(c->h_ints, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  int nv = c->h_ints[0];
  LM_LAUNCH(c, k_voxel_centroid, div_up(nv * 32, 256), 256, 0, d_in, c->v_vals_alt, c->v_heads, nv, nf, d_out,
            d_member);
  LM_CUDA(cudaGetLastError());
  *n_out = nv;
  return LMSF_OK;
}

}  // namespace lm
