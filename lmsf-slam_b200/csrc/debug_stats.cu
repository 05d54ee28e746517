// debug_stats.cu — instrumented copy of the kNN search (compiled with LMSF_KNN_STATS) used to tune it:
// per exit pass (A / B / C) the number of queries and the candidates scanned, box tests and cell lookups.
// Not part of the documented ABI; exported as lmsf_debug_knn_stats for scripts under profiles/.
#define LMSF_KNN_STATS 1
#include "common.cuh"
#include "knn.cuh"

namespace lm {
__global__ void __launch_bounds__(128) k_knn_stats(MapView mv, const float* __restrict__ q, int nq,
                                                   unsigned long long* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  Top5 nb;
  KnnStats ks;
  ks.cand = ks.boxes = ks.lookups = 0;
  ks.level = 0;
  knn5(mv, q[3 * i], q[3 * i + 1], q[3 * i + 2], nb, nullptr, nullptr, ks);
  int l = ks.level;
  atomicAdd(&out[l * 4 + 0], 1ull);
  atomicAdd(&out[l * 4 + 1], (unsigned long long)ks.cand);
  atomicAdd(&out[l * 4 + 2], (unsigned long long)ks.boxes);
  atomicAdd(&out[l * 4 + 3], (unsigned long long)ks.lookups);
  if (nb.full()) atomicAdd(&out[20 + l], 1ull);
}
}  // namespace lm

using namespace lm;
extern "C" int lmsf_debug_knn_stats(lmsf_ctx* c, int kind, const float* q_xyz, int nq, unsigned long long out[28]) {
  if (!c || kind < 0 || kind > 1 || nq <= 0) return LMSF_ERR_INVALID;
  if (cudaSetDevice(c->device) != cudaSuccess) return LMSF_ERR_NO_DEVICE;
  MapIndex& m = c->map[kind];
  if (!m.ready) return LMSF_ERR_STATE;
  MapView v;
  v.sorted = m.sorted;
  v.table = m.table;
  v.l1 = m.l1;
  v.l2_start = m.l2_start;
  v.dev = m.dev;
  float* d_q = nullptr;
  unsigned long long* d_o = nullptr;
  LM_CUDA(cudaMalloc(&d_q, (size_t)nq * 12));
  LM_CUDA(cudaMalloc(&d_o, 28 * 8));
  LM_CUDA(cudaMemsetAsync(d_o, 0, 28 * 8, c->stream));
  LM_CUDA(cudaMemcpyAsync(d_q, q_xyz, (size_t)nq * 12, cudaMemcpyHostToDevice, c->stream));
  k_knn_stats<<<div_up(nq, 128), 128, 0, c->stream>>>(v, d_q, nq, d_o);
  LM_CUDA(cudaMemcpyAsync(out, d_o, 28 * 8, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_q);
  cudaFree(d_o);
  return LMSF_OK;
}
