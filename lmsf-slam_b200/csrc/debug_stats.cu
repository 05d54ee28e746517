// debug_stats.cu — instrumented copy of the kNN search (compiled with LMSF_KNN_STATS) used to tune it:
// work counters summed over a batch of queries.  Not part of the documented ABI; exported as lmsf_debug_knn_stats for
// scripts under profiles/.
//   out[0] L0 cell lookups   out[1] L1 cells visited   out[2] candidates scanned   out[3] segments
//   out[8..11] sweeps started in state START / BALL / SPARSE1 / SPARSE2       out[16] queries, out[17] with 5 neighbours
#define LMSF_KNN_STATS 1
#include "common.cuh"
#include "knn.cuh"

namespace lm {
__global__ void __launch_bounds__(KG_BLOCK) k_knn_stats(MapView mv, const float* __restrict__ q, int nq,
                                                        unsigned long long* __restrict__ out) {
  __shared__ int s_seg[KQ_SMEM_INTS];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  KqTop top;
  const int n = kq_knn5<false>(mv, kq_list(s_seg), q[3 * i], q[3 * i + 1], q[3 * i + 2], nullptr, top);
  atomicAdd(&out[16], 1ull);
  if (n == 5) atomicAdd(&out[17], 1ull);
}
__global__ void k_knn_stats_collect(unsigned long long* __restrict__ out, int reset) {
  const int i = threadIdx.x;
  if (i < 16) {
    if (reset)
      g_knn_stat[i] = 0ull;
    else
      out[i] = g_knn_stat[i];
  }
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
  char* buf = nullptr;
  LM_TRY(hook_scratch(c, (size_t)nq * 12 + 28 * 8 + 64, (void**)&buf));
  unsigned long long* d_o = (unsigned long long*)buf;
  float* d_q = (float*)(buf + 28 * 8);
  LM_CUDA(cudaMemsetAsync(d_o, 0, 28 * 8, c->stream));
  LM_CUDA(cudaMemcpyAsync(d_q, q_xyz, (size_t)nq * 12, cudaMemcpyHostToDevice, c->stream));
  k_knn_stats_collect<<<1, 32, 0, c->stream>>>(d_o, 1);
  k_knn_stats<<<div_up(nq, KG_BLOCK), KG_BLOCK, 0, c->stream>>>(v, d_q, nq, d_o);
  k_knn_stats_collect<<<1, 32, 0, c->stream>>>(d_o, 0);
  LM_CUDA(cudaMemcpyAsync(out, d_o, 28 * 8, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  return LMSF_OK;
}
