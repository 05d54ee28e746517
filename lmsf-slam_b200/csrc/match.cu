// match.cu — correspondence search, residuals/Jacobians, normal-equation reduction and the
// device-resident Gauss-Newton / Huber-LM loops.
//
//   k_knn       one thread per scan feature, in ring order (permutation from the feature extraction):
//               pointAssociateToMap (fp64 -> fp32, edgeSurfFeatureRegistration.hpp:342-350) and the exact
//               5-NN over the local-map grid (knn.cuh), seeded with the previous outer iteration's neighbours.
//               (k_assoc + a cell sort of the queries replace the ring order for caller-supplied features.)
//   k_fit       EdgeFeatureMatch::Match (FeatureMatch/EdgeFeatureMatch.hpp:33-87: 3x3 scatter,
//               symmetric eigen, line test, point-to-line residual and gradient) or
//               SurfFeatureMatch::Match (surfFeatureMatch.hpp:32-87: 5x3 least squares plane,
//               0.2 m validity, signed distance truncated to float), then the 1x6 Jacobian row of
//               the selected solver and a warp-shuffle + block-tree reduction of J^T J / J^T r.
//               The last block to finish sums the per-block partials in a fixed order and runs the
//               6x6 step: GNOptimization (edgeSurfFeatureRegistration.hpp:218-330) or the first
//               trust-region proposal of the Ceres-style Huber-LM.
//   k_lm_eval   re-evaluates the stored correspondences at the LM candidate (se3PointEdgeFactor
//               ceres_factor/edge_factor.hpp:33-61, se3PointSurfFactor surf_factor.hpp:32-56,
//               HuberLoss(0.1) ceres_edgeSurfFeatureRegistration.hpp:107), same reduction; its last
//               block accepts / rejects the step and proposes the next one.
// No host round trip happens inside a solve: iteration control lives in SolveState on the device,
// and every kernel of the pre-enqueued sequence exits at once when its phase is over.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include <cub/cub.cuh>

#include "common.cuh"
#include "knn.cuh"

namespace lm {

#ifndef LMSF_MATCH_BLOCK
#define LMSF_MATCH_BLOCK 256  // 128: 1 570, 256 with two evaluation blocks per SM: 1 600-1 616 scans/s (r3n, r3o)
#endif
static constexpr int MATCH_BLOCK = LMSF_MATCH_BLOCK;
static constexpr int MATCH_WARPS = MATCH_BLOCK / 32;
#ifndef LMSF_EVAL_PREFETCH
#define LMSF_EVAL_PREFETCH 2  // records per thread that k_lm_eval fetches before its griddepcontrol.wait (2 x 148 x 256 threads: all of an HDL-64 sweep)
#endif
static constexpr int EVAL_PREFETCH = LMSF_EVAL_PREFETCH;

struct SolveParams {
  int solver;        // LMSF_SOLVER_*
  int iter;          // GN iteration index / LM outer index
  int lm_max_iters;  // 4
  int gn_min_rows;   // 10
  double huber;      // 0.1
  int poll_ns;       // k_solve: sleep between two polls of the barrier word (0 = spin)
};

struct MapPair {
  MapView edge, surf;
  const float4* edge_cat;  // map clouds in map order (neighbour coordinates by original index)
  const float4* surf_cat;
};

// ------------------------------------------------------------------ geometry of one match
// EdgeFeatureMatch::Match :44-84 given the five neighbours (ascending) and the fp32 world point
__device__ __forceinline__ bool fit_edge(const float4* __restrict__ cat, const Top5& nb, float px, float py, float pz,
                                         d3& n_out, double& r_out, d3& a_out, d3& b_out) {
  d3 pt[5];
  d3 c = mk3(0, 0, 0);
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    float4 q = __ldg(&cat[nb.id[j]]);
    pt[j] = mk3((double)q.x, (double)q.y, (double)q.z);
    c = add3(c, pt[j]);
  }
  c = mk3(c.x / 5.0, c.y / 5.0, c.z / 5.0);
  double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    d3 z = sub3(pt[j], c);
    double zz[3] = {z.x, z.y, z.z};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) S[r * 3 + cc] = S[r * 3 + cc] + zz[r] * zz[cc];
  }
  double w[3], V[9];
  jacobi_eig<3>(S, w, V);
  if (!(w[2] > 3 * w[1])) return false;
  d3 u = mk3(V[2], V[5], V[8]);
  d3 a = add3(scl3(0.1, u), c);
  d3 b = add3(scl3(-0.1, u), c);
  d3 p = mk3((double)px, (double)py, (double)pz);
  d3 nu = crs3(sub3(p, a), sub3(p, b));
  d3 de = sub3(a, b);
  double den = nrm3(de);
  r_out = nrm3(nu) / den;
  d3 g = crs3(de, nu);
  double gn = nrm3(g);
  n_out = (gn > 0.0) ? mk3(g.x / gn, g.y / gn, g.z / gn) : g;
  a_out = a;
  b_out = b;
  return true;
}

// SurfFeatureMatch::Match :44-85
__device__ __forceinline__ bool fit_surf(const float4* __restrict__ cat, const Top5& nb, float px, float py, float pz,
                                         d3& n_out, double& D_out, double& r_out) {
  double A[15], B[5] = {-1, -1, -1, -1, -1}, nn[3];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    float4 q = __ldg(&cat[nb.id[j]]);
    A[j * 3 + 0] = q.x;
    A[j * 3 + 1] = q.y;
    A[j * 3 + 2] = q.z;
  }
  cpqr_solve<5, 3>(A, B, nn);
  d3 n = mk3(nn[0], nn[1], nn[2]);
  double len = nrm3(n);
  double D = 1 / len;
  n = mk3(n.x / len, n.y / len, n.z / len);
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    if (fabs(n.x * A[j * 3 + 0] + n.y * A[j * 3 + 1] + n.z * A[j * 3 + 2] + D) > 0.2) return false;
  }
  d3 p = mk3((double)px, (double)py, (double)pz);
  float distance = (float)(dot3(n, p) + D);
  r_out = fabs((double)distance);
  if (distance >= 0) {
    n_out = n;
    D_out = D;
  } else {
    n_out = mk3(-n.x, -n.y, -n.z);
    D_out = -D;
  }
  return true;
}

// ------------------------------------------------------------------ accumulation
struct Acc {
  double v[LM_NSUM];  // H upper triangle row-major (21), g (6), cost, rows, edge rows
};

__device__ __forceinline__ void acc_zero(Acc& a) {
#pragma unroll
  for (int i = 0; i < LM_NSUM; ++i) a.v[i] = 0.0;
}

__device__ __forceinline__ void acc_row(Acc& a, const double* J, double r, double cost) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) a.v[k++] += J[i] * J[j];
#pragma unroll
  for (int i = 0; i < 6; ++i) a.v[21 + i] += J[i] * r;
  a.v[27] += cost;
  a.v[28] += 1.0;
}

// ceres::HuberLoss + Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho')
__device__ __forceinline__ void huber_apply(double a, double r, double& rho0, double& scale) {
  double s = r * r, b = a * a;
  if (s > b) {
    double sr = sqrt(s);
    rho0 = 2 * a * sr - b;
    double rho1 = fmax(DBL_MIN, a / sr);
    scale = sqrt(rho1);
  } else {
    rho0 = s;
    scale = 1.0;
  }
}

// se3PointEdgeFactor (edge_factor.hpp:39-57): residual and local Jacobian at (q,t)
__device__ __forceinline__ void edge_factor(const quat& q, d3 t, d3 pl, d3 a, d3 b, double huber, Acc& acc) {
  d3 lp = add3(qrot(q, pl), t);
  d3 nu = crs3(sub3(lp, a), sub3(lp, b));
  d3 de = sub3(a, b);
  double den = nrm3(de), nun = nrm3(nu);
  double r = nun / den;
  d3 u = mk3(nu.x / nun, nu.y / nun, nu.z / nun);
  d3 g = crs3(de, u);
  d3 jr = crs3(lp, g);
  double J[6] = {jr.x / den, jr.y / den, jr.z / den, g.x / den, g.y / den, g.z / den};
  double rho0, sc;
  huber_apply(huber, r, rho0, sc);
#pragma unroll
  for (int i = 0; i < 6; ++i) J[i] = sc * J[i];
  acc_row(acc, J, sc * r, 0.5 * rho0);
}

// se3PointSurfFactor (surf_factor.hpp:37-52)
__device__ __forceinline__ void surf_factor(const quat& q, d3 t, d3 pl, d3 n, double D, double huber, Acc& acc) {
  d3 lp = add3(qrot(q, pl), t);
  double r = dot3(n, lp) + D;
  d3 jr = crs3(lp, n);
  double J[6] = {jr.x, jr.y, jr.z, n.x, n.y, n.z};
  double rho0, sc;
  huber_apply(huber, r, rho0, sc);
#pragma unroll
  for (int i = 0; i < 6; ++i) J[i] = sc * J[i];
  acc_row(acc, J, sc * r, 0.5 * rho0);
}

// GNOptimization row (edgeSurfFeatureRegistration.hpp:236-266): J = grad^T [-R skew(p) | I], float residual
__device__ __forceinline__ void gn_row(const double* R, d3 pl, d3 grad, double res, Acc& acc) {
  float residual = (float)res;
  double sk[9] = {0, -pl.z, pl.y, pl.z, 0, -pl.x, -pl.y, pl.x, 0};
  double A[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      A[i * 3 + j] = -(R[i * 3 + 0] * sk[0 * 3 + j] + R[i * 3 + 1] * sk[1 * 3 + j] + R[i * 3 + 2] * sk[2 * 3 + j]);
  double J[6];
#pragma unroll
  for (int j = 0; j < 3; ++j) J[j] = grad.x * A[0 * 3 + j] + grad.y * A[1 * 3 + j] + grad.z * A[2 * 3 + j];
  J[3] = grad.x;
  J[4] = grad.y;
  J[5] = grad.z;
  double rr = residual;
  acc_row(acc, J, rr, 0.5 * rr * rr);
}

// warp-shuffle + block tree; returns true in the last block to finish, with the grid total in `tot`.
// Per-block partials are stored quantity-major (partial[k * gridDim.x + block]) so that the final sum,
// one warp per quantity with lanes striding over the blocks, reads them coalesced.  Every order of
// summation is fixed, so the result is reproducible bit for bit from run to run.
__device__ __forceinline__ bool reduce_grid(const Acc& acc, double* __restrict__ partial, SolveState* __restrict__ st,
                                            double* tot /* shared [LM_NSUM] */) {
  __shared__ double wsum[MATCH_WARPS][LM_NSUM];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nblk = gridDim.x;
  // Warp sums of the LM_NSUM quantities by a TRANSPOSING butterfly: at the exchange over distance d a lane keeps one
  // half of the quantities it still holds and hands the other half to its partner, so the 32 (padded) quantities cost
  // 16 + 8 + 4 + 2 + 1 = 31 exchanges instead of 30 x 5, and lane k ends up with the warp total of quantity k.  The
  // additions are the ones of the plain xor butterfly (v[l] + v[l ^ d] at every level, the same tree for every
  // quantity; fp addition commutes), so every bit of the result is what the 150-exchange form produced.
  {
    static_assert(LM_NSUM <= 32, "one quantity per lane after the transposing butterfly");
    double a[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = (k < LM_NSUM) ? acc.v[k] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const bool hi = (lane & d) != 0;
#pragma unroll
      for (int i = 0; i < d; ++i) {
        const double send = hi ? a[i] : a[i + d];
        const double keep = hi ? a[i + d] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
    }
    if (lane < LM_NSUM) wsum[warp][lane] = a[0];
  }
  __syncthreads();
  if (threadIdx.x < LM_NSUM) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < MATCH_WARPS; ++w) x += wsum[w][threadIdx.x];
    partial[(size_t)threadIdx.x * nblk + blockIdx.x] = x;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(&st->ticket, 1u);
    last = (t == nblk - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  // every warp owns ROWS of the LM_NSUM quantity rows; the loads of all its rows are issued together, so the sum
  // costs about one L2 round trip per 32 blocks instead of one per row (the per-lane order of additions, and with it
  // every bit of the result, is the same as summing row after row)
  constexpr int ROWS = (LM_NSUM + MATCH_WARPS - 1) / MATCH_WARPS;
  double x[ROWS];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) x[j] = 0.0;
  // FINAL_BATCH strides of 32 blocks are loaded before the first of them is added (one L2 round trip per batch instead
  // of one per stride: 296 blocks = 10 strides = 2 batches); the additions stay in ascending block order per lane
  constexpr int FINAL_BATCH = 5;
  for (unsigned b0 = lane; b0 < nblk; b0 += 32 * FINAL_BATCH) {
    double v[FINAL_BATCH][ROWS];
#pragma unroll
    for (int u = 0; u < FINAL_BATCH; ++u) {
      const unsigned b = b0 + 32u * u;
#pragma unroll
      for (int j = 0; j < ROWS; ++j) {
        const int k = warp + j * MATCH_WARPS;
        v[u][j] = (k < LM_NSUM && b < nblk) ? __ldcg(&partial[(size_t)k * nblk + b]) : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < FINAL_BATCH; ++u) {
      if (b0 + 32u * u < nblk) {
#pragma unroll
        for (int j = 0; j < ROWS; ++j) x[j] += v[u][j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    double y = x[j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) y += __shfl_xor_sync(0xffffffffu, y, d);
    const int k = warp + j * MATCH_WARPS;
    if (lane == 0 && k < LM_NSUM) tot[k] = y;
  }
  __syncthreads();
  if (threadIdx.x == 0) st->ticket = 0u;
  return true;
}

// block-cooperative copy of the solver state between global and shared memory (8-byte words)
static_assert(sizeof(SolveState) % 8 == 0, "SolveState is copied as 8-byte words");
__device__ __forceinline__ void state_load(SolveState* sh, const SolveState* st) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(st);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(sh);
  for (int i = threadIdx.x; i < (int)(sizeof(SolveState) / 8); i += blockDim.x) dst[i] = __ldcg(&src[i]);
  __syncthreads();
  if (threadIdx.x == 0) sh->ticket = 0u;  // reduce_grid has just re-armed it; never write a stale count back
}
__device__ __forceinline__ void state_store(SolveState* st, const SolveState* sh) {
  __syncthreads();
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(sh);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(st);
  for (int i = threadIdx.x; i < (int)(sizeof(SolveState) / 8); i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ void unpack_normal(const double* tot, double* H, double* g) {
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      H[i * 6 + j] = tot[k];
      H[j * 6 + i] = tot[k];
      ++k;
    }
  for (int i = 0; i < 6; ++i) g[i] = tot[21 + i];
}

__device__ __forceinline__ double norm7(const double* x) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s += x[i] * x[i];
  return sqrt(s);
}

// ------------------------------------------------------------------ 6x6 steps (one thread)
// GNOptimization :270-329
__device__ void gn_step(SolveState* st, const double* tot, const SolveParams& sp) {
  int rows = (int)tot[28];
  st->gn_iters += 1;
  if (rows < sp.gn_min_rows) return;
  st->cost = tot[27];  // "not enough feature": pose untouched, loop goes on
  double H[36], g[6], mg[6], X[6];
  unpack_normal(tot, H, g);
  for (int i = 0; i < 6; ++i) mg[i] = -g[i];
  cpqr_solve<6, 6>(H, mg, X);
  if (sp.iter == 0) {
    double w[6], V[36], V2[36], Vi[36];
    jacobi_eig<6>(H, w, V);
    for (int i = 0; i < 36; ++i) V2[i] = V[i];
    int deg = 0;
    float thresh = 100;
    for (int i = 5; i >= 0; --i) {  // literal: walks down from the LARGEST eigenvalue (:289-302)
      if (w[i] < thresh) {
        for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.0;
        deg = 1;
      } else {
        break;
      }
    }
    invert6(V, Vi);
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Vi[i * 6 + k] * V2[k * 6 + j];
        st->gn_map[i * 6 + j] = s;
      }
    st->gn_degenerate = deg;
  }
  if (st->gn_degenerate) {
    double Y[6];
    for (int i = 0; i < 6; ++i) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += st->gn_map[i * 6 + k] * X[k];
      Y[i] = s;
    }
    for (int i = 0; i < 6; ++i) X[i] = Y[i];
  }
  st->x[4] += X[3];
  st->x[5] += X[4];
  st->x[6] += X[5];
  d3 dr = mk3(X[0], X[1], X[2]);
  double dn = nrm3(dr);
  d3 axis = dn > 0.0 ? mk3(dr.x / dn, dr.y / dn, dr.z / dn) : dr;
  double ang = dn / 2;  // AngleAxisd(|d|/2, d^) :317
  double sh = sin(0.5 * ang), ch = cos(0.5 * ang);
  quat dq;
  dq.x = sh * axis.x;
  dq.y = sh * axis.y;
  dq.z = sh * axis.z;
  dq.w = ch;
  quat q;
  q.x = st->x[0];
  q.y = st->x[1];
  q.z = st->x[2];
  q.w = st->x[3];
  q = qmul(q, dq);
  st->x[0] = q.x;
  st->x[1] = q.y;
  st->x[2] = q.z;
  st->x[3] = q.w;
  float deltaR = (float)(dn / 2);
  double c3 = X[3] * 100, c4 = X[4] * 100, c5 = X[5] * 100;  // pow(., 2) == exact square
  float deltaT = (float)sqrt(c3 * c3 + c4 * c4 + c5 * c5);
  if (deltaR < 0.0009 && deltaT < 0.05) st->gn_done = 1;
}

__device__ __forceinline__ double grad_max(const double* g) {
  double m = 0;
  for (int j = 0; j < 6; ++j) m = fmax(m, fabs(g[j]));
  return m;
}

// LevenbergMarquardtStrategy::ComputeStep + TrustRegionMinimizer::ComputeTrustRegionStep (Ceres 1.14
// restated, see oracle): propose the next candidate, consuming iterations on invalid steps
__device__ void lm_propose(SolveState* st, const SolveParams& sp) {
  while (true) {
    if (st->lm_iter >= sp.lm_max_iters) {
      st->lm_active = 0;
      return;
    }
    st->lm_iter += 1;
    st->lm_steps_total += 1;
    double Hs[36], gs[6], A[36], y[6], step[6];
    for (int i = 0; i < 6; ++i) {
      gs[i] = st->scale[i] * st->g[i];
      for (int j = 0; j < 6; ++j) Hs[i * 6 + j] = st->scale[i] * st->H[i * 6 + j] * st->scale[j];
    }
    for (int i = 0; i < 36; ++i) A[i] = Hs[i];
    for (int j = 0; j < 6; ++j) {
      double dj = fmin(fmax(Hs[j * 6 + j], 1e-6), 1e32);
      A[j * 6 + j] += dj / st->radius;
    }
    bool ok = spd_solve6(A, gs, y);
    double mcc = 0.0;
    if (ok) {
      for (int j = 0; j < 6; ++j) {
        step[j] = -y[j];
        if (!isfinite(step[j])) ok = false;
      }
    }
    if (ok) {
      double sg = 0, sHs = 0;
      for (int i = 0; i < 6; ++i) {
        sg += step[i] * gs[i];
        double row = 0;
        for (int j = 0; j < 6; ++j) row += Hs[i * 6 + j] * step[j];
        sHs += step[i] * row;
      }
      mcc = -(sg + 0.5 * sHs);
    }
    if (!ok || !(mcc > 0.0)) {
      st->lm_invalid += 1;
      if (st->lm_invalid >= 5) {
        st->lm_active = 0;
        return;
      }
      st->radius *= 0.5;
      if (st->radius < 1e-32) {
        st->lm_active = 0;
        return;
      }
      continue;
    }
    st->lm_invalid = 0;
    double delta[6];
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * st->scale[j];
    se3_plus(st->x, delta, st->cand);
    st->model_change = mcc;
    return;
  }
}

// TrustRegionMinimizer::IterationZero: state of a fresh ceres::Solve at x
__device__ void lm_begin(SolveState* st, const double* tot, const SolveParams& sp) {
  unpack_normal(tot, st->H, st->g);
  st->cost = tot[27];
  st->lm_iter = 0;
  st->lm_invalid = 0;
  st->radius = 1e4;
  st->decrease = 2.0;
  for (int j = 0; j < 6; ++j) st->scale[j] = 1.0 / (1.0 + sqrt(st->H[j * 6 + j]));
  st->x_norm = norm7(st->x);
  st->lm_active = 1;
  if (tot[28] == 0.0 || grad_max(st->g) <= 1e-10) {
    st->lm_active = 0;
    return;
  }
  lm_propose(st, sp);
}

// the part of TrustRegionMinimizer::Minimize after the candidate has been evaluated
__device__ void lm_update(SolveState* st, const double* tot, const SolveParams& sp) {
  double cand_cost = tot[27];
  double diff[7];
  for (int i = 0; i < 7; ++i) diff[i] = st->x[i] - st->cand[i];
  if (norm7(diff) <= 1e-8 * (st->x_norm + 1e-8)) {  // parameter tolerance: candidate not taken
    st->lm_active = 0;
    return;
  }
  double cc = st->cost - cand_cost;
  if (fabs(cc) <= 1e-6 * st->cost) {  // function tolerance: candidate not taken
    st->lm_active = 0;
    return;
  }
  double rho = cc / st->model_change;
  if (rho > 1e-3) {
    for (int i = 0; i < 7; ++i) st->x[i] = st->cand[i];
    st->x_norm = norm7(st->x);
    unpack_normal(tot, st->H, st->g);
    st->cost = cand_cost;
    st->lm_steps_accepted += 1;
    double u = 2.0 * rho - 1.0;
    double f = 1.0 - u * u * u;
    st->radius = st->radius / fmax(1.0 / 3.0, f);
    st->radius = fmin(1e16, st->radius);
    st->decrease = 2.0;
    if (grad_max(st->g) <= 1e-10) {
      st->lm_active = 0;
      return;
    }
  } else {
    st->radius = st->radius / st->decrease;
    st->decrease *= 2.0;
  }
  if (st->radius < 1e-32) {
    st->lm_active = 0;
    return;
  }
  lm_propose(st, sp);
}

// ---- warp-cooperative versions of the LM 6x6 step -----------------------------------------------
// The serial code above is the specification (it is what the oracle does, operation for operation);
// the functions below give the same values — every sum keeps its serial order — but spread the 6x6
// algebra over the lanes of warp 0 of the last block and keep the matrices in shared memory.  A single
// thread walking ~2.5 k dependent fp64 instructions took ~45 us per k_fit launch (ncu: SMs idle for 65 %
// of the kernel); the warp version's critical path is the 6-column Cholesky and the two substitutions.
struct LmScratch {
  double Hs[36], A[36], L[36], gs[6], y[6], x[6], step[6], row[6], trig[4];
  int flag;
};

// se3_plus with the four trigonometric values supplied (cos(theta/2), sin(theta/2), cos(theta), sin(theta))
__device__ __forceinline__ void se3_plus_pre(const double* x, const double* d, const double* trig, double* out) {
  d3 om = mk3(d[0], d[1], d[2]), up = mk3(d[3], d[4], d[5]);
  double theta = nrm3(om);
  double re = trig[0], im;
  if (theta < 1e-10) {
    double t2 = theta * theta, t4 = t2 * t2;
    im = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    im = trig[1] / theta;
  }
  quat dq;
  dq.x = im * om.x;
  dq.y = im * om.y;
  dq.z = im * om.z;
  dq.w = re;
  double J[9];
  if (theta < 1e-10) {
    quat_to_mat(dq, J);
  } else {
    double O[9] = {0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3 + 0] * O[0 * 3 + j] + O[i * 3 + 1] * O[1 * 3 + j] + O[i * 3 + 2] * O[2 * 3 + j];
    double c1 = (1 - trig[2]) / (theta * theta);
    double c2 = (theta - trig[3]) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  d3 dt = mk3(J[0] * up.x + J[1] * up.y + J[2] * up.z, J[3] * up.x + J[4] * up.y + J[5] * up.z,
              J[6] * up.x + J[7] * up.y + J[8] * up.z);
  quat q;
  q.x = x[0];
  q.y = x[1];
  q.z = x[2];
  q.w = x[3];
  quat qp = qmul(dq, q);
  d3 tp = add3(qrot(dq, mk3(x[4], x[5], x[6])), dt);
  out[0] = qp.x;
  out[1] = qp.y;
  out[2] = qp.z;
  out[3] = qp.w;
  out[4] = tp.x;
  out[5] = tp.y;
  out[6] = tp.z;
}

// lm_propose, executed by all 32 lanes of one warp; st and w live in shared memory
__device__ void lm_propose_warp(SolveState* st, LmScratch* w, const SolveParams& sp) {
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  while (true) {
    __syncwarp();
    if (st->lm_iter >= sp.lm_max_iters) {
      __syncwarp();
      if (lane == 0) st->lm_active = 0;
      __syncwarp();
      return;
    }
    __syncwarp();
    if (lane == 0) {
      st->lm_iter += 1;
      st->lm_steps_total += 1;
    }
    for (int e = lane; e < 36; e += 32) {
      int i = e / 6, j = e % 6;
      double h = st->scale[i] * st->H[e] * st->scale[j];
      w->Hs[e] = h;
      w->A[e] = h;
    }
    if (lane < 6) w->gs[lane] = st->scale[lane] * st->g[lane];
    __syncwarp();
    if (lane < 6) {
      double dj = fmin(fmax(w->Hs[lane * 7], 1e-6), 1e32);
      w->A[lane * 7] += dj / st->radius;
    }
    __syncwarp();
    // Cholesky, column by column; lane i owns row i
    bool ok = true;
    for (int j = 0; j < 6; ++j) {
      double s = 0.0;
      if (lane >= j && lane < 6) {
        s = w->A[lane * 6 + j];
        for (int k = 0; k < j; ++k) s -= w->L[lane * 6 + k] * w->L[j * 6 + k];
      }
      double sd = __shfl_sync(FULL, s, j);
      if (!(sd > 0.0)) {
        ok = false;
        break;
      }
      double ljj = sqrt(sd);
      if (lane == j)
        w->L[j * 6 + j] = ljj;
      else if (lane > j && lane < 6)
        w->L[lane * 6 + j] = s / ljj;
      __syncwarp();
    }
    double mcc = 0.0;
    if (ok) {
      // forward substitution, column oriented (row i subtracts L[i][k] y[k] for k ascending)
      double acc = (lane < 6) ? w->gs[lane] : 0.0;
      for (int i = 0; i < 6; ++i) {
        double yi = __shfl_sync(FULL, acc, i) / w->L[i * 6 + i];
        if (lane == i) w->y[i] = yi;
        if (lane > i && lane < 6) acc -= w->L[lane * 6 + i] * yi;
      }
      __syncwarp();
      // back substitution: x[i] = (y[i] - sum_{k>i, ascending} L[k][i] x[k]) / L[i][i]
      if (lane == 0) {
        for (int i = 5; i >= 0; --i) {
          double s = w->y[i];
          for (int k = i + 1; k < 6; ++k) s -= w->L[k * 6 + i] * w->x[k];
          w->x[i] = s / w->L[i * 6 + i];
        }
      }
      __syncwarp();
      double stp = (lane < 6) ? -w->x[lane] : 0.0;
      bool fin = isfinite(stp);
      if (__any_sync(FULL, !fin)) ok = false;
      if (lane < 6) w->step[lane] = stp;
      __syncwarp();
      if (ok) {
        if (lane < 6) {
          double row = 0;
          for (int j = 0; j < 6; ++j) row += w->Hs[lane * 6 + j] * w->step[j];
          w->row[lane] = row;
        }
        __syncwarp();
        double sg = 0, sHs = 0;
        for (int i = 0; i < 6; ++i) {
          sg += w->step[i] * w->gs[i];
          sHs += w->step[i] * w->row[i];
        }
        mcc = -(sg + 0.5 * sHs);
      }
    }
    if (!ok || !(mcc > 0.0)) {  // uniform across the warp: every lane computed the same values
      int give_up = 0;
      __syncwarp();
      if (lane == 0) {
        st->lm_invalid += 1;
        if (st->lm_invalid >= 5) {
          st->lm_active = 0;
          give_up = 1;
        } else {
          st->radius *= 0.5;
          if (st->radius < 1e-32) {
            st->lm_active = 0;
            give_up = 1;
          }
        }
      }
      give_up = __shfl_sync(FULL, give_up, 0);
      __syncwarp();
      if (give_up) return;
      continue;
    }
    // candidate = Plus(x, step * scale); the four trigonometric values come from four lanes
    double delta[6];
    for (int j = 0; j < 6; ++j) delta[j] = w->step[j] * st->scale[j];
    double theta = nrm3(mk3(delta[0], delta[1], delta[2]));
    double half = 0.5 * theta;
    double tv = 0.0;
    if (lane == 0) tv = cos(half);
    if (lane == 1) tv = sin(half);
    if (lane == 2) tv = cos(theta);
    if (lane == 3) tv = sin(theta);
    if (lane < 4) w->trig[lane] = tv;
    __syncwarp();
    if (lane == 0) {
      st->lm_invalid = 0;
      se3_plus_pre(st->x, delta, w->trig, st->cand);
      st->model_change = mcc;
    }
    __syncwarp();
    return;
  }
}

__device__ void lm_begin_warp(SolveState* st, LmScratch* w, const double* tot, const SolveParams& sp) {
  const int lane = threadIdx.x & 31;
  for (int e = lane; e < 36; e += 32) {
    int i = e / 6, j = e % 6;
    int a = i < j ? i : j, b = i < j ? j : i;
    st->H[e] = tot[a * 6 - a * (a - 1) / 2 + (b - a)];  // upper-triangle packing of acc_row
  }
  if (lane < 6) {
    st->g[lane] = tot[21 + lane];
    st->scale[lane] = 1.0 / (1.0 + sqrt(tot[lane * 6 - lane * (lane - 1) / 2]));
  }
  __syncwarp();
  if (lane == 0) {
    st->cost = tot[27];
    st->lm_iter = 0;
    st->lm_invalid = 0;
    st->radius = 1e4;
    st->decrease = 2.0;
    st->x_norm = norm7(st->x);
    st->lm_active = 1;
    if (tot[28] == 0.0 || grad_max(st->g) <= 1e-10) st->lm_active = 0;
  }
  __syncwarp();
  if (!st->lm_active) return;
  lm_propose_warp(st, w, sp);
}

__device__ void lm_update_warp(SolveState* st, LmScratch* w, const double* tot, const SolveParams& sp) {
  const int lane = threadIdx.x & 31;
  int verdict = 0;  // 0 stop, 1 accepted, 2 rejected
  if (lane == 0) {
    double cand_cost = tot[27];
    double diff[7];
    for (int i = 0; i < 7; ++i) diff[i] = st->x[i] - st->cand[i];
    double cc = st->cost - cand_cost;
    if (norm7(diff) <= 1e-8 * (st->x_norm + 1e-8)) {
      st->lm_active = 0;  // parameter tolerance: candidate not taken
    } else if (fabs(cc) <= 1e-6 * st->cost) {
      st->lm_active = 0;  // function tolerance: candidate not taken
    } else {
      double rho = cc / st->model_change;
      if (rho > 1e-3) {
        for (int i = 0; i < 7; ++i) st->x[i] = st->cand[i];
        st->x_norm = norm7(st->x);
        st->cost = cand_cost;
        st->lm_steps_accepted += 1;
        double u = 2.0 * rho - 1.0;
        double f = 1.0 - u * u * u;
        st->radius = st->radius / fmax(1.0 / 3.0, f);
        st->radius = fmin(1e16, st->radius);
        st->decrease = 2.0;
        verdict = 1;
      } else {
        st->radius = st->radius / st->decrease;
        st->decrease *= 2.0;
        verdict = 2;
      }
    }
  }
  verdict = __shfl_sync(0xffffffffu, verdict, 0);
  if (verdict == 0) return;
  if (verdict == 1) {
    for (int e = lane; e < 36; e += 32) {
      int i = e / 6, j = e % 6;
      int a = i < j ? i : j, b = i < j ? j : i;
      st->H[e] = tot[a * 6 - a * (a - 1) / 2 + (b - a)];
    }
    if (lane < 6) st->g[lane] = tot[21 + lane];
    __syncwarp();
    int stop = 0;
    if (lane == 0 && grad_max(st->g) <= 1e-10) {
      st->lm_active = 0;
      stop = 1;
    }
    if (__shfl_sync(0xffffffffu, stop, 0)) return;
  }
  int stop = 0;
  if (lane == 0 && st->radius < 1e-32) {
    st->lm_active = 0;
    stop = 1;
  }
  if (__shfl_sync(0xffffffffu, stop, 0)) return;
  lm_propose_warp(st, w, sp);
}

// ------------------------------------------------------------------ optional in-kernel timing (tuning builds)
#ifdef LMSF_TIMING
__device__ unsigned long long g_dbg[32];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TSTAMP(var) unsigned long long var = gtime()
#define TACC(slot, a, b) atomicAdd(&g_dbg[slot], (b) - (a))
#define TCOUNT(slot) atomicAdd(&g_dbg[slot], 1ull)
#else
#define TSTAMP(var)
#define TACC(slot, a, b)
#define TCOUNT(slot)
#endif

// ------------------------------------------------------------------ kernels
// Programmatic dependent launch (sm_90+): a kernel of the solve chain launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while its predecessor is still draining; it must
// not touch anything the predecessor reads or writes before pdl_wait() (griddepcontrol.wait: the predecessor grid has
// completed and its memory is visible).  pdl_launch_next() then lets the successor be scheduled in turn — behind the
// wait, so at most one successor grid sits resident at a time.  Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_next() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct QueryBufs {
  unsigned* keys;            // [upper] sort keys (kind | L0 | L1 | L2 cell of the world point, top 32 bits)
  int* vals;                 // [upper] feature index
  float4* pw;                // [upper] fp32 world point by feature index
};

// 29-bit cell key of a query for sorting only (locality heuristic: any key is correct): 1 m cell relative to
// a 256 x 256 x 128 m box around the sensor (8 + 8 + 7 bits, clamped) and the 0.25 m sub-cell (6 bits)
struct QueryOrigin {
  int x, y, z;
};
__device__ __forceinline__ unsigned query_cell_key(QueryOrigin o, float x, float y, float z) {
  const float lim = 1.0e5f;
  x = fminf(fmaxf(x, -lim), lim);
  y = fminf(fmaxf(y, -lim), lim);
  z = fminf(fmaxf(z, -lim), lim);
  int ax = (int)floorf(x * 4.0f), ay = (int)floorf(y * 4.0f), az = (int)floorf(z * 4.0f);
  unsigned cx = (unsigned)min(max((ax >> 2) - o.x, 0), 255);
  unsigned cy = (unsigned)min(max((ay >> 2) - o.y, 0), 255);
  unsigned cz = (unsigned)min(max((az >> 2) - o.z, 0), 127);
  unsigned f1 = ((az & 3) << 4) | ((ay & 3) << 2) | (ax & 3);
  return (((((cz << 8) | cy) << 8) | cx) << 6) | f1;
}

// number of map points in the L2 cell that holds (x,y,z): a cheap density probe (scheduling heuristic only)
__device__ __forceinline__ int own_cell_population(const MapView& mv, float x, float y, float z) {
  const MapDev md = *mv.dev;
  if (md.n <= 0 || !(fabsf(x) < 2.0e5f && fabsf(y) < 2.0e5f && fabsf(z) < 2.0e5f)) return 0;
  int ax = (int)floorf(x * 16.0f), ay = (int)floorf(y * 16.0f), az = (int)floorf(z * 16.0f);
  const CellRec* rec = find_cell(mv, md, (ax >> 4) - md.min_c[0], (ay >> 4) - md.min_c[1], (az >> 4) - md.min_c[2]);
  if (!rec) return 0;
  int f1 = (((az >> 2) & 3) << 4) | (((ay >> 2) & 3) << 2) | ((ax >> 2) & 3);
  unsigned long long m1 = rec->mask;
  if (!((m1 >> f1) & 1ull)) return 0;
  int l1 = rec->fine_base + __popcll(m1 & ((1ull << f1) - 1ull));
  int f2 = ((az & 3) << 4) | ((ay & 3) << 2) | (ax & 3);
  const L1Rec lr = ld_l1(mv, l1);
  unsigned long long m2 = lr.mask;
  if (!((m2 >> f2) & 1ull)) return 0;
  int b = lr.first + __popcll(m2 & ((1ull << f2) - 1ull));
  return mv.l2_start[b + 1] - mv.l2_start[b];
}

// pointAssociateToMap (edgeSurfFeatureRegistration.hpp:342-350): world point in fp64, stored as fp32;
// plus the sort key that groups the queries of one map cell into one warp.  Queries whose own map cell is
// (nearly) empty are the expensive ones (wider sweeps): their key sorts them first, so the persistent
// k_knn warps start with them and the cheap dense-region queries fill the tail (longest job first).
__global__ void __launch_bounds__(256) k_assoc(const float4* __restrict__ feat, const int* __restrict__ counts,
                                               SolveState* __restrict__ st, MapPair maps, int has_edge_map,
                                               int has_surf_map, QueryOrigin qorg,
                                               int upper, int solver, int with_keys, QueryBufs qb) {
  pdl_wait();
  pdl_launch_next();
  if (solver == LMSF_SOLVER_GN && st->gn_done) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {  // work queues of the k_knn launch that follows
    st->knn_next = 0;
    st->defer_next = 0;
    st->n_defer = 0;
  }
  const bool in_range = i < upper;
  const int n_e = counts[0], n_s = counts[1];
  unsigned key = 0x80000000u;  // padding sorts behind every live query
  bool heavy = false;
  if (in_range && i < n_e + n_s) {
    quat q;
    q.x = st->x[0];
    q.y = st->x[1];
    q.z = st->x[2];
    q.w = st->x[3];
    d3 t = mk3(st->x[4], st->x[5], st->x[6]);
    float4 f = feat[i];
    d3 pw = add3(qrot(q, mk3((double)f.x, (double)f.y, (double)f.z)), t);
    float4 w = make_float4((float)pw.x, (float)pw.y, (float)pw.z, 0.f);
    qb.pw[i] = w;
    const bool is_edge = i < n_e;
    if (with_keys && (is_edge ? has_edge_map : has_surf_map)) {
      const MapView& mv = is_edge ? maps.edge : maps.surf;
      int pop = own_cell_population(mv, w.x, w.y, w.z);
      heavy = pop < 5;
      // [31] padding  [30] light (own cell populated)  [29] surf  [28..0] cell
      key = query_cell_key(qorg, w.x, w.y, w.z) | (is_edge ? 0u : (1u << 29)) | (heavy ? 0u : (1u << 30));
    } else {
      key = 0u;
      heavy = with_keys != 0;
    }
  }
  if (with_keys) {  // warp-aggregated count of the heavy queries (they occupy sorted positions [0, n_heavy))
    unsigned hb = __ballot_sync(0xffffffffu, heavy);
    if ((threadIdx.x & 31) == 0 && hb) atomicAdd(&st->n_heavy, __popc(hb));
  }
  if (with_keys && in_range) {
    qb.keys[i] = key;
    qb.vals[i] = i;
  }
}

// One deferred query, searched by the whole warp (kw_knn5); all lanes call with the same record (k_knn: world point,
// .w = position | edge flag in bit 30).
__device__ __forceinline__ void knn_sparse_one(const Warp32& x, KwScratch* s, float4 w, const MapPair& maps, int upper,
                                               int* __restrict__ nbr) {
  const int tw = __float_as_int(w.w);
  const int t = tw & 0x3fffffff;
  const bool is_edge = (tw & 0x40000000) != 0;
  KqTop top;
  int n = kw_knn5(x, is_edge ? maps.edge : maps.surf, s, w.x, w.y, w.z, top);
  if (n < 0) {
    // more occupied cells around the query than the item list holds: the complete per-thread search, by lane 0
    // (its segment list lives in the same shared memory)
    if (x.lane == 0) {
      KqList li;
      li.seg = reinterpret_cast<int*>(s);
      li.stride = 1;
      n = kq_knn5<false>(is_edge ? maps.edge : maps.surf, li, w.x, w.y, w.z, nullptr, top);
    }
    n = x.shfl(n, 0);
#pragma unroll
    for (int k = 0; k < 5; ++k) top.k[k] = x.shfl64(top.k[k], 0);
  }
  if (x.lane == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) nbr[k * upper + t] = (n == 5) ? kg_key_id(top.k[k]) : -1;
  }
  x.sync();
}

// exact 5-NN of every query, in processing order (t = position), one thread per query (knn.cuh).  Persistent warps pull
// chunks of 32 consecutive positions from a device-side counter: queries of sparse regions cost several times more
// than queries of dense ones and lie next to each other, so a static block->query map leaves a long tail of heavy
// blocks.
#ifndef KNN_MINBLOCKS
#define KNN_MINBLOCKS 4
#endif
#ifndef KNN_SEEDED_MINBLOCKS
#define KNN_SEEDED_MINBLOCKS 5  // seeded-only kernel, resident blocks per SM: 4: 1 620, 5: 1 632, 6: 1 626, 8 (spills): 1 552;
#endif                          // the general kernel for the seeded pass: 1 608-1 617 scans/s (r3q)
// assoc != 0 (ring-order launches): the kernel also does k_assoc's work for its own query — pointAssociateToMap in
// fp64, stored as fp32 in pw[] for k_fit — so no separate launch is needed.
// seeded != 0 (outer iterations after the first): nbr[] holds the previous iteration's neighbours of the same position;
// the largest of their keys at the query's new place bounds the 5th key from the start.
// SEEDED_ONLY (with DEFER, seeded launches): only the seeded search is compiled in; a query without five usable seeds (it
// had fewer than five neighbours within the radius in the previous pass) goes to the warp-cooperative search.
template <bool DEFER, bool SEEDED_ONLY = false>
__global__ void __launch_bounds__(KG_BLOCK, SEEDED_ONLY ? KNN_SEEDED_MINBLOCKS : KNN_MINBLOCKS) k_knn(const int* __restrict__ perm, float4* __restrict__ pw,
                                             const float4* __restrict__ feat,
                                             const int* __restrict__ counts, SolveState* __restrict__ st,
                                             MapPair maps, int has_edge_map, int has_surf_map, int upper, int solver,
                                             int seeded, int assoc, int chunk, float4* __restrict__ defer,
                                             int* __restrict__ nbr) {
  pdl_wait();
  pdl_launch_next();
  __shared__ int s_seg[KQ_SMEM_INTS];
  if (solver == LMSF_SOLVER_GN && st->gn_done) return;
  const int n_e = counts[0], n_s = counts[1];
  const int lane = threadIdx.x & 31;
  const KqList li = kq_list(s_seg);
  const int live = n_e + n_s;  // positions >= live are padding (sorted) or unset (ring order)
  while (true) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&st->knn_next, chunk);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= live) break;
    const int t = base + lane;
    const int f = (lane < chunk && t < upper && t < live) ? perm[t] : live;
    if (f < live) {
      const bool is_edge = f < n_e;
      float4 w;
      if (assoc) {
        quat q;
        q.x = st->x[0];
        q.y = st->x[1];
        q.z = st->x[2];
        q.w = st->x[3];
        const d3 tr = mk3(st->x[4], st->x[5], st->x[6]);
        const float4 fp = feat[f];
        const d3 pwd = add3(qrot(q, mk3((double)fp.x, (double)fp.y, (double)fp.z)), tr);
        w = make_float4((float)pwd.x, (float)pwd.y, (float)pwd.z, 0.f);
        pw[f] = w;
      } else {
        w = pw[f];
      }
      int n = 0;
      KqTop top;
      if (is_edge ? has_edge_map : has_surf_map) {
        unsigned long long seed[5];
        bool have_seed = false;
        if (seeded) {
          const float4* __restrict__ cat = is_edge ? maps.edge_cat : maps.surf_cat;
          int sid[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) sid[k] = nbr[k * upper + t];
          if (sid[4] >= 0) {  // five or none
            have_seed = true;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              const float4 m = __ldg(&cat[sid[k]]);
              const float dx = m.x - w.x, dy = m.y - w.y, dz = m.z - w.z;
              float r = dx * dx;
              r = r + dy * dy;
              r = r + dz * dz;
              seed[k] = kg_key(r, sid[k]);
            }
          }
        }
        if (SEEDED_ONLY)
          n = have_seed ? kq_knn5_seeded(is_edge ? maps.edge : maps.surf, li, w.x, w.y, w.z, seed, top) : -1;
        else
          n = kq_knn5<DEFER>(is_edge ? maps.edge : maps.surf, li, w.x, w.y, w.z, have_seed ? seed : nullptr, top);
      }
      if (n < 0) {  // a sparse case: left to the warp-cooperative search (one warp per query)
        // one 16-byte record per deferred query: the warp search starts from a single load (position, kind and point)
        defer[atomicAdd(&st->n_defer, 1)] = make_float4(w.x, w.y, w.z, __int_as_float(t | (is_edge ? (int)0x40000000 : 0)));
      } else {
        // -1 x 5 unless five neighbours lie within the search radius: both callers reject such a query
#pragma unroll
        for (int k = 0; k < 5; ++k) nbr[k * upper + t] = (n == 5) ? kg_key_id(top.k[k]) : -1;
      }
    }
  }
}

// the queries k_knn deferred, one per warp (kw_knn5): the launch is enqueued behind every k_knn, its size is only
// known on the device.
__global__ void __launch_bounds__(KG_BLOCK) k_knn_sparse(SolveState* __restrict__ st, MapPair maps, int upper, int solver,
                                                         const float4* __restrict__ defer, int* __restrict__ nbr) {
  pdl_wait();
  pdl_launch_next();
  __shared__ KwScratch scratch[KG_BLOCK / 32];
  if (solver == LMSF_SOLVER_GN && st->gn_done) return;
  const int n_def = st->n_defer;
  if (n_def == 0) return;
  const Warp32 x;
  KwScratch* s = &scratch[threadIdx.x >> 5];
  // every warp starts with the query of its own number (no round trip to the queue head); further ones from the queue
  const int n_warps = gridDim.x * (KG_BLOCK / 32);
  int i = blockIdx.x * (KG_BLOCK / 32) + (threadIdx.x >> 5);
  for (;;) {
    if (i >= n_def) break;
    TSTAMP(t_q0);
    knn_sparse_one(x, s, defer[i], maps, upper, nbr);
#ifdef LMSF_TIMING
    if (x.lane == 0) {  // per-query time of the warp search: max, sum, count, histogram < 5 / 10 / 20 / 40 / >= 40 us
      TSTAMP(t_q1);
      const unsigned long long dt = t_q1 - t_q0;
      atomicMax(&g_dbg[8], dt);
      atomicAdd(&g_dbg[9], dt);
      atomicAdd(&g_dbg[10], 1ull);
      atomicAdd(&g_dbg[11 + (dt < 5000 ? 0 : (dt < 10000 ? 1 : (dt < 20000 ? 2 : (dt < 40000 ? 3 : 4))))], 1ull);
    }
#endif
    if (x.lane == 0) i = n_warps + atomicAdd(&st->defer_next, 1);
    i = x.shfl(i, 0);
  }
}

struct RecBufs {
  double* d;      // [6][stride] edge: a, b   surf: n, D
  float* pl;      // [3][stride] scan point in the LiDAR frame
  uint8_t* kind;  // [stride] 0 = no correspondence, 1 = edge, 2 = surf
  int stride;
};

// The correspondences of this thread's positions (t = first, first + step, ...): fit, residual and Jacobian at (q, tr)
// into acc; for the Huber-LM solver also the per-match records k_lm_eval / k_solve re-evaluate.
template <bool ACC>
__device__ __forceinline__ void fit_positions(const float4* __restrict__ feat, const int* __restrict__ perm,
                                              const float4* __restrict__ pw, const int* __restrict__ nbr, int n_e,
                                              int n_s, const MapPair& maps, int upper, const quat& q, d3 tr,
                                              const double* R, const RecBufs& rb, const SolveParams& sp, Acc& acc) {
  for (int t = blockIdx.x * MATCH_BLOCK + threadIdx.x; t < upper; t += gridDim.x * MATCH_BLOCK) {
    int f = perm[t];
    const bool live = t < n_e + n_s && f < n_e + n_s;  // positions behind the live ones: kind cleared for k_lm_eval
    const bool is_edge = f < n_e;
    Top5 nb;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      nb.id[k] = live ? nbr[k * upper + t] : -1;
      nb.d[k] = 0.f;
    }
    uint8_t kind = 0;
    if (nb.full()) {
      float4 fp = feat[f];
      float4 w = pw[f];
      d3 pl = mk3((double)fp.x, (double)fp.y, (double)fp.z);
      if (is_edge) {
        d3 n, a, b;
        double r;
        if (fit_edge(maps.edge_cat, nb, w.x, w.y, w.z, n, r, a, b)) {
          kind = 1;
          if (ACC) acc.v[29] += 1.0;
          if (sp.solver == LMSF_SOLVER_GN) {
            gn_row(R, pl, n, r, acc);
          } else {
            rb.d[0 * rb.stride + t] = a.x;
            rb.d[1 * rb.stride + t] = a.y;
            rb.d[2 * rb.stride + t] = a.z;
            rb.d[3 * rb.stride + t] = b.x;
            rb.d[4 * rb.stride + t] = b.y;
            rb.d[5 * rb.stride + t] = b.z;
            if (ACC) edge_factor(q, tr, pl, a, b, sp.huber, acc);
          }
        }
      } else {
        d3 n;
        double D, r;
        if (fit_surf(maps.surf_cat, nb, w.x, w.y, w.z, n, D, r)) {
          kind = 2;
          if (sp.solver == LMSF_SOLVER_GN) {
            gn_row(R, pl, n, r, acc);
          } else {
            rb.d[0 * rb.stride + t] = n.x;
            rb.d[1 * rb.stride + t] = n.y;
            rb.d[2 * rb.stride + t] = n.z;
            rb.d[3 * rb.stride + t] = D;
            if (ACC) surf_factor(q, tr, pl, n, D, sp.huber, acc);
          }
        }
      }
      if (kind && sp.solver != LMSF_SOLVER_GN) {
        rb.pl[0 * rb.stride + t] = fp.x;
        rb.pl[1 * rb.stride + t] = fp.y;
        rb.pl[2 * rb.stride + t] = fp.z;
      }
    }
    rb.kind[t] = kind;
  }
}

// fit + residual/Jacobian + reduction, grid-stride over the sorted queries; the last block runs the 6x6 step
__global__ void __launch_bounds__(MATCH_BLOCK) k_fit(const float4* __restrict__ feat, const int* __restrict__ perm,
                                                     const float4* __restrict__ pw, const int* __restrict__ nbr,
                                                     const int* __restrict__ counts, MapPair maps, int upper,
                                                     SolveState* __restrict__ st, RecBufs rb,
                                                     double* __restrict__ partial, SolveParams sp) {
  pdl_wait();
  pdl_launch_next();
  __shared__ double tot[LM_NSUM];
  if (sp.solver == LMSF_SOLVER_GN && st->gn_done) return;
  TSTAMP(t_begin);
  const int n_e = counts[0], n_s = counts[1];
  quat q;
  q.x = st->x[0];
  q.y = st->x[1];
  q.z = st->x[2];
  q.w = st->x[3];
  d3 tr = mk3(st->x[4], st->x[5], st->x[6]);
  double R[9];
  quat_to_mat(q, R);
  Acc acc;
  acc_zero(acc);
  fit_positions<true>(feat, perm, pw, nbr, n_e, n_s, maps, upper, q, tr, R, rb, sp, acc);
  TSTAMP(t_loop);
  if (!reduce_grid(acc, partial, st, tot)) return;
  TSTAMP(t_red);
  // the serial 6x6 step runs on a shared-memory copy of the state (one thread, latency bound:
  // global-memory round trips on every st-> field would dominate it)
  __shared__ SolveState sh;
  __shared__ LmScratch scratch;
  state_load(&sh, st);
  if (threadIdx.x == 0) {
    sh.n_edge_ok = (int)tot[29];
    sh.n_surf_ok = (int)tot[28] - (int)tot[29];
    sh.knn_next = 0;  // work-queue heads of the next k_knn / k_knn_sparse launches (k_assoc does it when it runs)
    sh.defer_next = 0;
    sh.n_defer = 0;
  }
  if (sp.solver == LMSF_SOLVER_GN) {
    if (threadIdx.x == 0) gn_step(&sh, tot, sp);
  } else if (threadIdx.x < 32) {
    lm_begin_warp(&sh, &scratch, tot, sp);
  }
  state_store(st, &sh);
#ifdef LMSF_TIMING
  if (threadIdx.x == 0) {
    TSTAMP(t_end);
    TACC(0, t_begin, t_loop);
    TACC(1, t_loop, t_red);
    TACC(2, t_red, t_end);
    TCOUNT(3);
  }
#endif
}

// the stored correspondences of this thread's positions re-evaluated at (q, tr)
// one stored correspondence as k_fit_records left it (76 bytes: kind, scan point, a / b or n / D)
struct RecRow {
  uint8_t kind;
  float pl[3];
  double d[6];
};

__device__ __forceinline__ void rec_load(RecRow& r, const RecBufs& rb, int t, int upper) {
  r.kind = 0;
  if (t >= upper) return;
  r.kind = rb.kind[t];
  if (!r.kind) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) r.pl[i] = rb.pl[i * rb.stride + t];
#pragma unroll
  for (int i = 0; i < 4; ++i) r.d[i] = rb.d[i * rb.stride + t];
  if (r.kind == 1) {
    r.d[4] = rb.d[4 * rb.stride + t];
    r.d[5] = rb.d[5 * rb.stride + t];
  }
}

template <bool COUNT_EDGES>
__device__ __forceinline__ void rec_eval(const RecRow& r, const quat& q, d3 tr, const SolveParams& sp, Acc& acc) {
  if (!r.kind) return;
  d3 pl = mk3((double)r.pl[0], (double)r.pl[1], (double)r.pl[2]);
  if (r.kind == 1) {
    if (COUNT_EDGES) acc.v[29] += 1.0;
    edge_factor(q, tr, pl, mk3(r.d[0], r.d[1], r.d[2]), mk3(r.d[3], r.d[4], r.d[5]), sp.huber, acc);
  } else {
    surf_factor(q, tr, pl, mk3(r.d[0], r.d[1], r.d[2]), r.d[3], sp.huber, acc);
  }
}

// the positions of this thread from `first` on, in steps of the grid
template <bool COUNT_EDGES = false>
__device__ __forceinline__ void eval_positions(int upper, const quat& q, d3 tr, const RecBufs& rb, const SolveParams& sp,
                                               Acc& acc, int skip = 0) {
  const int step = gridDim.x * MATCH_BLOCK;
  for (int t = blockIdx.x * MATCH_BLOCK + threadIdx.x + skip * step; t < upper; t += step) {
    RecRow r;
    rec_load(r, rb, t, upper);
    rec_eval<COUNT_EDGES>(r, q, tr, sp, acc);
  }
}

// ---- the Huber-LM outer iteration in two launches instead of k_fit ----------------------------------------------
// k_fit carries 30 fp64 accumulators through the fits (168 registers, two resident blocks per SM, every thread four
// features one after the other: a 22 us loop of dependent fp64 chains).  k_fit_records only fits and stores the records
// (one feature per thread, more resident warps), k_lm_eval0 then evaluates the records at x — the same sums k_fit forms,
// in the order of the evaluation grid — and its last block opens the trust-region loop (lm_begin).
__global__ void __launch_bounds__(MATCH_BLOCK) k_fit_records(const float4* __restrict__ feat, const int* __restrict__ perm,
                                                             const float4* __restrict__ pw, const int* __restrict__ nbr,
                                                             const int* __restrict__ counts, MapPair maps, int upper,
                                                             RecBufs rb, SolveParams sp) {
  pdl_wait();
  pdl_launch_next();
  const int n_e = counts[0], n_s = counts[1];
  quat q;
  q.x = q.y = q.z = 0.0;
  q.w = 1.0;
  Acc acc;  // unused (ACC = false)
  fit_positions<false>(feat, perm, pw, nbr, n_e, n_s, maps, upper, q, mk3(0, 0, 0), nullptr, rb, sp, acc);
}

__global__ void __launch_bounds__(MATCH_BLOCK) k_lm_eval0(int upper, SolveState* __restrict__ st, RecBufs rb,
                                                          double* __restrict__ partial, SolveParams sp) {
  pdl_wait();
  pdl_launch_next();
  __shared__ double tot[LM_NSUM];
  quat q;
  q.x = st->x[0];
  q.y = st->x[1];
  q.z = st->x[2];
  q.w = st->x[3];
  d3 tr = mk3(st->x[4], st->x[5], st->x[6]);
  Acc acc;
  acc_zero(acc);
  eval_positions<true>(upper, q, tr, rb, sp, acc);
  if (!reduce_grid(acc, partial, st, tot)) return;
  __shared__ SolveState sh;
  __shared__ LmScratch scratch;
  state_load(&sh, st);
  if (threadIdx.x == 0) {
    sh.n_edge_ok = (int)tot[29];
    sh.n_surf_ok = (int)tot[28] - (int)tot[29];
    sh.knn_next = 0;  // work-queue heads of the next k_knn / k_knn_sparse launches
    sh.defer_next = 0;
    sh.n_defer = 0;
  }
  if (threadIdx.x < 32) lm_begin_warp(&sh, &scratch, tot, sp);
  state_store(st, &sh);
}

// re-evaluate the stored correspondences at the LM candidate; last block accepts / rejects / proposes
__global__ void __launch_bounds__(MATCH_BLOCK) k_lm_eval(int upper, SolveState* __restrict__ st, RecBufs rb,
                                                         double* __restrict__ partial, SolveParams sp, int early_records) {
  // The records were written by k_fit_records, which had completed before the launch that precedes this one passed
  // its own griddepcontrol.wait (k_lm_eval0 or a k_lm_eval) — so this thread's first EVAL_PREFETCH records can be
  // fetched BEFORE the wait, while the previous launch is still in its reduction and 6x6 step: the candidate pose is
  // the only thing the wait protects.  (Ordinary launches: the wait is a no-op and this is a plain load.)  The launch
  // that follows k_fit directly (LMSF_ONE_FIT=1: the records are still being written) passes early_records = 0.
  RecRow pre[EVAL_PREFETCH];
  const int step = gridDim.x * MATCH_BLOCK;
  if (early_records) {
#pragma unroll
    for (int u = 0; u < EVAL_PREFETCH; ++u) rec_load(pre[u], rb, blockIdx.x * MATCH_BLOCK + threadIdx.x + u * step, upper);
  }
  pdl_wait();
  pdl_launch_next();
  if (!early_records) {
#pragma unroll
    for (int u = 0; u < EVAL_PREFETCH; ++u) rec_load(pre[u], rb, blockIdx.x * MATCH_BLOCK + threadIdx.x + u * step, upper);
  }
  __shared__ double tot[LM_NSUM];
  if (!st->lm_active) return;
  TSTAMP(t_begin);
  quat q;
  q.x = st->cand[0];
  q.y = st->cand[1];
  q.z = st->cand[2];
  q.w = st->cand[3];
  d3 tr = mk3(st->cand[4], st->cand[5], st->cand[6]);
  Acc acc;
  acc_zero(acc);
#pragma unroll
  for (int u = 0; u < EVAL_PREFETCH; ++u) rec_eval<false>(pre[u], q, tr, sp, acc);
  eval_positions(upper, q, tr, rb, sp, acc, EVAL_PREFETCH);
  TSTAMP(t_loop);
  if (!reduce_grid(acc, partial, st, tot)) return;
  TSTAMP(t_red);
  __shared__ SolveState sh;
  __shared__ LmScratch scratch;
  state_load(&sh, st);
  if (threadIdx.x < 32) lm_update_warp(&sh, &scratch, tot, sp);
  state_store(st, &sh);
#ifdef LMSF_TIMING
  if (threadIdx.x == 0) {
    TSTAMP(t_end);
    TACC(4, t_begin, t_loop);
    TACC(5, t_loop, t_red);
    TACC(6, t_red, t_end);
    TCOUNT(7);
  }
#endif
}

// ---- k_solve: one outer iteration of the Huber-LM solver in ONE persistent launch ------------------------------------
// k_fit followed by up to lm_max_iters k_lm_eval launches is a chain of dependent kernels that all work on the same
// correspondences: every launch boundary is a drain + a launch gap + (once the trust region has converged) an idle
// launch.  Here the grid stays resident (cooperative launch, <= 2 CTAs per SM): every thread fits its own positions and
// re-evaluates the same positions at each candidate, and between the phases the grid meets at a barrier whose last
// arriver is the block that sums the partials and takes the 6x6 step (exactly reduce_grid + lm_*_warp of the split
// kernels).  The other blocks wait on a generation word, read the verdict (lm_active, candidate pose) and go on or
// leave.  Arithmetic and summation order per phase are those of the split kernels launched with this grid.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(MATCH_BLOCK) k_solve(const float4* __restrict__ feat, const int* __restrict__ perm,
                                                       const float4* __restrict__ pw, const int* __restrict__ nbr,
                                                       const int* __restrict__ counts, MapPair maps, int upper,
                                                       SolveState* __restrict__ st, RecBufs rb,
                                                       double* __restrict__ partial, SolveParams sp,
                                                       unsigned* __restrict__ bar) {
  __shared__ double tot[LM_NSUM];
  __shared__ SolveState sh;
  __shared__ LmScratch scratch;
  __shared__ double s_cand[7];
  __shared__ int s_go;
  __shared__ unsigned s_gen;
  TSTAMP(t_begin);
  if (threadIdx.x == 0) s_gen = ld_acquire_u32(bar);  // no block can release before every block has read this
  const int n_e = counts[0], n_s = counts[1];
  Acc acc;
  {
    quat q;
    q.x = st->x[0];
    q.y = st->x[1];
    q.z = st->x[2];
    q.w = st->x[3];
    const d3 tr = mk3(st->x[4], st->x[5], st->x[6]);
    acc_zero(acc);
    fit_positions<true>(feat, perm, pw, nbr, n_e, n_s, maps, upper, q, tr, nullptr, rb, sp, acc);
  }
  __syncthreads();
  unsigned gen = s_gen;
  TSTAMP(t_loop);
  for (int k = 0;; ++k) {
    // ---- barrier; the last block to arrive takes the step and opens the next generation
    TSTAMP(t_arrive);
    if (reduce_grid(acc, partial, st, tot)) {
      TSTAMP(t_tot);
      state_load(&sh, st);
      TSTAMP(t_ld);
      if (k == 0) {
        if (threadIdx.x == 0) {
          sh.n_edge_ok = (int)tot[29];
          sh.n_surf_ok = (int)tot[28] - (int)tot[29];
          sh.knn_next = 0;  // work queues of the next k_knn launch
          sh.defer_next = 0;
          sh.n_defer = 0;
              }
        if (threadIdx.x < 32) lm_begin_warp(&sh, &scratch, tot, sp);
      } else {
        if (threadIdx.x < 32) lm_update_warp(&sh, &scratch, tot, sp);
      }
      TSTAMP(t_up);
      state_store(st, &sh);
      TSTAMP(t_st);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) st_release_u32(bar, gen + 1u);
#ifdef LMSF_TIMING
      if (threadIdx.x == 0 && k >= 1) {
        TSTAMP(t_done);
        TACC(8, t_arrive, t_tot);
        TACC(9, t_tot, t_done);
        TCOUNT(10);
        TACC(4, t_tot, t_ld);
        TACC(5, t_ld, t_up);
        TACC(6, t_up, t_st);
        TACC(7, t_st, t_done);
      }
#endif
    }
    if (k == sp.lm_max_iters) break;  // the budget is spent: nothing left to wait for
    ++gen;
    TSTAMP(t_arr);
    if (threadIdx.x == 0) {
      while (ld_acquire_u32(bar) != gen)
        if (sp.poll_ns) __nanosleep(sp.poll_ns);
      const int go = __ldcg(&st->lm_active);
      s_go = go;
      if (go) {
#pragma unroll
        for (int i = 0; i < 7; ++i) s_cand[i] = __ldcg(&st->cand[i]);
      }
    }
    __syncthreads();
#ifdef LMSF_TIMING
    TSTAMP(t_rel);
    if (threadIdx.x == 0 && blockIdx.x == 0 && k < 5) TACC(18 + k, t_arr, t_rel);
#endif
    if (!s_go) break;
    quat q;
    q.x = s_cand[0];
    q.y = s_cand[1];
    q.z = s_cand[2];
    q.w = s_cand[3];
    const d3 tr = mk3(s_cand[4], s_cand[5], s_cand[6]);
    acc_zero(acc);
    eval_positions(upper, q, tr, rb, sp, acc);
#ifdef LMSF_TIMING
    TSTAMP(t_ev);
    if (threadIdx.x == 0 && blockIdx.x == 0 && k < 5) {
      TACC(23 + k, t_rel, t_ev);
      TCOUNT(28);
    }

#endif
  }
#ifdef LMSF_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    TSTAMP(t_end);
    TACC(16, t_begin, t_loop);
    TACC(29, t_begin, t_end);
    TCOUNT(17);
  }
#endif
}

struct Pose7 {
  double v[7];
};
// the prior pose travels as a kernel argument: no host-to-device copy in front of the solve
__global__ void k_state_init(SolveState* st, Pose7 pose) {
  if (threadIdx.x != 0) return;
  for (int i = 0; i < 7; ++i) st->x[i] = st->cand[i] = pose.v[i];
  st->cost = 0;
  st->n_edge_ok = st->n_surf_ok = 0;
  st->gn_done = st->gn_degenerate = st->gn_iters = 0;
  st->lm_active = 0;
  st->lm_iter = st->lm_invalid = 0;
  st->lm_steps_total = st->lm_steps_accepted = 0;
  st->ticket = 0u;
  st->knn_next = 0;
  st->defer_next = 0;
  st->n_defer = 0;
  st->n_heavy = 0;
}

// ------------------------------------------------------------------ test hooks
// lmsf_knn5 runs the registration's own pipeline: the per-thread search with deferral, then the deferred queries one
// per warp.  ctr[0] = deferred so far, ctr[1] = work-queue head of the second kernel (zeroed by the host).
__global__ void __launch_bounds__(KG_BLOCK) k_knn_hook(MapView mv, const float* __restrict__ q, int nq,
                                                       int* __restrict__ idx, float* __restrict__ d2,
                                                       int* __restrict__ ctr, int* __restrict__ defer) {
  __shared__ int s_seg[KQ_SMEM_INTS];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  KqTop top;
  const int n = kq_knn5<true>(mv, kq_list(s_seg), q[3 * i], q[3 * i + 1], q[3 * i + 2], nullptr, top);
  if (n < 0) {
    defer[atomicAdd(&ctr[0], 1)] = i;
    return;
  }
  for (int k = 0; k < 5; ++k) {
    const bool in = k < n;
    idx[5 * i + k] = in ? kg_key_id(top.k[k]) : -1;
    d2[5 * i + k] = in ? kg_key_d2(top.k[k]) : __int_as_float(0x7f800000);
  }
}
__global__ void __launch_bounds__(KG_BLOCK) k_knn_hook_sparse(MapView mv, const float* __restrict__ q,
                                                              int* __restrict__ idx, float* __restrict__ d2,
                                                              int* __restrict__ ctr, const int* __restrict__ defer) {
  __shared__ KwScratch scratch[KG_BLOCK / 32];
  const int n_def = ctr[0];
  const Warp32 x;
  KwScratch* s = &scratch[threadIdx.x >> 5];
  while (true) {
    int j = 0;
    if (x.lane == 0) j = atomicAdd(&ctr[1], 1);
    j = x.shfl(j, 0);
    if (j >= n_def) break;
    const int i = defer[j];
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    KqTop top;
    int n = kw_knn5(x, mv, s, qx, qy, qz, top);
    if (n < 0) {  // item list overflow: the complete per-thread search, by lane 0
      if (x.lane == 0) {
        KqList li;
        li.seg = reinterpret_cast<int*>(s);
        li.stride = 1;
        n = kq_knn5<false>(mv, li, qx, qy, qz, nullptr, top);
      }
      n = x.shfl(n, 0);
#pragma unroll
      for (int k = 0; k < 5; ++k) top.k[k] = x.shfl64(top.k[k], 0);
    }
    if (x.lane == 0) {
      for (int k = 0; k < 5; ++k) {
        const bool in = k < n;
        idx[5 * i + k] = in ? kg_key_id(top.k[k]) : -1;
        d2[5 * i + k] = in ? kg_key_d2(top.k[k]) : __int_as_float(0x7f800000);
      }
    }
    x.sync();
  }
}

__global__ void __launch_bounds__(KG_BLOCK) k_match_hook(MapView mv, const float4* __restrict__ cat, int kind,
                                                         const float* __restrict__ q, int nq,
                                                         uint8_t* __restrict__ okv, double* __restrict__ out10) {
  __shared__ int s_seg[KQ_SMEM_INTS];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float px = q[3 * i], py = q[3 * i + 1], pz = q[3 * i + 2];
  KqTop top;
  const int n = kq_knn5<false>(mv, kq_list(s_seg), px, py, pz, nullptr, top);
  double o[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool ok = false;
  if (n == 5) {
    Top5 nb;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      nb.id[k] = kg_key_id(top.k[k]);
      nb.d[k] = kg_key_d2(top.k[k]);
    }
    if (kind == LMSF_KIND_EDGE) {
      d3 nn, a, b;
      double r;
      ok = fit_edge(cat, nb, px, py, pz, nn, r, a, b);
      if (ok) {
        o[0] = nn.x, o[1] = nn.y, o[2] = nn.z, o[3] = r;
        o[4] = a.x, o[5] = a.y, o[6] = a.z, o[7] = b.x, o[8] = b.y, o[9] = b.z;
      }
    } else {
      d3 nn;
      double D, r;
      ok = fit_surf(cat, nb, px, py, pz, nn, D, r);
      if (ok) o[0] = nn.x, o[1] = nn.y, o[2] = nn.z, o[3] = r, o[4] = D;
    }
  }
  okv[i] = ok ? 1 : 0;
  for (int k = 0; k < 10; ++k) out10[10 * i + k] = o[k];
}

// ------------------------------------------------------------------ alignment score (row f2)
// PointCloudAlignmentEvaluate::AlignmentScore (registration/alignEvaluate.hpp:55-87): pcl::transformPointCloud with a
// Matrix4f (fp32: ((m0 x + m1 y) + m2 z) + m3, no FMA), 1-NN in the target, inliers = squared distance <= thresh.
struct Rigid12f {
  float m[12];  // rows 0..2 of the 4x4, row-major
};
__global__ void __launch_bounds__(KG_BLOCK) k_align_score(MapView mv, const float4* __restrict__ pts, int n, Rigid12f T,
                                                          float thresh, double* __restrict__ part_sum,
                                                          int* __restrict__ part_cnt) {
  __shared__ int s_seg[KQ_SMEM_INTS];
  __shared__ double s_sum[4];
  __shared__ int s_cnt[4];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  int in = 0;
  if (i < n) {
    const float4 p = pts[i];
    const float qx = T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z + T.m[3];
    const float qy = T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z + T.m[7];
    const float qz = T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z + T.m[11];
    KqTop top;
    const int found = kq_knn5<false>(mv, kq_list(s_seg), qx, qy, qz, nullptr, top);
    if (found > 0) {
      const float d0 = kg_key_d2(top.k[0]);
      if (d0 <= thresh) {
        v = (double)d0;
        in = 1;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, d);
    in += __shfl_xor_sync(0xffffffffu, in, d);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_sum[warp] = v;
    s_cnt[warp] = in;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    part_sum[blockIdx.x] = ((s_sum[0] + s_sum[1]) + s_sum[2]) + s_sum[3];
    part_cnt[blockIdx.x] = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
  }
}

// fixed-order final sum (reproducible bit for bit)
__global__ void __launch_bounds__(256) k_align_finish(const double* __restrict__ part_sum,
                                                      const int* __restrict__ part_cnt, int nblk,
                                                      double* __restrict__ out_sum, int* __restrict__ out_cnt) {
  __shared__ double s_sum[256];
  __shared__ int s_cnt[256];
  double v = 0.0;
  int in = 0;
  for (int b = threadIdx.x; b < nblk; b += 256) {
    v += part_sum[b];
    in += part_cnt[b];
  }
  s_sum[threadIdx.x] = v;
  s_cnt[threadIdx.x] = in;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + d];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + d];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out_sum = s_sum[0];
    *out_cnt = s_cnt[0];
  }
}

// ------------------------------------------------------------------ host side
#ifndef KNN_GRID_PER_SM
#define KNN_GRID_PER_SM 4
#endif
// launch on the context's stream with the programmatic-stream-serialization attribute (see pdl_wait) when `pdl`
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(Ctx* c, bool pdl, void (*kern)(KArgs...), int grid, int block, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  c->launches++;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define LM_CHAIN(c, pdl, kern, grid, block, ...) LM_CUDA(launch_chain(c, pdl, kern, grid, block, __VA_ARGS__))

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  int x = v ? atoi(v) : 0;
  return x > 0 ? x : dflt;
}

static MapView view_of(const MapIndex& m) {
  MapView v;
  v.sorted = m.sorted;
  v.table = m.table;
  v.l1 = m.l1;
  v.l2_start = m.l2_start;
  v.dev = m.dev;
  return v;
}

int solve_alloc(Ctx* c) {
  size_t cap = (size_t)c->prm.max_points;
  LM_CUDA(cudaMalloc(&c->d_rec, 6 * cap * sizeof(double)));
  LM_CUDA(cudaMalloc(&c->d_recf, 3 * cap * sizeof(float)));
  LM_CUDA(cudaMalloc(&c->d_ok, cap));
  LM_CUDA(cudaMalloc(&c->q_keys, cap * 8));
  LM_CUDA(cudaMalloc(&c->q_keys_alt, cap * 8));
  LM_CUDA(cudaMalloc(&c->q_vals, cap * 4));
  LM_CUDA(cudaMalloc(&c->q_vals_alt, cap * 4));
  LM_CUDA(cudaMalloc(&c->d_pw, cap * sizeof(float4)));
  LM_CUDA(cudaMalloc(&c->d_nbr, 5 * cap * sizeof(int)));
  LM_CUDA(cudaMalloc(&c->d_defer, cap * sizeof(float4)));
  c->partial_blocks = div_up((int)cap, MATCH_BLOCK);
  LM_CUDA(cudaMalloc(&c->d_partial, (size_t)c->partial_blocks * LM_NSUM * sizeof(double)));
  LM_CUDA(cudaMalloc(&c->d_state, sizeof(SolveState)));
  LM_CUDA(cudaMemset(c->d_state, 0, sizeof(SolveState)));
  LM_CUDA(cudaMalloc(&c->d_bar, 64));
  LM_CUDA(cudaMemset(c->d_bar, 0, 64));
  return LMSF_OK;
}

void solve_free(Ctx* c) {
  cudaFree(c->d_rec);
  cudaFree(c->d_recf);
  cudaFree(c->d_ok);
  cudaFree(c->q_keys);
  cudaFree(c->q_keys_alt);
  cudaFree(c->q_vals);
  cudaFree(c->q_vals_alt);
  cudaFree(c->d_pw);
  cudaFree(c->d_nbr);
  cudaFree(c->d_defer);
  cudaFree(c->d_partial);
  cudaFree(c->d_state);
  cudaFree(c->d_bar);
  cudaFree(c->hook_buf);
}

int hook_scratch(Ctx* c, size_t bytes, void** out) {
  if (bytes > c->hook_bytes) {
    size_t want = bytes + bytes / 4 + 4096;
    if (c->hook_buf) LM_CUDA(cudaFree(c->hook_buf));
    c->hook_buf = nullptr;
    c->hook_bytes = 0;
    LM_CUDA(cudaMalloc(&c->hook_buf, want));
    c->hook_bytes = want;
  }
  *out = c->hook_buf;
  return LMSF_OK;
}

int knn_hook(Ctx* c, int kind, const float* d_q, int nq, int* d_idx, float* d_d2, int* d_work) {
  if (!c->map[kind].ready) return LMSF_ERR_STATE;
  if (nq == 0) return LMSF_OK;
  // d_work: [2] counters, [nq] deferral list
  int* ctr = d_work;
  LM_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(int), c->stream));
  LM_LAUNCH(c, k_knn_hook, div_up(nq, KG_BLOCK), KG_BLOCK, 0, view_of(c->map[kind]), d_q, nq, d_idx, d_d2, ctr,
            d_work + 2);
  int sgrid = div_up(nq, 4);
  LM_LAUNCH(c, k_knn_hook_sparse, sgrid < 148 * 2 ? sgrid : 148 * 2, KG_BLOCK, 0, view_of(c->map[kind]), d_q, d_idx, d_d2,
            ctr, d_work + 2);
  LM_CUDA(cudaGetLastError());
  return LMSF_OK;
}

int match_hook(Ctx* c, int kind, const float* d_q, int nq, uint8_t* d_ok, double* d_out10) {
  if (!c->map[kind].ready) return LMSF_ERR_STATE;
  if (nq == 0) return LMSF_OK;
  LM_LAUNCH(c, k_match_hook, div_up(nq, KG_BLOCK), KG_BLOCK, 0, view_of(c->map[kind]), c->map[kind].cat, kind, d_q, nq, d_ok,
            d_out10);
  LM_CUDA(cudaGetLastError());
  return LMSF_OK;
}

// device cloud -> (sum of inlier squared distances, inlier count) read back with one sync
int align_hook(Ctx* c, int kind, const float4* d_pts, int n, const float T12[12], float thresh, double* sum, int* cnt) {
  if (!c->map[kind].ready) return LMSF_ERR_STATE;
  const int nblk = div_up(n, KG_BLOCK);
  // partial sums live behind the caller's points in the hook arena (d_pts is its first n float4)
  double* d_part = (double*)(d_pts + n);
  int* d_cnt = (int*)(d_part + nblk + 1);
  Rigid12f T;
  for (int i = 0; i < 12; ++i) T.m[i] = T12[i];
  LM_LAUNCH(c, k_align_score, nblk, KG_BLOCK, 0, view_of(c->map[kind]), d_pts, n, T, thresh, d_part, d_cnt);
  LM_LAUNCH(c, k_align_finish, 1, 256, 0, d_part, d_cnt, nblk, d_part + nblk, d_cnt + nblk);
  int rc = LMSF_OK;
  if (cudaMemcpyAsync(sum, d_part + nblk, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaMemcpyAsync(cnt, d_cnt + nblk, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess) {
    c->last_error = cudaGetErrorString(cudaGetLastError());
    rc = LMSF_ERR_CUDA;
  }
  return rc;
}

#ifdef LMSF_TIMING
extern "C" int lmsf_debug_kernel_times(unsigned long long out[32], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 32);
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_dbg, z, sizeof z);
  }
  return 0;
}
#endif

// Enqueue one whole solve on the stream (solve_enqueue), then read pose + statistics back with one sync
// (solve_finish).  upper = host-side upper bound of n_edge + n_surf (the device counts are authoritative).
int solve_run(Ctx* c, int solver, double pose[7], lmsf_reg_stats* stats, int upper, int outer_count) {
  LM_TRY(solve_enqueue(c, solver, pose, upper, outer_count));
  return solve_finish(c, solver, pose, stats, outer_count);
}

// lmsf_params.huber_delta is a float; the reference writes the double literal HuberLoss(0.1)
// (ceres_edgeSurfFeatureRegistration.hpp:107) and (double)0.1f = 0.10000000149 moves the pose by 1e-9 m.  The parameter
// is read as the decimal constant it was written as (shortest 6-digit decimal -> double); host code, same as the oracle.
static double decimal_of_float(float f) {
  char buf[48];
  snprintf(buf, sizeof buf, "%.6g", (double)f);
  return strtod(buf, nullptr);
}

int solve_enqueue(Ctx* c, int solver, const double pose[7], int upper, int outer_count) {
  if (upper < 0 || outer_count < 0) return LMSF_ERR_INVALID;
  if (upper > c->prm.max_points) return LMSF_ERR_CAPACITY;
  const auto h_enq0 = std::chrono::steady_clock::now();
  SolveParams sp;
  sp.solver = solver;
  sp.iter = 0;
  sp.lm_max_iters = c->prm.lm_inner_iters;
  sp.gn_min_rows = 10;
  sp.huber = decimal_of_float(c->prm.huber_delta);
  static const int poll_ns = env_int("LMSF_SOLVE_POLL_NS", 0);
  sp.poll_ns = poll_ns;
  MapPair maps;
  maps.edge = view_of(c->map[0]);
  maps.surf = view_of(c->map[1]);
  maps.edge_cat = c->map[0].cat;
  maps.surf_cat = c->map[1].cat;
  const int he = c->map[0].ready ? 1 : 0, hs = c->map[1].ready ? 1 : 0;
  const int up = upper > 0 ? upper : 1;
  // persistent-style grids: a few CTAs per SM, grid-stride over the queries (per-SM factors tunable through the
  // environment for profiling sweeps only)
  static const int fit_per_sm = env_int("LMSF_FIT_GRID", 1), eval_per_sm = env_int("LMSF_EVAL_GRID", 2),
                   knn_per_sm = env_int("LMSF_KNN_GRID", KNN_GRID_PER_SM),
                   seeded_per_sm = env_int("LMSF_KNN_SEEDED_GRID", KNN_SEEDED_MINBLOCKS);
  int fit_grid = div_up(up, MATCH_BLOCK);
  if (fit_grid > 148 * fit_per_sm) fit_grid = 148 * fit_per_sm;
  int eval_grid = div_up(up, MATCH_BLOCK);
  if (eval_grid > 148 * eval_per_sm) eval_grid = 148 * eval_per_sm;
  RecBufs rb;
  rb.d = c->d_rec;
  rb.pl = c->d_recf;
  rb.kind = c->d_ok;
  rb.stride = c->prm.max_points;
  // Processing order of the queries (position -> feature index).  Features from our own extraction come with a
  // permutation that walks the surfs along their rings (consecutive positions are neighbours in space): no sort.
  // Caller-supplied or voxel-filtered features are sorted by map cell once per solve instead.
  static const int force_qsort = env_int("LMSF_FORCE_QSORT", 0);  // tuning experiments
  static const bool want_pdl = env_int("LMSF_NO_PDL", 0) == 0;  // A/B: 1 = ordinary launches
  static const bool split_fit = env_int("LMSF_ONE_FIT", 0) == 0;  // A/B: 1 = k_fit (fit + evaluation at x in one launch)
  // 1 = k_solve (one persistent launch per outer iteration) instead of k_fit + k_lm_eval launches: built for round 2,
  // measured equal in isolation and 12 us per launch slower inside the pipeline (r2p-r2r) — kept selectable, not default
  static const bool fused_solve = env_int("LMSF_FUSED_SOLVE", 0) == 1;
  const bool ring_order = c->perm_valid && !force_qsort;
  // programmatic dependent launches along the chain k_knn -> k_knn_sparse -> k_fit -> k_lm_eval x n -> k_knn ...: only when
  // every launch of the chain is one of ours (the cell-sort path has a library sort in it) and no profiling event is
  // recorded between them
  const bool pdl = want_pdl && ring_order && !c->prof;
  const int* perm = ring_order ? c->d_perm : c->q_vals_alt;
  QueryBufs qb;
  qb.keys = (unsigned*)c->q_keys;
  qb.vals = c->q_vals;
  qb.pw = c->d_pw;
  // query sort keys are relative to a box around the sensor position of the prior pose
  QueryOrigin qorg;
  {
    auto cell = [](double v) { return (int)floor(v < -1.0e5 ? -1.0e5 : (v > 1.0e5 ? 1.0e5 : v)); };
    qorg.x = cell(pose[4]) - 128;
    qorg.y = cell(pose[5]) - 128;
    qorg.z = cell(pose[6]) - 64;
  }
  Pose7 prior;
  for (int i = 0; i < 7; ++i) prior.v[i] = pose[i];
  LM_LAUNCH(c, k_state_init, 1, 32, 0, c->d_state, prior);
  const double alg_bytes = 16.0 * ((double)upper + (double)c->map[0].n_host + (double)c->map[1].n_host) + 216.0;
  for (int it = 0; it < outer_count; ++it) {
    sp.iter = it;
    {
      // the cell sort is a locality heuristic: done once per solve, its permutation is kept for the later
      // outer iterations (the pose moves by millimetres between them), which also keeps the sorted position
      // of a feature stable so that k_knn can seed its search with the previous iteration's neighbours
      StageScope scope(c, LMSF_STAGE_ASSOC);
      if (!ring_order)
        LM_LAUNCH(c, k_assoc, div_up(up, 256), 256, 0, c->d_feat, c->ex.counts, c->d_state, maps, he, hs, qorg, up,
                  solver, it == 0 ? 1 : 0, qb);
      if (it == 0 && !ring_order) {
        size_t tmp = c->cub_tmp_bytes;
        LM_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp, tmp, (unsigned*)c->q_keys, (unsigned*)c->q_keys_alt,
                                                c->q_vals, c->q_vals_alt, up, 0, 32, c->stream));
        c->launches++;
      }
    }
    {
      StageScope scope(c, LMSF_STAGE_MATCH);
      static const int knn_width = env_int("LMSF_KNN_WIDTH", 0);  // tuning experiments
      int chunk = 32;  // queries per warp: halve it while the sweep has fewer chunks than ~1.5x the resident warps
      while (chunk > 8 && up / chunk < 148 * knn_per_sm * (KG_BLOCK / 32) * 3 / 2) chunk >>= 1;
      if (knn_width >= 1 && knn_width <= 32) chunk = knn_width;
      int knn_grid = div_up(up, 4 * chunk);
      if (knn_grid > 148 * knn_per_sm) knn_grid = 148 * knn_per_sm;  // persistent: every resident warp pulls work
      // Deferral pays when the sparse cases are the exception (a raw window: 0.6 % of the queries, but all of the long
      // single-thread chains).  Over a voxel-filtered map (leaf >= 12.5 cm) nearly every query is such a case: a warp
      // per query would be the slow way, and with uniform work there is no tail to remove — each thread searches its own.
      static const int force_defer = env_int("LMSF_KNN_DEFER", 0);  // tuning: 1 = always, 2 = never
      const bool sparse_map = c->prm.map_leaf_surf >= 0.125f;
      const bool defer = force_defer == 1 || (force_defer != 2 && !sparse_map);
      if (defer) {
        static const bool seeded_kernel = env_int("LMSF_NO_SEEDED_KERNEL", 0) == 0;  // A/B: 1 = the general kernel throughout
        if (it > 0 && seeded_kernel) {
          int sgrid = div_up(up, 4 * chunk);
          if (sgrid > 148 * seeded_per_sm) sgrid = 148 * seeded_per_sm;
          LM_CHAIN(c, pdl, (k_knn<true, true>), sgrid, KG_BLOCK, perm, c->d_pw, c->d_feat, c->ex.counts, c->d_state, maps, he,
                   hs, up, solver, 1, ring_order ? 1 : 0, chunk, c->d_defer, c->d_nbr);
        } else {
          LM_CHAIN(c, pdl, k_knn<true>, knn_grid, KG_BLOCK, perm, c->d_pw, c->d_feat, c->ex.counts, c->d_state, maps, he, hs,
                   up, solver, it == 0 ? 0 : 1, ring_order ? 1 : 0, chunk, c->d_defer, c->d_nbr);
        }
        // the sparse cases of this pass, one warp each (nothing to do when k_knn deferred none).  Serving them in the
        // tail of k_knn itself (warps that find the chunk queue empty) was built and measured 3-8x SLOWER, with or without
        // waiting for the producers (r2q, r2s): two large code paths alive on one SM at a time
        static const int sparse_per_sm = env_int("LMSF_SPARSE_GRID", 4);  // 2: 1 559, 3: 1 574, 4: 1 588, 6: 1 575 scans/s (r3j)
        LM_CHAIN(c, pdl, k_knn_sparse, 148 * sparse_per_sm, KG_BLOCK, c->d_state, maps, up, solver, c->d_defer, c->d_nbr);
      } else {
        LM_CHAIN(c, pdl, k_knn<false>, knn_grid, KG_BLOCK, perm, c->d_pw, c->d_feat, c->ex.counts, c->d_state, maps, he, hs,
                 up, solver, it == 0 ? 0 : 1, ring_order ? 1 : 0, chunk, c->d_defer, c->d_nbr);
      }
      c->match_bytes += alg_bytes;
      c->match_launches += 1;
    }
    if (solver == LMSF_SOLVER_HUBER_LM && fused_solve) {
      // fit + the whole trust-region loop of this outer iteration in one persistent cooperative launch
      StageScope scope(c, LMSF_STAGE_FIT);
      const float4* a_feat = c->d_feat;
      const float4* a_pw = c->d_pw;
      const int* a_nbr = c->d_nbr;
      const int* a_counts = c->ex.counts;
      int a_up = up;
      void* args[] = {(void*)&a_feat, (void*)&perm,       (void*)&a_pw, (void*)&a_nbr,        (void*)&a_counts, (void*)&maps,
                      (void*)&a_up,   (void*)&c->d_state, (void*)&rb,   (void*)&c->d_partial, (void*)&sp,       (void*)&c->d_bar};
      static const bool coop = env_int("LMSF_SOLVE_COOP", 0) == 1;
      if (coop) {
        LM_CUDA(cudaLaunchCooperativeKernel((const void*)k_solve, dim3(fit_grid), dim3(MATCH_BLOCK), args, 0, c->stream));
        c->launches++;
      } else {
        LM_LAUNCH(c, k_solve, fit_grid, MATCH_BLOCK, 0, a_feat, perm, a_pw, a_nbr, a_counts, maps, a_up, c->d_state, rb,
                  c->d_partial, sp, c->d_bar);
      }
    } else {
      if (solver == LMSF_SOLVER_HUBER_LM && split_fit) {
        StageScope scope(c, LMSF_STAGE_FIT);
        LM_CHAIN(c, pdl, k_fit_records, div_up(up, MATCH_BLOCK), MATCH_BLOCK, c->d_feat, perm, c->d_pw, c->d_nbr,
                 c->ex.counts, maps, up, rb, sp);
        LM_CHAIN(c, pdl, k_lm_eval0, eval_grid, MATCH_BLOCK, up, c->d_state, rb, c->d_partial, sp);
      } else {
        StageScope scope(c, LMSF_STAGE_FIT);
        LM_CHAIN(c, pdl, k_fit, fit_grid, MATCH_BLOCK, c->d_feat, perm, c->d_pw, c->d_nbr, c->ex.counts, maps, up,
                 c->d_state, rb, c->d_partial, sp);
      }
      if (solver == LMSF_SOLVER_HUBER_LM) {
        StageScope scope(c, LMSF_STAGE_SOLVE);
        for (int k = 0; k < c->prm.lm_inner_iters; ++k)
          LM_CHAIN(c, pdl, k_lm_eval, eval_grid, MATCH_BLOCK, up, c->d_state, rb, c->d_partial, sp,
                   (split_fit || k > 0) ? 1 : 0);
      }
    }
  }
  LM_CUDA(cudaGetLastError());
  c->host_us[2] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h_enq0).count();
  c->host_n[2] += 1;
  LM_CUDA(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(c->h_ints + 32, c->ex.counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  return LMSF_OK;
}

int solve_finish(Ctx* c, int solver, double pose[7], lmsf_reg_stats* stats, int outer_count) {
  LM_CUDA(cudaStreamSynchronize(c->stream));
  c->n_edge = c->h_ints[32];
  c->n_surf = c->h_ints[33];
  for (int k = 0; k < 2; ++k) {  // table occupancy of the last index build; a full table is an error, never silent
    MapIndex& m = c->map[k];
    if (!m.ready) continue;
    m.n_cells_seen = m.h_cnt[2];
    if (m.h_cnt[3]) {
      c->last_error = "local-map hash table overflow";
      m.ready = false;
      return LMSF_ERR_CAPACITY;
    }
  }
  const SolveState& s = *c->h_state;
  for (int i = 0; i < 7; ++i) pose[i] = s.x[i];
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->n_edge_matched = s.n_edge_ok;
    stats->n_surf_matched = s.n_surf_ok;
    stats->final_cost = s.cost;
    if (solver == LMSF_SOLVER_GN) {
      stats->outer_iters = s.gn_iters;
      stats->converged = s.gn_done;
      stats->degenerate = s.gn_degenerate;
    } else {
      stats->outer_iters = outer_count;
      stats->lm_steps_total = s.lm_steps_total;
      stats->lm_steps_accepted = s.lm_steps_accepted;
    }
  }
  return LMSF_OK;
}

}  // namespace lm

#ifdef LMSF_KNN_CHECK
// debugging builds (make EXTRA=-DLMSF_KNN_CHECK): the first index violation a search launched from match.cu met
namespace lm {
__global__ void k_knn_check_collect(int* __restrict__ out) {
  if (threadIdx.x < 8) out[threadIdx.x] = g_knn_check[threadIdx.x];
}
}  // namespace lm
extern "C" int lmsf_debug_knn_check(lmsf_ctx* c, int out[8]) {
  int* d = nullptr;
  LM_TRY(hook_scratch(c, 64, (void**)&d));
  lm::k_knn_check_collect<<<1, 32, 0, c->stream>>>(d);
  LM_CUDA(cudaMemcpyAsync(out, d, 32, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  return LMSF_OK;
}
#endif
