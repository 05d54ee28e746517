// api.cu — the C ABI (include/lmsf_b200.h): context lifetime, the three seams, and the
// scan-to-map tracker (LidarTracker/LidarTrackerLocalMap.hpp:107-262) whose point data never
// leaves the device: one H2D copy per sweep in, 7+ doubles out.
#include <math.h>
#include <string.h>

#include <chrono>
#include <new>
#include <utility>

#include "common.cuh"

namespace lm {

// ------------------------------------------------------------------ profiling
StageScope::StageScope(Ctx* c_, int s, cudaStream_t stream) : c(c_), stage(s), st(stream ? stream : c_->stream), l0(c_->launches) {
  if (!c->prof) return;
  auto take = [&]() {
    cudaEvent_t e = nullptr;
    if (!c->ev_pool.empty()) {
      e = c->ev_pool.back();
      c->ev_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  };
  a = take();
  b = take();
  cudaEventRecord(a, st);
}
StageScope::~StageScope() {
  c->prof_launch[stage] += c->launches - l0;
  if (!a) return;
  cudaEventRecord(b, st);
  c->spans.push_back(Ctx::Span{stage, a, b});
}

// pcl::transformPointCloud(cloud, out, Matrix4d) (LidarTrackerLocalMap.hpp:217): double math, float store
struct Rigid12 {
  double R[9], t[3];
};
__global__ void __launch_bounds__(256) k_transform(const float4* __restrict__ in, const int* __restrict__ counts,
                                                   int kind, Rigid12 T, float4* __restrict__ out) {
  const int n_e = counts[0];
  const int n = kind ? counts[1] : n_e;
  const float4* src = in + (kind ? n_e : 0);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = src[i];
  double x = p.x, y = p.y, z = p.z;
  float4 o;
  o.x = (float)(T.R[0] * x + T.R[1] * y + T.R[2] * z + T.t[0]);
  o.y = (float)(T.R[3] * x + T.R[4] * y + T.R[5] * z + T.t[1]);
  o.z = (float)(T.R[6] * x + T.R[7] * y + T.R[8] * z + T.t[2]);
  o.w = p.w;
  out[i] = o;
}

static int set_counts(Ctx* c, int n_e, int n_s) {
  c->h_ints[40] = n_e;
  c->h_ints[41] = n_s;
  LM_CUDA(cudaMemcpyAsync(c->ex.counts, c->h_ints + 40, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  c->n_edge = n_e;
  c->n_surf = n_s;
  return LMSF_OK;
}

static int fetch_counts(Ctx* c) {
  LM_CUDA(cudaMemcpyAsync(c->h_ints + 32, c->ex.counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  c->n_edge = c->h_ints[32];
  c->n_surf = c->h_ints[33];
  return LMSF_OK;
}

// choose the cloud the index is built over (raw window or its voxel-filtered copy) and build
static int index_window(Ctx* c, int kind, cudaStream_t st, bool fixed_grid) {
  MapIndex& m = c->map[kind];
  int total = 0;
  for (int f : m.frame_n) total += f;
  float leaf = kind ? c->prm.map_leaf_surf : c->prm.map_leaf_edge;
  int n = total;
  if (leaf > 0.f && total > 0) {
    LM_TRY(voxel_run(c, m.win, total, leaf, m.vox, &n, nullptr));  // main stream, synchronising
    m.cat = m.vox;
  } else {
    m.cat = m.win;
  }
  if (n == 0) return LMSF_OK;  // an empty source is ignored (ceres_edgeSurfFeatureRegistration.hpp:58)
  // density hint for the search: over a voxel-filtered map the points are about a leaf apart, five of them lie within
  // 1.25-1.6 leaves — the box grows to that width at once instead of cell by cell
  static const int grow_pct = [] {
    const char* v = getenv("LMSF_GROW_PCT");  // tuning
    int x = v ? atoi(v) : 0;
    return x > 0 ? x : 160;  // voxel-map bench: 60 % 1 021, 100 % 1 066, 125 % 1 068, 160 % 1 092 scans/s (r3m)
  }();
  int g = leaf > 0.f ? (int)ceilf(leaf * 16.0f * (float)grow_pct / 100.0f) : 2;
  m.grow0 = g < 2 ? 2 : (g > 12 ? 12 : g);
  return map_build(c, m, n, st, fixed_grid);
}

// updateLocalMap (LidarTrackerLocalMap.hpp:205-232) over the sliding-window local map (the reference's
// PointCloudLocalMapBase is missing from its tree; contract inferred, SURVEY.md §8 a6'):
// type 1 = AddFrameForMotion (append, evict the oldest beyond `window`), 2 = AddFrameForTime
// (replace the newest frame).  c->n_edge / c->n_surf must be valid on the host and the main stream idle.
//
// When the features came from our own extraction (finite, range-gated) and no map voxel filter is set, the
// whole update runs asynchronously on the map stream inside a grid frozen around the sensor: no host sync,
// and the index rebuild overlaps the next sweep's upload and feature extraction.  ev_feat_free releases
// d_feat to the next extraction, ev_map_done gates the next solve.
static int update_map(Ctx* c, const rigid& T, int type) {
  Rigid12 T12;
  for (int i = 0; i < 9; ++i) T12.R[i] = T.R[i];
  for (int i = 0; i < 3; ++i) T12.t[i] = T.t[i];
  const bool async = c->feat_from_extract && !(c->prm.map_leaf_edge > 0.f) && !(c->prm.map_leaf_surf > 0.f);
  cudaStream_t st = async ? c->stream_map : c->stream;
  const double reach = 1.2 * (double)c->prm.max_range + 2.0;
  bool fixed_ok[2] = {false, false};
  for (int kind = 0; kind < 2; ++kind) {  // capacity of both maps first: a refused update leaves both windows untouched
    int nk = kind ? c->n_surf : c->n_edge;
    if (nk == 0) continue;
    const MapIndex& m = c->map[kind];
    int total = 0;
    for (int f : m.frame_n) total += f;
    if (type == 1) {
      if ((int)m.frame_n.size() + 1 > c->prm.window) total -= m.frame_n.front();
    } else if (!m.frame_n.empty()) {
      total -= m.frame_n.back();
    }
    if (total + nk > m.cap) return LMSF_ERR_CAPACITY;
  }
  for (int kind = 0; kind < 2; ++kind) {
    int nk = kind ? c->n_surf : c->n_edge;
    if (nk == 0) continue;
    MapIndex& m = c->map[kind];
    StageScope scope(c, LMSF_STAGE_MAP, st);
    int total = 0;
    for (int f : m.frame_n) total += f;
    int drop = 0;
    if (type == 1) {
      if ((int)m.frame_n.size() + 1 > c->prm.window) drop = m.frame_n.front();
    } else if (!m.frame_n.empty()) {
      total -= m.frame_n.back();
      m.frame_n.pop_back();
      m.frame_pos.pop_back();
    }
    int kept = total - drop;
    if (kept + nk > m.cap) return LMSF_ERR_CAPACITY;
    if (drop > 0) {
      LM_CUDA(cudaMemcpyAsync(m.win_alt, m.win + drop, (size_t)kept * sizeof(float4), cudaMemcpyDeviceToDevice, st));
      float4* t = m.win;
      m.win = m.win_alt;
      m.win_alt = t;
      m.frame_n.erase(m.frame_n.begin());
      m.frame_pos.erase(m.frame_pos.begin());
    }
    LM_LAUNCH_ON(c, st, k_transform, div_up(nk, 256), 256, 0, c->d_feat, c->ex.counts, kind, T12, m.win + kept);
    m.frame_n.push_back(nk);
    m.frame_pos.push_back({T.t[0], T.t[1], T.t[2]});
    if (async) {
      // keep the frozen grid while every frame of the window fits it; otherwise re-centre it on this pose
      bool ok = m.fixed;
      for (const auto& fp : m.frame_pos) ok = ok && map_grid_covers(m, fp.data(), reach);
      if (!ok) {
        ok = map_freeze_grid(m, T.t);
        for (const auto& fp : m.frame_pos) ok = ok && map_grid_covers(m, fp.data(), reach);
      }
      fixed_ok[kind] = ok;
    }
  }
  if (async) {  // the map stream has read this slot's features: the front end may refill it
    Ctx::FeatSlot& sl = c->slot[c->slot_cur];
    LM_CUDA(cudaEventRecord(sl.freed, st));
    sl.freed_pending = true;
  }
  // the two index builds are independent: in the asynchronous case the (small) edge build runs on a second stream
  // beside the surf build instead of in front of it
  const bool fork = async && c->n_edge > 0 && c->n_surf > 0;
  if (fork) {
    LM_CUDA(cudaEventRecord(c->ev_map_fork, st));
    LM_CUDA(cudaStreamWaitEvent(c->stream_map2, c->ev_map_fork, 0));
  }
  for (int kind = 0; kind < 2; ++kind) {
    int nk = kind ? c->n_surf : c->n_edge;
    if (nk == 0) continue;
    LM_TRY(index_window(c, kind, (fork && kind == 0) ? c->stream_map2 : st, async && fixed_ok[kind]));
  }
  if (fork) {
    LM_CUDA(cudaEventRecord(c->ev_map_join, c->stream_map2));
    LM_CUDA(cudaStreamWaitEvent(st, c->ev_map_join, 0));
  }
  if (async) {
    LM_CUDA(cudaEventRecord(c->ev_map_done, st));
    c->map_pending = true;
  }
  return LMSF_OK;
}

// needUpdataLocalMap (LidarTrackerLocalMap.hpp:239-262)
static int need_update(Ctx* c, const rigid& curr, double stamp) {
  if (stamp - c->last_kf_time > c->prm.kf_time) return 2;
  rigid d = rigid_mul(rigid_inv(c->last_kf), curr);
  double dt = sqrt(d.t[0] * d.t[0] + d.t[1] * d.t[1] + d.t[2] * d.t[2]);
  quat q = mat_to_quat(d.R);
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  double da = acos(q.w / n) * 2;
  if (dt > c->prm.kf_trans || da > c->prm.kf_rot) return 1;
  return 0;
}

static int next_lm_outer(Ctx* c) {
  if (c->lm_count > 2) c->lm_count--;  // ceres_edgeSurfFeatureRegistration.hpp:100-101
  return c->lm_count;
}

static int run_solver(Ctx* c, int solver, double pose[7], lmsf_reg_stats* st, int upper) {
  int outer = (solver == LMSF_SOLVER_GN) ? c->prm.gn_max_iters : next_lm_outer(c);
  return solve_run(c, solver, pose, st, upper, outer);
}

// optional voxel filter of the scan features (off by default); needs the counts on the host
static int scan_filter(Ctx* c) {
  const float le = c->prm.scan_leaf_edge, ls = c->prm.scan_leaf_surf;
  if (!(le > 0.f) && !(ls > 0.f)) return LMSF_OK;
  LM_TRY(fetch_counts(c));
  c->perm_valid = false;  // the features are about to be replaced by voxel centroids
  const int ne = c->n_edge, ns = c->n_surf;
  if (ne + ns == 0) return LMSF_OK;
  LM_CUDA(cudaMemcpyAsync(c->d_tmp, c->d_feat, (size_t)(ne + ns) * sizeof(float4), cudaMemcpyDeviceToDevice,
                          c->stream));
  int ne2 = ne, ns2 = ns;
  if (le > 0.f) LM_TRY(voxel_run(c, c->d_tmp, ne, le, c->d_feat, &ne2, nullptr));
  if (ls > 0.f) {
    LM_TRY(voxel_run(c, c->d_tmp + ne, ns, ls, c->d_feat + ne2, &ns2, nullptr));
  } else if (ne2 != ne) {
    LM_CUDA(cudaMemcpyAsync(c->d_feat + ne2, c->d_tmp + ne, (size_t)ns * sizeof(float4), cudaMemcpyDeviceToDevice,
                            c->stream));
  }
  return set_counts(c, ne2, ns2);
}

// LidarTrackerLocalMap::Solve (:107-160) in two halves; features already in d_feat, counts on the device.
// tracker_begin: prediction + the whole registration enqueued on the main stream, no host wait.
static int tracker_begin(Ctx* c, int upper, double stamp, const double delta[7]) {
  if (c->pending.active) return LMSF_ERR_STATE;  // one sweep in flight per tracker
  LM_TRY(scan_filter(c));
  c->pending = Ctx::Pending();
  c->pending.stamp = stamp;
  if (!c->init) {
    c->pending.first = true;
    c->pending.active = true;
    return LMSF_OK;
  }
  bool ident = delta[0] == 0 && delta[1] == 0 && delta[2] == 0 && delta[3] == 1 && delta[4] == 0 && delta[5] == 0 &&
               delta[6] == 0;
  c->curr = ident ? rigid_mul(c->prev, c->motion) : rigid_mul(c->prev, rigid_from_pose(delta));
  double p[7];
  rigid_to_pose(c->curr, p);  // Quaterniond(T.rotation())
  LM_TRY(wait_map(c));  // the previous keyframe's index rebuild (map stream) must be complete
  c->pending.solver = c->prm.solver;
  c->pending.outer = (c->prm.solver == LMSF_SOLVER_GN) ? c->prm.gn_max_iters : next_lm_outer(c);
  LM_TRY(solve_enqueue(c, c->pending.solver, p, upper, c->pending.outer));
  c->pending.active = true;
  return LMSF_OK;
}

// tracker_end: wait for the pose, motion update, keyframe test (:239-262), local-map update (:205-232)
static int tracker_end(Ctx* c, double delta[7], double pose_out[7], lmsf_track_stats* out) {
  if (!c->pending.active) return LMSF_ERR_STATE;
  c->pending.active = false;
  const double stamp = c->pending.stamp;
  lmsf_track_stats s;
  memset(&s, 0, sizeof s);
  if (c->pending.first) {
    LM_TRY(fetch_counts(c));
    LM_TRY(update_map(c, rigid_identity(), 1));
    c->curr = c->prev = c->motion = c->last_kf = rigid_identity();
    c->last_kf_time = stamp;
    c->init = true;
    s.first = 1;
    s.keyframe = 1;
  } else {
    double p[7];
    const auto h0 = std::chrono::steady_clock::now();
    LM_TRY(solve_finish(c, c->pending.solver, p, &s.reg, c->pending.outer));  // leaves n_edge / n_surf on the host
    c->host_us[0] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count();
    c->host_n[0] += 1;
    c->curr = rigid_from_pose(p);  // T.linear() = q.toRotationMatrix()
    c->motion = rigid_mul(rigid_inv(c->prev), c->curr);
    rigid_to_pose(c->motion, delta);
    c->prev = c->curr;
    int ut = need_update(c, c->curr, stamp);
    s.keyframe = ut;
    if (ut) {
      c->last_kf = c->curr;
      c->last_kf_time = stamp;
      const auto h1 = std::chrono::steady_clock::now();
      LM_TRY(update_map(c, c->curr, ut));
      c->host_us[1] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h1).count();
      c->host_n[1] += 1;
    }
  }
  s.n_edge = c->n_edge;
  s.n_surf = c->n_surf;
  s.map_edge = c->map[0].n_host;
  s.map_surf = c->map[1].n_host;
  rigid_to_pose(c->curr, pose_out);
  if (out) *out = s;
  return LMSF_OK;
}

static int tracker_core(Ctx* c, int upper, double stamp, double delta[7], double pose_out[7], lmsf_track_stats* out) {
  LM_TRY(tracker_begin(c, upper, stamp, delta));
  return tracker_end(c, delta, pose_out, out);
}

// make slot `si` the one the main stream works on (d_feat / ex.counts alias it)
static void alias_slot(Ctx* c, int si) {
  c->slot_cur = si;
  c->d_feat = c->slot[si].feat;
  c->ex.counts = c->slot[si].counts;
  c->d_perm = c->slot[si].perm;
  c->perm_valid = false;
}

// a slot that holds no unconsumed prefetch (the current one preferred); when both do, the older prefetch is dropped
static int writable_slot(Ctx* c) {
  if (!c->slot[c->slot_cur].filled) return c->slot_cur;
  if (!c->slot[c->slot_cur ^ 1].filled) return c->slot_cur ^ 1;
  int si = c->slot[0].seq < c->slot[1].seq ? 0 : 1;
  c->slot[si].filled = false;
  return si;
}

static int upload_features(Ctx* c, const float* edge, int n_e, const float* surf, int n_s) {
  if (n_e < 0 || n_s < 0 || (n_e > 0 && !edge) || (n_s > 0 && !surf)) return LMSF_ERR_INVALID;
  if (n_e + n_s > c->prm.max_points) return LMSF_ERR_CAPACITY;
  if (c->pending.active) return LMSF_ERR_STATE;
  alias_slot(c, writable_slot(c));
  LM_TRY(wait_feat(c));
  c->feat_from_extract = false;  // caller-supplied features: arbitrary coordinates, the map update stays synchronous
  if (n_e) memcpy(c->h_pts, edge, (size_t)n_e * sizeof(float4));
  if (n_s) memcpy(c->h_pts + n_e, surf, (size_t)n_s * sizeof(float4));
  if (n_e + n_s)
    LM_CUDA(cudaMemcpyAsync(c->d_feat, c->h_pts, (size_t)(n_e + n_s) * sizeof(float4), cudaMemcpyHostToDevice,
                            c->stream));
  return set_counts(c, n_e, n_s);
}

// host sweep -> d_sweep on the front-end stream, through one of two pinned staging buffers
static int upload_sweep(Ctx* c, const float* xyzi, int n) {
  if (n < 0 || (n > 0 && !xyzi)) return LMSF_ERR_INVALID;
  if (n > c->prm.max_points) return LMSF_ERR_CAPACITY;
  if (n) {
    const int b = c->stage_next;
    c->stage_next ^= 1;
    LM_CUDA(cudaEventSynchronize(c->ev_stage[b]));  // the copy that last read this staging buffer is done
    memcpy(c->h_stage[b], xyzi, (size_t)n * sizeof(float4));
    LM_CUDA(cudaMemcpyAsync(c->d_sweep, c->h_stage[b], (size_t)n * sizeof(float4), cudaMemcpyHostToDevice,
                            c->stream_fe));
    LM_CUDA(cudaEventRecord(c->ev_stage[b], c->stream_fe));
  }
  return LMSF_OK;
}

// Feature extraction of one sweep into slot `si` on the front-end stream (a1).  host != 0: xyzi is a host buffer,
// staged and copied first; otherwise it is a device pointer.  Returns with the work enqueued.
static int extract_into_slot(Ctx* c, int si, const float* xyzi, int n, bool host) {
  if (n < 0 || (n > 0 && !xyzi)) return LMSF_ERR_INVALID;
  if (n > c->prm.max_points) return LMSF_ERR_CAPACITY;
  Ctx::FeatSlot& sl = c->slot[si];
  if (sl.freed_pending) {  // a local-map update may still be reading this slot
    LM_CUDA(cudaStreamWaitEvent(c->stream_fe, sl.freed, 0));
    sl.freed_pending = false;
  }
  const float4* d_in = (const float4*)xyzi;
  if (host) {
    LM_TRY(upload_sweep(c, xyzi, n));
    d_in = c->d_sweep;
  }
  LM_TRY(extract_run(c, d_in, n, c->stream_fe, sl.feat, sl.counts, sl.perm));
  LM_CUDA(cudaEventRecord(sl.ready, c->stream_fe));
  return LMSF_OK;
}

// adopt a slot whose extraction has been enqueued on the front-end stream
static int adopt_slot(Ctx* c, int si) {
  Ctx::FeatSlot& sl = c->slot[si];
  alias_slot(c, si);
  sl.filled = false;
  LM_CUDA(cudaStreamWaitEvent(c->stream, sl.ready, 0));
  c->feat_from_extract = true;
  c->perm_valid = true;
  return LMSF_OK;
}

// extraction for the synchronous entry points: main stream ordered behind it
static int extract_current(Ctx* c, const float* xyzi, int n, bool host) {
  if (c->pending.active) return LMSF_ERR_STATE;  // a submitted sweep is still reading its feature slot
  const int si = writable_slot(c);
  LM_TRY(extract_into_slot(c, si, xyzi, n, host));
  return adopt_slot(c, si);
}

// lmsf_tracker_prefetch(_dev): extraction of a coming sweep into a slot that is not waiting to be consumed.
// Called as prefetch(k+1) after submit(k): slot A feeds the solve in flight, k+1 goes to slot B.  The prefetched
// sweep is identified by the ticket handed back here and by nothing else — a caller buffer recycled at the same
// address can never be mistaken for it.
static int prefetch(Ctx* c, const float* xyzi, int n, bool host, int64_t* ticket) {
  if (!ticket) return LMSF_ERR_INVALID;
  *ticket = 0;
  int si = c->slot_cur ^ 1;
  if (c->slot[si].filled) si = c->slot_cur;
  if (c->slot[si].filled) return LMSF_ERR_STATE;  // two sweeps already waiting for their tracker step
  if (c->pending.active && si == c->slot_cur) return LMSF_ERR_STATE;  // that slot feeds the solve in flight
  LM_TRY(extract_into_slot(c, si, xyzi, n, host));
  c->slot[si].filled = true;
  c->slot[si].n = n;
  c->slot[si].seq = ++c->prefetch_seq;
  *ticket = c->slot[si].seq;
  return LMSF_OK;
}

// tracker entry by ticket: the prefetched sweep becomes the current one; *n_out = its point count
static int take_ticket(Ctx* c, int64_t ticket, int* n_out) {
  if (ticket <= 0) return LMSF_ERR_INVALID;
  for (int si = 0; si < 2; ++si) {
    const Ctx::FeatSlot& sl = c->slot[si];
    if (sl.filled && sl.seq == ticket) {
      *n_out = sl.n;
      return adopt_slot(c, si);
    }
  }
  return LMSF_ERR_STATE;  // unknown, cancelled, displaced or already consumed
}

// tracker entry with a sweep: always extracted now (a waiting prefetch is only displaced when both slots hold one)
static int take_sweep(Ctx* c, const float* xyzi, int n, bool host) { return extract_current(c, xyzi, n, host); }

static void destroy(Ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream_fe) cudaStreamSynchronize(c->stream_fe);
  if (c->stream_map2) cudaStreamSynchronize(c->stream_map2);
  if (c->stream_map) cudaStreamSynchronize(c->stream_map);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& sp : c->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  scdb_free(c);
  extract_free(c);
  voxel_free(c);
  solve_free(c);
  map_free(c->map[0]);
  map_free(c->map[1]);
  cudaFree(c->d_sweep);
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->slot[i].feat);
    cudaFree(c->slot[i].counts);
    cudaFree(c->slot[i].perm);
    if (c->slot[i].ready) cudaEventDestroy(c->slot[i].ready);
    if (c->slot[i].freed) cudaEventDestroy(c->slot[i].freed);
    cudaFreeHost(c->h_stage[i]);
    if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]);
  }
  cudaFree(c->d_tmp);
  cudaFreeHost(c->h_pts);
  cudaFreeHost(c->h_pose);
  cudaFreeHost(c->h_ints);
  cudaFreeHost(c->h_state);
  if (c->ev_map_done) cudaEventDestroy(c->ev_map_done);
  if (c->ev_map_fork) cudaEventDestroy(c->ev_map_fork);
  if (c->ev_map_join) cudaEventDestroy(c->ev_map_join);
  if (c->stream_map2) cudaStreamDestroy(c->stream_map2);
  if (c->stream_map) cudaStreamDestroy(c->stream_map);
  if (c->stream_fe) cudaStreamDestroy(c->stream_fe);
  if (c->stream) cudaStreamDestroy(c->stream);
}

static int create(Ctx* c) {
  size_t cap = (size_t)c->prm.max_points;
  // The registration and the local-map update are the critical path of a sweep; the front end (extraction of the NEXT
  // sweep) only has to be ready when they are done.  Stream priorities make the block scheduler hand freed SM resources to
  // the critical path first, so the extraction fills its gaps instead of delaying its kernels (measured neutral on the
  // bench sequence, r2r: the extraction already fitted into the tails of k_knn / k_fit).
  int prio_lo = 0, prio_hi = 0;
  LM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // lo = numerically greatest = least priority
  LM_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
  LM_CUDA(cudaStreamCreateWithPriority(&c->stream_map, cudaStreamNonBlocking, prio_hi));
  LM_CUDA(cudaStreamCreateWithPriority(&c->stream_map2, cudaStreamNonBlocking, prio_hi));
  LM_CUDA(cudaEventCreateWithFlags(&c->ev_map_fork, cudaEventDisableTiming));
  LM_CUDA(cudaEventCreateWithFlags(&c->ev_map_join, cudaEventDisableTiming));
  LM_CUDA(cudaStreamCreateWithPriority(&c->stream_fe, cudaStreamNonBlocking, prio_lo));
  LM_CUDA(cudaEventCreateWithFlags(&c->ev_map_done, cudaEventDisableTiming));
  LM_CUDA(cudaMalloc(&c->d_sweep, cap * sizeof(float4)));
  for (int i = 0; i < 2; ++i) {
    LM_CUDA(cudaMalloc(&c->slot[i].feat, cap * sizeof(float4)));
    LM_CUDA(cudaMalloc(&c->slot[i].counts, 4 * sizeof(int)));
    LM_CUDA(cudaMemset(c->slot[i].counts, 0, 4 * sizeof(int)));
    LM_CUDA(cudaMalloc(&c->slot[i].perm, cap * sizeof(int)));
    LM_CUDA(cudaEventCreateWithFlags(&c->slot[i].ready, cudaEventDisableTiming));
    LM_CUDA(cudaEventCreateWithFlags(&c->slot[i].freed, cudaEventDisableTiming));
    LM_CUDA(cudaMallocHost(&c->h_stage[i], cap * sizeof(float4)));
    LM_CUDA(cudaEventCreateWithFlags(&c->ev_stage[i], cudaEventDisableTiming));
  }
  c->slot_cur = 0;
  c->d_feat = c->slot[0].feat;
  c->d_perm = c->slot[0].perm;
  LM_CUDA(cudaMalloc(&c->d_tmp, cap * sizeof(float4)));
  LM_CUDA(cudaMallocHost(&c->h_pts, cap * sizeof(float4)));
  LM_CUDA(cudaMallocHost(&c->h_pose, 16 * sizeof(double)));
  LM_CUDA(cudaMallocHost(&c->h_ints, 64 * sizeof(int)));
  LM_CUDA(cudaMallocHost(&c->h_state, sizeof(SolveState)));
  LM_TRY(extract_alloc(c));
  c->ex.counts = c->slot[0].counts;
  LM_TRY(voxel_alloc(c));
  LM_TRY(solve_alloc(c));
  LM_TRY(map_alloc(c, c->map[0], c->prm.max_map_points));
  LM_TRY(map_alloc(c, c->map[1], c->prm.max_map_points));
  c->curr = c->prev = c->motion = c->last_kf = rigid_identity();
  c->lm_count = c->prm.lm_outer_start;
  return LMSF_OK;
}

}  // namespace lm

using namespace lm;

#define ENTER(c)                      \
  if (!(c)) return LMSF_ERR_INVALID;  \
  if (cudaSetDevice((c)->device) != cudaSuccess) return LMSF_ERR_NO_DEVICE

extern "C" {

int lmsf_params_default(lmsf_params* p) {
  if (!p) return LMSF_ERR_INVALID;
  memset(p, 0, sizeof *p);
  p->n_scans = 16;
  p->min_range = 2.f;
  p->max_range = 80.f;
  p->edge_thresh = 1.f;
  p->remove_bad_points = 1;
  p->max_points = 262144;
  p->window = 10;
  p->solver = LMSF_SOLVER_HUBER_LM;
  p->gn_max_iters = 10;
  p->lm_outer_start = 10;
  p->lm_inner_iters = 4;
  p->huber_delta = 0.1f;
  p->kf_trans = 0.3;
  p->kf_rot = 0.1;
  p->kf_time = 10.0;
  p->max_map_points = 0;
  return LMSF_OK;
}

int lmsf_ctx_create(int device, const lmsf_params* p, lmsf_ctx** out) {
  if (!p || !out) return LMSF_ERR_INVALID;
  *out = nullptr;
  if (p->n_scans != 16 && p->n_scans != 32 && p->n_scans != 64) return LMSF_ERR_INVALID;
  if (p->max_points <= 0 || p->window <= 0 || p->lm_inner_iters < 0 || p->gn_max_iters < 0) return LMSF_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LMSF_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return LMSF_ERR_NO_DEVICE;
  lmsf_ctx* c = new (std::nothrow) lmsf_ctx();
  if (!c) return LMSF_ERR_INVALID;
  c->device = device;
  c->prm = *p;
  if (c->prm.max_map_points <= 0) {
    long long m = (long long)c->prm.window * c->prm.max_points;
    c->prm.max_map_points = (int)(m > 16777216LL ? 16777216LL : m);
  }
  int rc = create(c);
  if (rc != LMSF_OK) {
    static thread_local std::string keep;
    keep = c->last_error;
    fprintf(stderr, "lmsf_ctx_create: %s\n", keep.c_str());
    destroy(c);
    delete c;
    return rc;
  }
  *out = c;
  return LMSF_OK;
}

void lmsf_ctx_destroy(lmsf_ctx* c) {
  if (!c) return;
  destroy(c);
  delete c;
}

const char* lmsf_strerror(int code) {
  switch (code) {
    case LMSF_OK: return "ok";
    case LMSF_ERR_INVALID: return "invalid argument";
    case LMSF_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU path)";
    case LMSF_ERR_CUDA: return "CUDA error";
    case LMSF_ERR_CAPACITY: return "input exceeds the capacity the context was created with";
    case LMSF_ERR_STATE: return "call out of order (no local map set)";
    default: return "unknown error";
  }
}

const char* lmsf_last_cuda_error(lmsf_ctx* c) { return c ? c->last_error.c_str() : ""; }
int64_t lmsf_launch_count(lmsf_ctx* c) { return c ? c->launches : 0; }
void* lmsf_stream(lmsf_ctx* c) { return c ? (void*)c->stream : nullptr; }

int lmsf_extract_features(lmsf_ctx* c, const float* xyzi, int n, uint8_t* label_out, float* edge_xyzi, int* n_edge,
                          float* surf_xyzi, int* n_surf) {
  ENTER(c);
  LM_TRY(extract_current(c, xyzi, n, true));
  LM_TRY(fetch_counts(c));
  const int ne = c->n_edge, ns = c->n_surf;
  if (label_out && n) LM_CUDA(cudaMemcpyAsync(label_out, c->ex.label, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (edge_xyzi && ne)
    LM_CUDA(cudaMemcpyAsync(edge_xyzi, c->d_feat, (size_t)ne * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (surf_xyzi && ns)
    LM_CUDA(cudaMemcpyAsync(surf_xyzi, c->d_feat + ne, (size_t)ns * sizeof(float4), cudaMemcpyDeviceToHost,
                            c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  if (n_edge) *n_edge = ne;
  if (n_surf) *n_surf = ns;
  return LMSF_OK;
}

int lmsf_extract_features_to_dev(lmsf_ctx* c, const float* xyzi, int n, void* d_feat_out, int cap, int* n_edge,
                                 int* n_surf) {
  ENTER(c);
  if (!n_edge || !n_surf || (cap > 0 && !d_feat_out)) return LMSF_ERR_INVALID;
  LM_TRY(extract_current(c, xyzi, n, true));
  LM_TRY(fetch_counts(c));
  const int ne = c->n_edge, ns = c->n_surf;
  if (ne + ns > cap) return LMSF_ERR_CAPACITY;
  if (ne + ns)
    LM_CUDA(cudaMemcpyAsync(d_feat_out, c->d_feat, (size_t)(ne + ns) * sizeof(float4), cudaMemcpyDeviceToDevice,
                            c->stream));
  *n_edge = ne;
  *n_surf = ns;
  return LMSF_OK;
}

int lmsf_voxel_downsample(lmsf_ctx* c, const float* xyzi, int n, float leaf, float* out_xyzi, int* n_out,
                          int32_t* voxel_of_point) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (n < 0 || !(leaf > 0.f) || !n_out || (n > 0 && !xyzi)) return LMSF_ERR_INVALID;
  if (n > c->vox_cap) return LMSF_ERR_CAPACITY;
  *n_out = 0;
  if (n == 0) return LMSF_OK;
  // staged through pageable->device copies in chunks of the pinned buffer
  const int chunk = c->prm.max_points;
  for (int o = 0; o < n; o += chunk) {
    int m = (n - o < chunk) ? n - o : chunk;
    memcpy(c->h_pts, xyzi + 4 * (size_t)o, (size_t)m * sizeof(float4));
    LM_CUDA(cudaMemcpyAsync(c->v_in + o, c->h_pts, (size_t)m * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    LM_CUDA(cudaStreamSynchronize(c->stream));
  }
  int nv = 0;
  LM_TRY(voxel_run(c, c->v_in, n, leaf, c->v_out, &nv, voxel_of_point ? c->v_member : nullptr));
  if (out_xyzi && nv)
    LM_CUDA(cudaMemcpyAsync(out_xyzi, c->v_out, (size_t)nv * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (voxel_of_point)
    LM_CUDA(cudaMemcpyAsync(voxel_of_point, c->v_member, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  *n_out = nv;
  return LMSF_OK;
}

int lmsf_rotary_preprocess(lmsf_ctx* c, const float* xyzi, int n, float scan_period, float* out_xyzi, int* n_out) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (n < 0 || !n_out || !(scan_period > 0.f) || (n > 0 && (!xyzi || !out_xyzi))) return LMSF_ERR_INVALID;
  if (n > c->prm.max_points) return LMSF_ERR_CAPACITY;
  *n_out = 0;
  if (n == 0) return LMSF_OK;
  LM_CUDA(cudaMemcpyAsync(c->d_tmp, xyzi, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  LM_TRY(rotary_apply(c, c->d_tmp, n, scan_period, c->stream));
  int m = 0;
  LM_TRY(filter_run(c, c->d_tmp, n, 0, 0.f, 0.f, c->v_out, &m));  // removeNaN: order-preserving
  if (m > 0) LM_CUDA(cudaMemcpyAsync(out_xyzi, c->v_out, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  *n_out = m;
  return LMSF_OK;
}

int lmsf_common_process(lmsf_ctx* c, const float* xyzi, int n, int remove_nan, float leaf, float dist_near,
                        float dist_far, float* out_xyzi, int* n_out) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (n < 0 || !n_out || (n > 0 && (!xyzi || !out_xyzi)) || leaf < 0.f) return LMSF_ERR_INVALID;
  if (n > c->vox_cap) return LMSF_ERR_CAPACITY;
  *n_out = 0;
  if (n == 0) return LMSF_OK;
  const int chunk = c->prm.max_points;
  for (int o = 0; o < n; o += chunk) {
    int m = (n - o < chunk) ? n - o : chunk;
    memcpy(c->h_pts, xyzi + 4 * (size_t)o, (size_t)m * sizeof(float4));
    LM_CUDA(cudaMemcpyAsync(c->v_in + o, c->h_pts, (size_t)m * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    LM_CUDA(cudaStreamSynchronize(c->stream));
  }
  float4 *cur = c->v_in, *other = c->v_out;
  int m = n;
  if (remove_nan) {  // common_processing.hpp:93-97
    int k = 0;
    LM_TRY(filter_run(c, cur, m, 0, 0.f, 0.f, other, &k));
    std::swap(cur, other);
    m = k;
  }
  if (leaf > 0.f && m > 0) {  // :103 (a VoxelGridFilter without a leaf returns its input)
    int k = 0;
    LM_TRY(voxel_run(c, cur, m, leaf, other, &k, nullptr));
    std::swap(cur, other);
    m = k;
  }
  // :105 outlier removal: not offered here (out of scope) — the reference's default is "none"
  if (!(dist_near == 0.f && dist_far == 0.f) && m > 0) {  // :107-108, distance_filter.hpp:27-31
    int k = 0;
    LM_TRY(filter_run(c, cur, m, 1, dist_near, dist_far, other, &k));
    std::swap(cur, other);
    m = k;
  }
  if (m > 0) LM_CUDA(cudaMemcpyAsync(out_xyzi, cur, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  *n_out = m;
  return LMSF_OK;
}

int lmsf_map_set(lmsf_ctx* c, int kind, const float* xyzi, int n) {
  ENTER(c);
  if (kind < 0 || kind > 1 || n < 0 || (n > 0 && !xyzi)) return LMSF_ERR_INVALID;
  if (n == 0) return LMSF_OK;  // ceres_edgeSurfFeatureRegistration.hpp:58
  MapIndex& m = c->map[kind];
  if (n > m.cap) return LMSF_ERR_CAPACITY;
  LM_TRY(wait_map(c));
  const int chunk = c->prm.max_points;
  for (int o = 0; o < n; o += chunk) {
    int k = (n - o < chunk) ? n - o : chunk;
    memcpy(c->h_pts, xyzi + 4 * (size_t)o, (size_t)k * sizeof(float4));
    LM_CUDA(cudaMemcpyAsync(m.win + o, c->h_pts, (size_t)k * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    LM_CUDA(cudaStreamSynchronize(c->stream));
  }
  m.frame_n.assign(1, n);
  m.frame_pos.assign(1, {0.0, 0.0, 0.0});
  m.fixed = false;  // arbitrary cloud: bounding-box build
  m.cat = m.win;  // SetInputSource takes the cloud as given: no map voxel filter on this path
  m.grow0 = 2;
  return map_build(c, m, n, c->stream, false);
}

int lmsf_knn5(lmsf_ctx* c, int kind, const float* q_xyz, int nq, int32_t* idx5, float* d2_5) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (kind < 0 || kind > 1 || nq < 0 || (nq > 0 && (!q_xyz || !idx5 || !d2_5))) return LMSF_ERR_INVALID;
  if (!c->map[kind].ready) return LMSF_ERR_STATE;
  if (nq == 0) return LMSF_OK;
  char* buf = nullptr;
  LM_TRY(hook_scratch(c, (size_t)nq * 56 + 8, (void**)&buf));
  float* d_q = (float*)buf;
  int* d_i = (int*)(buf + (size_t)nq * 12);
  float* d_d = (float*)(buf + (size_t)nq * 32);
  LM_CUDA(cudaMemcpyAsync(d_q, q_xyz, (size_t)nq * 12, cudaMemcpyHostToDevice, c->stream));
  LM_TRY(knn_hook(c, kind, d_q, nq, d_i, d_d, (int*)(buf + (size_t)nq * 52)));
  LM_CUDA(cudaMemcpyAsync(idx5, d_i, (size_t)nq * 20, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(d2_5, d_d, (size_t)nq * 20, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  return LMSF_OK;
}

int lmsf_match(lmsf_ctx* c, int kind, const float* q_xyz, int nq, uint8_t* ok, double* out10) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (kind < 0 || kind > 1 || nq < 0 || (nq > 0 && (!q_xyz || !ok || !out10))) return LMSF_ERR_INVALID;
  if (!c->map[kind].ready) return LMSF_ERR_STATE;
  if (nq == 0) return LMSF_OK;
  char* buf = nullptr;
  LM_TRY(hook_scratch(c, (size_t)nq * 96 + 64, (void**)&buf));
  double* d_o = (double*)buf;                       // 80 B per query, 8-byte aligned first
  float* d_q = (float*)(buf + (size_t)nq * 80);
  uint8_t* d_ok = (uint8_t*)(buf + (size_t)nq * 92);
  LM_CUDA(cudaMemcpyAsync(d_q, q_xyz, (size_t)nq * 12, cudaMemcpyHostToDevice, c->stream));
  LM_TRY(match_hook(c, kind, d_q, nq, d_ok, d_o));
  LM_CUDA(cudaMemcpyAsync(ok, d_ok, (size_t)nq, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaMemcpyAsync(out10, d_o, (size_t)nq * 80, cudaMemcpyDeviceToHost, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  return LMSF_OK;
}

int lmsf_align_score(lmsf_ctx* c, int kind, const float* xyzi, int n, const float relpose16[16], double inlier_thresh,
                     double inlier_ratio_thresh, double* score, double* overlap, int32_t* n_inlier) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (kind < 0 || kind > 1 || n < 0 || (n > 0 && !xyzi) || !relpose16 || !score || !overlap) return LMSF_ERR_INVALID;
  if (!(inlier_thresh >= 0.0 && inlier_thresh <= 1.0)) return LMSF_ERR_INVALID;  // the index answers d2 < 1
  if (!c->map[kind].ready) return LMSF_ERR_STATE;
  if (n_inlier) *n_inlier = 0;
  if (n == 0) {  // alignEvaluate.hpp:61
    *score = 1.7976931348623157e308;
    *overlap = 0.0;
    return LMSF_OK;
  }
  float4* d_pts = nullptr;
  const size_t nblk = (size_t)(n + 127) / 128;
  LM_TRY(hook_scratch(c, (size_t)n * sizeof(float4) + (nblk + 1) * (sizeof(double) + sizeof(int)) + 64, (void**)&d_pts));
  double sum = 0;
  int cnt = 0;
  LM_CUDA(cudaMemcpyAsync(d_pts, xyzi, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  LM_TRY(align_hook(c, kind, d_pts, n, relpose16, (float)inlier_thresh, &sum, &cnt));
  const double ratio = (double)cnt / (double)n;  // :80
  *overlap = ratio;
  *score = (ratio > inlier_ratio_thresh) ? sum / cnt : 1.7976931348623157e308;  // :82-85
  if (n_inlier) *n_inlier = cnt;
  return LMSF_OK;
}

int lmsf_register(lmsf_ctx* c, const float* edge_xyzi, int n_e, const float* surf_xyzi, int n_s, int solver,
                  double pose[7], lmsf_reg_stats* st) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (!pose || (solver != LMSF_SOLVER_GN && solver != LMSF_SOLVER_HUBER_LM)) return LMSF_ERR_INVALID;
  LM_TRY(upload_features(c, edge_xyzi, n_e, surf_xyzi, n_s));
  return run_solver(c, solver, pose, st, n_e + n_s);
}

int lmsf_set_lm_outer(lmsf_ctx* c, int count) {
  ENTER(c);
  if (count < 0) return LMSF_ERR_INVALID;
  c->lm_count = count;
  return LMSF_OK;
}

int lmsf_tracker_step(lmsf_ctx* c, const float* xyzi, int n, double stamp, double delta[7], double pose_out[7],
                      lmsf_track_stats* st) {
  ENTER(c);
  if (!delta || !pose_out) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;  // before any slot state is touched
  LM_TRY(take_sweep(c, xyzi, n, true));
  return tracker_core(c, n, stamp, delta, pose_out, st);
}

int lmsf_tracker_step_dev(lmsf_ctx* c, const float* d_xyzi, int n, double stamp, double delta[7], double pose_out[7],
                          lmsf_track_stats* st) {
  ENTER(c);
  if (!delta || !pose_out || n < 0 || (n > 0 && !d_xyzi)) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;
  LM_TRY(take_sweep(c, d_xyzi, n, false));
  return tracker_core(c, n, stamp, delta, pose_out, st);
}

int lmsf_tracker_step_ticket(lmsf_ctx* c, int64_t ticket, double stamp, double delta[7], double pose_out[7],
                             lmsf_track_stats* st) {
  ENTER(c);
  if (!delta || !pose_out) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;
  int n = 0;
  LM_TRY(take_ticket(c, ticket, &n));
  return tracker_core(c, n, stamp, delta, pose_out, st);
}

int lmsf_tracker_submit_ticket(lmsf_ctx* c, int64_t ticket, double stamp, const double delta[7]) {
  ENTER(c);
  if (!delta) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;
  int n = 0;
  LM_TRY(take_ticket(c, ticket, &n));
  return tracker_begin(c, n, stamp, delta);
}

int lmsf_tracker_prefetch_cancel(lmsf_ctx* c, int64_t ticket) {
  ENTER(c);
  for (int si = 0; si < 2; ++si) {
    Ctx::FeatSlot& sl = c->slot[si];
    if (sl.filled && sl.seq == ticket) {
      sl.filled = false;  // the extraction already enqueued runs to completion; its slot is simply reusable
      return LMSF_OK;
    }
  }
  return LMSF_ERR_STATE;
}

int lmsf_tracker_submit(lmsf_ctx* c, const float* xyzi, int n, double stamp, const double delta[7]) {
  ENTER(c);
  if (!delta) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;
  LM_TRY(take_sweep(c, xyzi, n, true));
  return tracker_begin(c, n, stamp, delta);
}

int lmsf_tracker_submit_dev(lmsf_ctx* c, const float* d_xyzi, int n, double stamp, const double delta[7]) {
  ENTER(c);
  if (!delta || n < 0 || (n > 0 && !d_xyzi)) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;
  LM_TRY(take_sweep(c, d_xyzi, n, false));
  return tracker_begin(c, n, stamp, delta);
}

int lmsf_tracker_wait(lmsf_ctx* c, double delta_out[7], double pose_out[7], lmsf_track_stats* st) {
  ENTER(c);
  if (!delta_out || !pose_out) return LMSF_ERR_INVALID;
  const double ident[7] = {0, 0, 0, 1, 0, 0, 0};  // the initialising sweep has no motion increment
  for (int i = 0; i < 7; ++i) delta_out[i] = ident[i];
  return tracker_end(c, delta_out, pose_out, st);
}

int lmsf_tracker_prefetch(lmsf_ctx* c, const float* xyzi, int n, int64_t* ticket) {
  ENTER(c);
  return prefetch(c, xyzi, n, true, ticket);
}

int lmsf_tracker_prefetch_dev(lmsf_ctx* c, const float* d_xyzi, int n, int64_t* ticket) {
  ENTER(c);
  return prefetch(c, d_xyzi, n, false, ticket);
}

int lmsf_tracker_step_features(lmsf_ctx* c, const float* edge_xyzi, int n_e, const float* surf_xyzi, int n_s,
                               double stamp, double delta[7], double pose_out[7], lmsf_track_stats* st) {
  ENTER(c);
  if (!delta || !pose_out) return LMSF_ERR_INVALID;
  if (c->pending.active) return LMSF_ERR_STATE;
  LM_TRY(upload_features(c, edge_xyzi, n_e, surf_xyzi, n_s));
  return tracker_core(c, n_e + n_s, stamp, delta, pose_out, st);
}

int lmsf_tracker_reset(lmsf_ctx* c) {
  ENTER(c);
  LM_CUDA(cudaStreamSynchronize(c->stream_fe));
  LM_CUDA(cudaStreamSynchronize(c->stream_map));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  c->map_pending = false;
  c->pending = Ctx::Pending();
  for (int i = 0; i < 2; ++i) c->slot[i].freed_pending = c->slot[i].filled = false;
  c->init = false;
  for (int k = 0; k < 2; ++k) {
    c->map[k].frame_n.clear();
    c->map[k].frame_pos.clear();
    c->map[k].fixed = false;
    c->map[k].n_cells_seen = 0;
    c->map[k].ready = false;
    c->map[k].n_host = 0;
    c->map[k].cat = c->map[k].win;
  }
  c->curr = c->prev = c->motion = c->last_kf = rigid_identity();
  c->lm_count = c->prm.lm_outer_start;
  return LMSF_OK;
}

int lmsf_tracker_register_aux(lmsf_ctx* c, const float* xyzi, int n, double pose[7], lmsf_reg_stats* st) {
  ENTER(c);
  if (!pose) return LMSF_ERR_INVALID;
  LM_TRY(extract_current(c, xyzi, n, true));
  LM_TRY(scan_filter(c));
  rigid T = rigid_from_pose(pose);  // Solve(Isometry3d&): quaternion <-> matrix round trip
  double p[7];
  rigid_to_pose(T, p);
  LM_TRY(wait_map(c));
  LM_TRY(run_solver(c, c->prm.solver, p, st, n));
  T = rigid_from_pose(p);
  rigid_to_pose(T, pose);
  return LMSF_OK;
}

int lmsf_tracker_register_aux_features_dev(lmsf_ctx* c, const void* d_feat, int n_edge, int n_surf, double pose[7],
                                           lmsf_reg_stats* st) {
  ENTER(c);
  if (!pose || n_edge < 0 || n_surf < 0 || (n_edge + n_surf > 0 && !d_feat)) return LMSF_ERR_INVALID;
  if (n_edge + n_surf > c->prm.max_points) return LMSF_ERR_CAPACITY;
  if (c->pending.active) return LMSF_ERR_STATE;
  alias_slot(c, writable_slot(c));
  LM_TRY(wait_feat(c));
  c->feat_from_extract = false;
  if (n_edge + n_surf)
    LM_CUDA(cudaMemcpyAsync(c->d_feat, d_feat, (size_t)(n_edge + n_surf) * sizeof(float4), cudaMemcpyDeviceToDevice,
                            c->stream));
  LM_TRY(set_counts(c, n_edge, n_surf));
  LM_TRY(scan_filter(c));
  rigid T = rigid_from_pose(pose);  // Solve(Isometry3d&): quaternion <-> matrix round trip
  double p[7];
  rigid_to_pose(T, p);
  LM_TRY(wait_map(c));
  LM_TRY(run_solver(c, c->prm.solver, p, st, n_edge + n_surf));
  T = rigid_from_pose(p);
  rigid_to_pose(T, pose);
  return LMSF_OK;
}

int lmsf_get_map(lmsf_ctx* c, int kind, float* xyzi, int cap, int* n) {
  ENTER(c);
  LM_TRY(wait_map(c));
  if (kind < 0 || kind > 1 || !n) return LMSF_ERR_INVALID;
  const MapIndex& m = c->map[kind];
  *n = m.n_host;
  if (!xyzi) return LMSF_OK;
  if (cap < m.n_host) return LMSF_ERR_CAPACITY;
  if (m.n_host) {
    LM_CUDA(cudaMemcpyAsync(xyzi, m.cat, (size_t)m.n_host * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    LM_CUDA(cudaStreamSynchronize(c->stream));
  }
  return LMSF_OK;
}

int lmsf_dev_alloc(lmsf_ctx* c, int64_t bytes, void** d_ptr) {
  ENTER(c);
  if (bytes <= 0 || !d_ptr) return LMSF_ERR_INVALID;
  LM_CUDA(cudaMalloc(d_ptr, (size_t)bytes));
  return LMSF_OK;
}

int lmsf_dev_free(lmsf_ctx* c, void* d_ptr) {
  ENTER(c);
  LM_CUDA(cudaStreamSynchronize(c->stream));
  LM_CUDA(cudaFree(d_ptr));
  return LMSF_OK;
}

int lmsf_dev_upload(lmsf_ctx* c, void* d_dst, const void* h_src, int64_t bytes) {
  ENTER(c);
  if (!d_dst || !h_src || bytes < 0) return LMSF_ERR_INVALID;
  LM_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  return LMSF_OK;
}

/* tuning aid (not part of the documented ABI): host microseconds spent inside 0 = solve (enqueue + wait for the
 * pose), 1 = local-map update enqueue, 2 = solve enqueue only; counts in n[] */
int lmsf_debug_host_times(lmsf_ctx* c, double us[4], int64_t n[4], int reset) {
  if (!c) return LMSF_ERR_INVALID;
  for (int i = 0; i < 4; ++i) {
    us[i] = c->host_us[i];
    n[i] = c->host_n[i];
    if (reset) {
      c->host_us[i] = 0;
      c->host_n[i] = 0;
    }
  }
  return LMSF_OK;
}

int lmsf_profile_enable(lmsf_ctx* c, int enable) {
  ENTER(c);
  c->prof = enable != 0;
  return LMSF_OK;
}

int lmsf_profile_read(lmsf_ctx* c, double ms[LMSF_N_STAGES], int64_t launches[LMSF_N_STAGES], double* match_alg_bytes,
                      int reset) {
  ENTER(c);
  LM_CUDA(cudaStreamSynchronize(c->stream_fe));
  LM_CUDA(cudaStreamSynchronize(c->stream_map));
  LM_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& sp : c->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) c->prof_ms[sp.stage] += t;
    c->ev_pool.push_back(sp.a);
    c->ev_pool.push_back(sp.b);
  }
  c->spans.clear();
  for (int i = 0; i < LMSF_N_STAGES; ++i) {
    if (ms) ms[i] = c->prof_ms[i];
    if (launches) launches[i] = c->prof_launch[i];
  }
  if (match_alg_bytes) *match_alg_bytes = c->match_bytes;
  if (reset) {
    for (int i = 0; i < LMSF_N_STAGES; ++i) {
      c->prof_ms[i] = 0;
      c->prof_launch[i] = 0;
    }
    c->match_bytes = 0;
  }
  return LMSF_OK;
}

}  // extern "C"
