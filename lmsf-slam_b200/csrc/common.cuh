// common.cuh — context layout, launch/err helpers shared by the .cu files of liblmsf_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <string>
#include <vector>

#include "../../include/lmsf_b200.h"
#include "dmath.cuh"

namespace lm {

// ---------------------------------------------------------------- errors
#define LM_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      c->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                    \
      return LMSF_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define LM_TRY(expr)          \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != LMSF_OK) return rc__; \
  } while (0)

// launch on the context's stream and count it
#define LM_LAUNCH(c, kern, grid, block, smem, ...)                    \
  do {                                                                \
    kern<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__);      \
    (c)->launches++;                                                  \
  } while (0)
// same, on an explicit stream (the local-map update runs on its own stream)
#define LM_LAUNCH_ON(c, st, kern, grid, block, smem, ...)             \
  do {                                                                \
    kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);             \
    (c)->launches++;                                                  \
  } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- feature extraction state
struct ExtractBufs {
  int cap = 0;        // max points per sweep
  int nblk_cap = 0;   // ceil(cap / RING_BLOCK)
  int* ring_id = nullptr;      // [cap] ring of each input point or -1
  int* blk_cnt = nullptr;      // [64][nblk_cap] per-block ring histogram -> exclusive offsets
  int* ring_cnt = nullptr;     // [64] points per ring
  int* ring_off = nullptr;     // [65] exclusive offsets of rings in ring order
  float4* ring_pts = nullptr;  // [cap] points grouped by ring, firing order kept
  int* ring_src = nullptr;     // [cap] original index of ring_pts[i]
  double* curv = nullptr;      // [cap] curvature by ring-order position
  int* sorted = nullptr;       // [cap] per sector: ring-local ids ascending by (curvature, id)
  double* sort_key = nullptr;  // [2*cap] global scratch for sectors too long for shared memory
  int* sort_val = nullptr;     // [2*cap]
  uint8_t* flag = nullptr;     // [cap] disable flags for rings too long for shared memory
  uint8_t* btype = nullptr;    // [cap] bad-point classes, same
  uint8_t* is_edge = nullptr;  // [cap] by ring-order position
  int* edge_ids = nullptr;     // [64*6*20] picked edges per sector in pick order
  int* sec_cnt = nullptr;      // [64*6] edges picked per sector
  uint8_t* label = nullptr;    // [cap] by original index
  int* surf_rank = nullptr;    // [cap] scratch: feature index of the surf at a ring-order position
  int* counts = nullptr;       // [4] n_edge, n_surf, n_valid
  int* rotary = nullptr;       // [4] RotaryLidarPreProcess state of the sweep in flight: first / last finite point, the
                               //     point at which half_passed turns true
};

// ---------------------------------------------------------------- local-map index
// One coarse cell (edge = 1 m, the kNN radius) of the hash grid.  32 bytes so a
// probe is a single 32-byte sector.
struct __align__(32) CellRec {
  unsigned long long key;   // packed relative cell coords, ~0 = empty slot
  unsigned long long mask;  // occupancy of the 4x4x4 L1 sub-cells (0.25 m)
  int start, end;           // range in the cell-sorted point array
  int fine_base;            // index of this cell's first occupied L1 cell (l1[])
  int pad;
};

// One occupied L1 cell (0.25 m): occupancy of its 4x4x4 L2 cells and the index of its first L2 cell in l2_start.
// 16 bytes: one load per visited cell.
struct __align__(16) L1Rec {
  unsigned long long mask;
  int first;
  int pad;
};

struct MapDev {  // device-visible description of one map index (lives in device memory)
  int n;              // points in the sorted array (finite points of the map)
  int min_c[3];       // coarse cell coords of the bbox minimum
  int dim[3];         // coarse cells per axis
  int bits[3];        // bits per axis in the packed key
  unsigned table_mask;  // slots - 1
  int n_fine;         // occupied fine cells
  int grow0;          // kNN: first grown box (half width in L2 cells) when START's block holds fewer than five points;
                      // 2 for raw clouds, wider over a voxel-filtered map whose points are a leaf apart (0 = 2)
};

struct MapIndex {
  int cap = 0;
  int n_host = 0;            // points in cat (host copy of the count)
  bool ready = false;
  float4* win = nullptr;     // [cap] raw sliding window: keyframe clouds concatenated oldest -> newest
  float4* win_alt = nullptr; // [cap] double buffer used when the oldest frame is evicted
  float4* vox = nullptr;     // [cap] voxel-filtered window (only when a map leaf size is set)
  float4* cat = nullptr;     // not owned: win or vox — the map cloud the index is built over
  float4* sorted = nullptr;  // [cap] cell-sorted copy, .w = bit pattern of the original index
  float4* grouped = nullptr;               // [cap] points grouped by L0 cell (build scratch), .w = original index
  int2* slot_rank = nullptr;               // [cap] build scratch: hash slot of the point's L0 cell, rank inside it
  int* cell_list = nullptr;                // [cap] build scratch: hash slots of the occupied L0 cells
  int* l2_start = nullptr;                 // [2 cap+1] start of every occupied L2 cell; one sentinel per L0 cell
  L1Rec* l1 = nullptr;                     // [cap] one record per occupied L1 cell
  int* d_cnt = nullptr;                    // [8] device counters: L2 entries, L1 cells, L0 cells, insert-failed flag, point cursor
  unsigned* d_box = nullptr;               // [8] bbox scratch of this map (ordered-uint min/max, finite count)
  int* h_cnt = nullptr;                    // [4] pinned mirror of d_cnt (read back with the next pose)
  int grow0 = 2;                           // density hint for the search (MapDev::grow0), set by the caller of map_build
  bool fixed = false;                      // grid origin / key width frozen (tracker maps): builds need no host sync
  int n_cells_seen = 0;                    // occupied L0 cells of the last build that was read back
  CellRec* table = nullptr;                // [table_cap]
  unsigned table_cap = 0;
  MapDev* dev = nullptr;     // device copy
  MapDev host;               // host mirror (after build)
  std::vector<int> frame_n;  // sliding window frame sizes, oldest first
  std::vector<std::array<double, 3>> frame_pos;  // sensor position of every window frame (frozen-grid cover test)
};

// ---------------------------------------------------------------- solver state (device resident)
#define LM_NSUM 30  // 21 (H upper) + 6 (g) + cost + rows + edge rows
struct SolveState {
  double x[7];        // current pose estimate {qx,qy,qz,qw,tx,ty,tz}
  double cand[7];     // LM candidate
  double H[36], g[6]; // normal equations at x
  double cost;        // cost at x
  double scale[6];    // Jacobi column scaling of this solve
  double radius, decrease;
  double model_change;
  double x_norm;
  double gn_map[36];  // GN degeneracy remap
  int n_edge_ok, n_surf_ok;
  int gn_done, gn_degenerate, gn_iters;
  int lm_active;      // inner loop still running
  int lm_iter;        // trust-region iterations used in this solve
  int lm_invalid;
  int lm_steps_total, lm_steps_accepted;
  unsigned ticket;    // last-block-done counter
  int knn_next;       // work-queue head of the running k_knn launch (light queries, chunks of 32)
  int defer_next;     // work-queue head of the k_knn_sparse launch that follows (one deferred query per warp)
  int n_defer;        // queries the running k_knn launch handed to k_knn_sparse (positions in Ctx::d_defer)
  int n_heavy;        // queries whose own map cell is (nearly) empty; they sort first
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  lmsf_params prm;
  std::string last_error;
  int64_t launches = 0;

  // sweep + features
  float4* d_sweep = nullptr;   // [max_points]
  float4* d_feat = nullptr;    // alias of slot[slot_cur].feat: edges first, then surfs
  float4* d_tmp = nullptr;     // [max_points] scratch (transformed frame / voxel input)
  ExtractBufs ex;
  MapIndex map[2];

  // per-feature match records (7 doubles SoA) + validity
  double* d_rec = nullptr;     // [6][max_points] edge: a, b   surf: n, D  (by sorted query position)
  float* d_recf = nullptr;     // [3][max_points] scan point, LiDAR frame
  uint8_t* d_ok = nullptr;     // [max_points] 0 none, 1 edge, 2 surf
  unsigned long long* q_keys = nullptr;      // [max_points] query sort keys (cell of the world point)
  unsigned long long* q_keys_alt = nullptr;
  int* q_vals = nullptr;                     // [max_points] feature index
  int* q_vals_alt = nullptr;                 // sorted: position -> feature index
  float4* d_pw = nullptr;      // [max_points] fp32 world point by feature index
  int* d_nbr = nullptr;        // [5][max_points] neighbour indices by sorted query position
  float4* d_defer = nullptr;   // [max_points] the queries k_knn left to k_knn_sparse: world point, .w = position | edge flag
  double* d_partial = nullptr; // [grid][LM_NSUM]
  int partial_blocks = 0;
  SolveState* d_state = nullptr;
  unsigned* d_bar = nullptr;   // [1] generation word of k_solve's grid barrier

  // voxel scratch (sized for max(max_points, max_map_points))
  int vox_cap = 0;
  unsigned long long* v_keys = nullptr;
  unsigned long long* v_keys_alt = nullptr;
  int* v_vals = nullptr;
  int* v_vals_alt = nullptr;
  int* v_heads = nullptr;
  uint8_t* v_flags = nullptr;
  float4* v_sorted = nullptr;
  float4* v_in = nullptr;
  float4* v_out = nullptr;
  int* v_member = nullptr;
  void* v_params = nullptr;  // VoxelParams on device
  unsigned* d_bbox = nullptr;  // [8] ordered-uint min/max + finite count

  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  void* hook_buf = nullptr;    // grow-only scratch of the test / loop-closure hooks (lmsf_knn5, lmsf_match, lmsf_align_score)
  size_t hook_bytes = 0;

  // local-map update pipeline: its own stream, ordered against the main stream with two events
  cudaStream_t stream_map = nullptr;
  cudaStream_t stream_map2 = nullptr;   // the edge map's index build runs beside the surf map's
  cudaEvent_t ev_map_fork = nullptr, ev_map_join = nullptr;
  cudaEvent_t ev_map_done = nullptr;   // map stream has finished the index builds
  bool map_pending = false;            // main stream has not yet waited for ev_map_done
  bool feat_from_extract = false;      // d_feat came from extract_run: finite and range-gated

  // front end (feature extraction) pipeline: its own stream and two feature slots, so that the extraction of
  // the next sweep overlaps the registration of the current one.  d_feat / ex.counts alias slot[slot_cur].
  struct FeatSlot {
    float4* feat = nullptr;         // [max_points] edges first, then surfs
    int* counts = nullptr;          // [4] n_edge, n_surf (device)
    int* perm = nullptr;            // [max_points] processing order of the registration: position -> feature index
    cudaEvent_t ready = nullptr;    // front-end stream: extraction into this slot finished
    cudaEvent_t freed = nullptr;    // map stream: the local-map update has consumed this slot
    bool freed_pending = false;     // nobody has waited for `freed` yet
    bool filled = false;            // holds a prefetched sweep that no step has consumed yet
    int n = 0;                      // its point count
    int64_t seq = 0;                // the ticket lmsf_tracker_prefetch* handed out for it (order of the prefetches)
  };
  FeatSlot slot[2];
  int slot_cur = 0;
  int* d_perm = nullptr;          // alias of slot[slot_cur].perm
  bool perm_valid = false;        // d_perm describes d_feat (features from our own extraction, unfiltered)
  int64_t prefetch_seq = 0;
  cudaStream_t stream_fe = nullptr;
  float4* h_stage[2] = {nullptr, nullptr};  // pinned staging of host sweeps, alternating
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};  // H2D out of h_stage[i] finished
  int stage_next = 0;

  // pinned host staging
  float4* h_pts = nullptr;     // [max_points]
  double* h_pose = nullptr;    // [16]
  int* h_ints = nullptr;       // [64]
  SolveState* h_state = nullptr;

  // tracker (LidarTrackerLocalMap members)
  struct Pending {            // a submitted sweep whose pose has not been waited for (lmsf_tracker_submit / _wait)
    bool active = false;
    bool first = false;       // the initialising sweep: nothing was enqueued
    double stamp = 0;
    int solver = 0, outer = 0;
  } pending;
  bool init = false;
  rigid prev, curr, motion, last_kf;
  double last_kf_time = 0;
  int lm_count = 10;
  int n_edge = 0, n_surf = 0;  // features currently in d_feat

  // profiling
  bool prof = false;
  struct Span {
    int stage;
    cudaEvent_t a, b;
  };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[LMSF_N_STAGES] = {0, 0, 0, 0, 0, 0, 0};
  int64_t prof_launch[LMSF_N_STAGES] = {0, 0, 0, 0, 0, 0, 0};
  double match_bytes = 0;
  int64_t match_launches = 0;  // correspondence passes (k_assoc + sort + k_knn + k_fit)

  double host_us[4] = {0, 0, 0, 0};  // tuning aid: host time per phase (lmsf_debug_host_times)
  int64_t host_n[4] = {0, 0, 0, 0};

  // loop-closure descriptor database (scancontext.cu), allocated on first use
  void* scdb = nullptr;
};

struct StageScope {  // records CUDA events around a stage when profiling is on
  Ctx* c;
  int stage;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  int64_t l0;
  StageScope(Ctx* c_, int s, cudaStream_t stream = nullptr);
  ~StageScope();
};

// ---- implemented in extract.cu
int extract_alloc(Ctx* c);
void extract_free(Ctx* c);
// sweep d_in (n points, device) -> feat_out (edges, then surfs), counts_out[0..1], labels in ex.label; all launches
// on stream `st`.  The scratch in ExtractBufs is shared: extractions must be ordered on one stream (the front end's).
int extract_run(Ctx* c, const float4* d_in, int n, cudaStream_t st, float4* feat_out, int* counts_out, int* perm_out);
// RotaryLidarPreProcess::Process on d_pts[0..n) in place (intensity := relative time; non-finite points are passed over as
// if removeNaN had dropped them), on stream `st`
int rotary_apply(Ctx* c, float4* d_pts, int n, float scan_period, cudaStream_t st);

// ---- implemented in voxel.cu
int voxel_alloc(Ctx* c);
void voxel_free(Ctx* c);
// device in -> device out; *n_out read back (synchronises the stream)
int voxel_run(Ctx* c, const float4* d_in, int n, float leaf, float4* d_out, int* n_out, int* d_member);

// order-preserving point filter on the main stream: mode 0 = finite x/y/z, mode 1 = near < |p| < far (fp32 norm)
int filter_run(Ctx* c, const float4* d_in, int n, int mode, float near_t, float far_t, float4* d_out, int* n_out);

// ---- implemented in mapindex.cu
int map_alloc(Ctx* c, MapIndex& m, int cap);
void map_free(MapIndex& m);
// (re)build the grid index over m.cat[0..n) on stream `st`.  fixed_grid = false: bounding box first (one host
// sync; any input).  fixed_grid = true: the map's frozen origin / key width is reused, nothing is read back
// (tracker maps whose points are known to lie inside the frozen grid).
int map_build(Ctx* c, MapIndex& m, int n, cudaStream_t st, bool fixed_grid);
// freeze the grid of a tracker map around a sensor position (1024 x 1024 x 256 m); false if p is unusable
bool map_freeze_grid(MapIndex& m, const double p[3]);
// true when a sweep taken at sensor position p (range-gated to `reach` metres) lies inside the frozen grid
bool map_grid_covers(const MapIndex& m, const double p[3], double reach);
// make the main stream wait for the pending local-map work (no-op when nothing is pending)
int wait_map(Ctx* c);
int wait_feat(Ctx* c);

// ---- implemented in scancontext.cu
void scdb_free(Ctx* c);

// ---- implemented in match.cu
int solve_alloc(Ctx* c);
void solve_free(Ctx* c);
// grow-only scratch arena of the hooks: `bytes` of device memory, valid until the next hook call (a regrow
// synchronises the device, steady-state calls do not allocate)
int hook_scratch(Ctx* c, size_t bytes, void** out);
// d_work: nq + 2 ints of scratch (deferral counters and list)
int knn_hook(Ctx* c, int kind, const float* d_q, int nq, int* d_idx, float* d_d2, int* d_work);
int match_hook(Ctx* c, int kind, const float* d_q, int nq, uint8_t* d_ok, double* d_out10);
int align_hook(Ctx* c, int kind, const float4* d_pts, int n, const float T12[12], float thresh, double* sum, int* cnt);
// features in c->d_feat (counts in c->ex.counts on the device; `upper` bounds their sum on the
// host); enqueues the whole solve, reads pose + statistics back (one stream sync)
int solve_run(Ctx* c, int solver, double pose[7], lmsf_reg_stats* st, int upper, int outer_count);
// the two halves of solve_run: enqueue everything (no host wait) / wait for the pose and read the statistics
int solve_enqueue(Ctx* c, int solver, const double pose[7], int upper, int outer_count);
int solve_finish(Ctx* c, int solver, double pose[7], lmsf_reg_stats* st, int outer_count);

}  // namespace lm

struct lmsf_ctx : lm::Ctx {};
