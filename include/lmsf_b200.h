/*
 * lmsf_b200.h — C ABI of the B200-native LiDAR scan-to-map registration hot path.
 *
 * Drop-in boundary for Robot-WH/LMSF-Slam (src/MultiSensorFusionEstimator3D,
 * paths below are relative to its include/ directory).  Every entry point is
 * extern "C", takes plain pointers and sizes (caller-owned HOST buffers unless
 * the name ends in _dev) and returns an int status: 0 = OK, <0 = error
 * (lmsf_strerror).  The reference's three seams return void and have no error
 * convention (registration_base.hpp:31-33, process_base.hpp:37,
 * Filter/filter_base.hpp:34); the C++ adapters in lmsf_b200_adapters.hpp log a
 * non-zero status and leave the output untouched, which is what the reference
 * does on failure (edgeSurfFeatureRegistration.hpp:221-225).
 *
 * There is NO CPU fallback behind this ABI: every call runs hand-written
 * sm_100a CUDA kernels on the context's device and fails with
 * LMSF_ERR_NO_DEVICE when none is usable.
 *
 * A context is one LiDAR's worth of state (one CUDA stream, pre-sized device
 * arenas, the local-map index, the tracker state).  Contexts share nothing
 * and may be used concurrently from different threads, one thread per context
 * at a time — the reference runs one processor/tracker instance per LiDAR
 * under `omp parallel for` (System/ML_System.hpp:137-141,248-256).
 *
 * Layouts: a point is 4 packed floats {x, y, z, intensity}; a pose is 7
 * doubles {qx, qy, qz, qw, tx, ty, tz} — the parameter block layout of
 * ceres_edgeSurfFeatureRegistration.hpp:38-40.
 */
#ifndef LMSF_B200_H_
#define LMSF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMSF_OK 0
#define LMSF_ERR_INVALID (-1)   /* bad argument */
#define LMSF_ERR_NO_DEVICE (-2) /* no usable CUDA device: there is no CPU path */
#define LMSF_ERR_CUDA (-3)      /* a CUDA call failed; see lmsf_last_cuda_error */
#define LMSF_ERR_CAPACITY (-4)  /* input larger than the context was sized for */
#define LMSF_ERR_STATE (-5)     /* call order (e.g. register before map_set) */

#define LMSF_KIND_EDGE 0 /* "loam_edge" */
#define LMSF_KIND_SURF 1 /* "loam_surf" */

#define LMSF_SOLVER_GN 0      /* EdgeSurfFeatureRegistration (edgeSurfFeatureRegistration.hpp:27) */
#define LMSF_SOLVER_HUBER_LM 1 /* CeresEdgeSurfFeatureRegistration (factory default, ML_SystemFactory.hpp:189) */

typedef struct lmsf_ctx lmsf_ctx;

/* Construction parameters.  Defaults (lmsf_params_default) are the constants
 * the reference hard-codes; the comment gives the file:line of each. */
typedef struct {
  int32_t n_scans;           /* 16; 16|32|64 (ML_SystemFactory.hpp:197, LOAMFeatureProcessor_base.hpp:309-340) */
  float min_range;           /* 2    (ML_SystemFactory.hpp:197) */
  float max_range;           /* 80 */
  float edge_thresh;         /* 1.0  (LOAMFeatureProcessor_base.hpp:37) */
  int32_t remove_bad_points; /* 1    (LOAMFeatureProcessor_base.hpp:38) */
  int32_t max_points;        /* capacity of one sweep, default 262144 */
  int32_t window;            /* 10   sliding-window keyframes (ndt_test.yaml:18) */
  int32_t solver;            /* LMSF_SOLVER_HUBER_LM */
  float map_leaf_edge;       /* 0 = raw concatenation (shipped tracker); >0 voxel-filter the map */
  float map_leaf_surf;
  float scan_leaf_edge;      /* 0 = off; >0 voxel-filter scan features before registration */
  float scan_leaf_surf;
  int32_t gn_max_iters;      /* 10   (edgeSurfFeatureRegistration.hpp:65) */
  int32_t lm_outer_start;    /* 10, decremented before every solve, floor 2 (ceres_...:46,100-101) */
  int32_t lm_inner_iters;    /* 4    (ceres_...:118) */
  float huber_delta;         /* 0.1  (ceres_...:107); read as the decimal it was written as ("%.6g" -> double),
                                because the reference's literal is the double 0.1, not (double)0.1f */
  double kf_trans;           /* 0.3 m   (LidarTrackerLocalMap.hpp:65) */
  double kf_rot;             /* 0.1 rad */
  double kf_time;            /* 10 s */
  int32_t max_map_points;    /* capacity of one local map; default window * max_points */
  float rotary_scan_period;  /* 0 = off (sweeps arrive preprocessed); > 0: every sweep that enters the extraction first
                                goes through the node's removeNaN + RotaryLidarPreProcess::Process
                                (Preprocess/RotaryLidar_preprocessing.hpp:31-71: intensity := relative time in the
                                sweep, SCAN_PERIOD_ = this value, 0.1 in the reference), fused into the ring pass */
  int32_t reserved[10];      /* must be zero */
} lmsf_params;

typedef struct {
  int32_t outer_iters;       /* outer iterations executed */
  int32_t n_edge_matched;    /* accepted correspondences in the last outer iteration */
  int32_t n_surf_matched;
  int32_t converged;         /* GN: broke on the convergence test (edgeSurf...:326) */
  int32_t degenerate;        /* GN: isDegenerate (edgeSurf...:290) */
  int32_t lm_steps_total;    /* LM: trust-region iterations over all outer iterations */
  int32_t lm_steps_accepted;
  int32_t pad;
  double final_cost;         /* GN: 0.5*sum r^2 of last iteration; LM: 0.5*sum rho(r^2) */
} lmsf_reg_stats;

typedef struct {
  int32_t n_edge, n_surf;    /* features extracted from this sweep */
  int32_t keyframe;          /* 0 none, 1 motion, 2 time (LidarTrackerLocalMap.hpp:239-262) */
  int32_t map_edge, map_surf; /* local-map sizes after this step */
  int32_t first;             /* 1 on the initialising sweep (LidarTrackerLocalMap.hpp:111-121) */
  lmsf_reg_stats reg;
} lmsf_track_stats;

/* ---- lifetime ------------------------------------------------------- */
int lmsf_params_default(lmsf_params* p);
int lmsf_ctx_create(int device, const lmsf_params* p, lmsf_ctx** out);
void lmsf_ctx_destroy(lmsf_ctx* c);
const char* lmsf_strerror(int code);
const char* lmsf_last_cuda_error(lmsf_ctx* c);
/* number of kernels this context has launched so far (bench.py gpu_launches) */
int64_t lmsf_launch_count(lmsf_ctx* c);
/* raw cudaStream_t of the context, for CUDA-event timing on the launching stream */
void* lmsf_stream(lmsf_ctx* c);

/* ---- seam 1: feature extraction ------------------------------------ */
/* Replaces LOAMFeatureProcessorBase::Process (FeatureExtract/LOAMFeatureProcessor_base.hpp:59-126):
 * splitScan :290-343, checkBadEdgePoint :216-282, curvature :97-118,
 * featureExtractionFromSector :145-207.  label_out (optional, n bytes):
 * 0 = neither, 1 = "loam_edge", 2 = "loam_surf".  edge_xyzi / surf_xyzi
 * receive the clouds in the reference's order (rings 0..R-1, sectors 0..5;
 * edges by descending curvature, surfs by ascending curvature); each must
 * have room for n points. */
int lmsf_extract_features(lmsf_ctx* c, const float* xyzi, int n, uint8_t* label_out,
                          float* edge_xyzi, int* n_edge, float* surf_xyzi, int* n_surf);

/* ---- seam 2: voxel-grid filter -------------------------------------- */
/* Replaces FilterBase::Filter with a pcl::VoxelGrid (Filter/filter_base.hpp:34-45,
 * Filter/voxel_grid.hpp:25-29, factory/processing/pointcloud/filter/filter_factory.hpp:36-41).
 * out_xyzi needs room for n points; voxel_of_point (optional, n ints) receives
 * the output row each input point was averaged into (-1 for a non-finite point). */
int lmsf_voxel_downsample(lmsf_ctx* c, const float* xyzi, int n, float leaf, float* out_xyzi,
                          int* n_out, int32_t* voxel_of_point);

/* = PointCloudCommonProcess::Process (Algorithm/PointClouds/processing/common_processing.hpp:87-111), the front
 * end of the direct-method branch: (1) pcl::removeNaNFromPointCloud when remove_nan != 0, (2) the VoxelGrid
 * down-sampling when leaf > 0, (3) — outlier removal is not offered (out of scope; the reference's default is
 * none) — (4) DistanceFilter (Filter/distance_filter.hpp:24-44): keep near < |p| < far, skipped when both are 0.
 * Order preserving.  out_xyzi needs room for n points. */
int lmsf_common_process(lmsf_ctx* c, const float* xyzi, int n, int remove_nan, float leaf, float dist_near,
                        float dist_far, float* out_xyzi, int* n_out);
/* Replaces what the node does to a raw sweep before the estimator sees it (src/apps/src/MultiLidarSLAM_node.cpp:126-133,
 * Preprocess/RotaryLidar_preprocessing.hpp:31-104): pcl::removeNaNFromPointCloud (x, y, z finite; order kept), then
 * RotaryLidarPreProcess<PointXYZI>::Process — the per-point relative time in the sweep, from the azimuth between the
 * first and the last point (findStartEndAngle :80-94) with the half-sweep state machine of :38-69, written into the
 * intensity channel (:100-104).  scan_period = SCAN_PERIOD_ (0.1).  out_xyzi needs room for n points. */
int lmsf_rotary_preprocess(lmsf_ctx* c, const float* xyzi, int n, float scan_period, float* out_xyzi, int* n_out);

/* ---- seam 3: registration ------------------------------------------- */
/* = RegistrationBase::SetInputSource (registration/registration_base.hpp:31): upload a
 * local map and build its kNN index (replaces pcl::KdTreeFLANN::setInputCloud,
 * FeatureMatch/FeatureMatchBase.hpp:40-44).  An empty map is ignored like
 * ceres_edgeSurfFeatureRegistration.hpp:58. */
int lmsf_map_set(lmsf_ctx* c, int kind, const float* xyzi, int n);
/* Test hook for nearestKSearch(point, 5) (EdgeFeatureMatch.hpp:38, surfFeatureMatch.hpp:37).
 * Neighbours are ascending by (squared distance, index); only neighbours with
 * squared distance < 1.0 (FeatureMatchBase.hpp:29) are reported, the rest of
 * the five slots hold idx -1 and d2 +inf — the matchers reject such queries. */
int lmsf_knn5(lmsf_ctx* c, int kind, const float* q_xyz, int nq, int32_t* idx5, float* d2_5);
/* Test hook for EdgeFeatureMatch::Match (EdgeFeatureMatch.hpp:33-87) / SurfFeatureMatch::Match
 * (surfFeatureMatch.hpp:32-87) on world-frame fp32 points.  out10 per query:
 * kind 0: {nx,ny,nz, r, ax,ay,az, bx,by,bz}; kind 1: {nx,ny,nz, r, D, 0,0,0,0,0}. */
int lmsf_match(lmsf_ctx* c, int kind, const float* q_xyz, int nq, uint8_t* ok, double* out10);
/* = PointCloudAlignmentEvaluate::AlignmentScore (registration/alignEvaluate.hpp:55-87) against the map index
 * `kind` (SetTargetPoints :42-46 = lmsf_map_set): the cloud is transformed by the row-major 4x4 fp32 relpose
 * (pcl::transformPointCloud with a Matrix4f), every point looks up its nearest target point, points with
 * squared distance <= inlier_thresh are inliers.  overlap = inliers / n; score = mean inlier squared distance
 * when overlap > inlier_ratio_thresh, else DBL_MAX (also for an empty cloud).  inlier_thresh must be <= 1.0
 * (the reference's callers use 0.1 and 1, loopDetection.hpp:177-178,411-412, backend_lifelong.hpp:319-320). */
int lmsf_align_score(lmsf_ctx* c, int kind, const float* xyzi, int n, const float relpose16[16],
                     double inlier_thresh, double inlier_ratio_thresh, double* score, double* overlap,
                     int32_t* n_inlier);
/* = SetInputTarget + Solve (registration_base.hpp:32-33): pose holds the
 * prediction on entry and the result on return.  solver: LMSF_SOLVER_*; the
 * Huber-LM outer-iteration budget is context state, as in the reference. */
int lmsf_register(lmsf_ctx* c, const float* edge_xyzi, int n_e, const float* surf_xyzi, int n_s,
                  int solver, double pose[7], lmsf_reg_stats* st);
/* CeresEdgeSurfFeatureRegistration::SetMaxIteration / reset of the stateful budget */
int lmsf_set_lm_outer(lmsf_ctx* c, int count);

/* ---- a6: scan-to-map tracker, device resident ----------------------- */
/* = LOAMFeatureProcessorBase::Process followed by LidarTrackerLocalMap::Solve
 * (LidarTracker/LidarTrackerLocalMap.hpp:107-160): prediction, registration,
 * keyframe test :239-262, local-map update and index rebuild :205-232.
 * delta: in = caller's motion prior (identity {0,0,0,1,0,0,0} = use the
 * constant-velocity model), out = motion increment.  pose_out = curr_pose_. */
int lmsf_tracker_step(lmsf_ctx* c, const float* xyzi, int n, double stamp, double delta[7],
                      double pose_out[7], lmsf_track_stats* st);
/* Same, with the sweep already resident in device memory (d_xyzi is a device
 * pointer on the context's device). */
int lmsf_tracker_step_dev(lmsf_ctx* c, const float* d_xyzi, int n, double stamp, double delta[7],
                          double pose_out[7], lmsf_track_stats* st);
/* Front-end / back-end pipelining.  The reference runs LOAMFeatureProcessorBase::Process on the sensor
 * thread and LidarTrackerLocalMap::Solve on estimate_thread_, connected by a queue
 * (System/ML_System.hpp:137-141, :210-227): the features of sweep k+1 are extracted while sweep k is
 * registered.  lmsf_tracker_prefetch(_dev) enqueues the upload and feature extraction of a COMING sweep on
 * the context's front-end stream, hands back a ticket (> 0) and returns at once; lmsf_tracker_step_ticket /
 * lmsf_tracker_submit_ticket consume exactly that sweep.  A prefetched sweep is identified by its ticket only
 * (never by the caller's pointer), a host sweep is copied at prefetch time, at most two sweeps wait at a time
 * (a third prefetch returns LMSF_ERR_STATE), lmsf_tracker_prefetch_cancel gives a waiting one up, and a
 * ticket that was consumed, cancelled or displaced is refused with LMSF_ERR_STATE.  lmsf_tracker_step /
 * _submit with a sweep pointer always extract that sweep now. */
int lmsf_tracker_prefetch(lmsf_ctx* c, const float* xyzi, int n, int64_t* ticket);
int lmsf_tracker_prefetch_dev(lmsf_ctx* c, const float* d_xyzi, int n, int64_t* ticket);
int lmsf_tracker_prefetch_cancel(lmsf_ctx* c, int64_t ticket);
int lmsf_tracker_step_ticket(lmsf_ctx* c, int64_t ticket, double stamp, double delta[7], double pose_out[7],
                             lmsf_track_stats* st);
/* lmsf_tracker_step in two halves, for callers that have host work to do while the GPU registers the sweep
 * (typically: prefetch the next one).  submit = prediction + the whole registration enqueued, returns at once;
 * wait = pose read-back, motion update, keyframe test, local-map update.  One sweep in flight per context:
 * a second submit, a step, or any call that rewrites the feature slot in use, returns LMSF_ERR_STATE until wait.
 * lmsf_tracker_step(x) == lmsf_tracker_submit(x) + lmsf_tracker_wait(). */
int lmsf_tracker_submit(lmsf_ctx* c, const float* xyzi, int n, double stamp, const double delta[7]);
int lmsf_tracker_submit_dev(lmsf_ctx* c, const float* d_xyzi, int n, double stamp, const double delta[7]);
int lmsf_tracker_submit_ticket(lmsf_ctx* c, int64_t ticket, double stamp, const double delta[7]);
int lmsf_tracker_wait(lmsf_ctx* c, double delta_out[7], double pose_out[7], lmsf_track_stats* st);
/* Same, fed with features instead of a raw sweep — the exact argument of
 * LidarTrackerLocalMap::Solve. */
int lmsf_tracker_step_features(lmsf_ctx* c, const float* edge_xyzi, int n_e, const float* surf_xyzi,
                               int n_s, double stamp, double delta[7], double pose_out[7],
                               lmsf_track_stats* st);
int lmsf_tracker_reset(lmsf_ctx* c);
/* = LidarTrackerLocalMap::RegistrationLocalMap (:168-177) as used for an auxiliary
 * LiDAR against the primary's map (System/ML_System.hpp:304): extract features
 * from xyzi and register them against this tracker's current local map. */
int lmsf_tracker_register_aux(lmsf_ctx* c, const float* xyzi, int n, double pose[7],
                              lmsf_reg_stats* st);
/* Multi-LiDAR rigs with one LiDAR per GPU (System/ML_System.hpp:296-310 on several devices): the auxiliary LiDAR's GPU
 * extracts (lmsf_extract_features_to_dev: edges first, then surfs, into a DEVICE buffer of `cap` points — the caller
 * ships that buffer to the primary's GPU, peer copy or NCCL), the primary's context registers the shipped features
 * against its local map (lmsf_tracker_register_aux_features_dev = lmsf_tracker_register_aux without the extraction).
 * The caller orders its transfer against the context's stream (lmsf_stream). */
int lmsf_extract_features_to_dev(lmsf_ctx* c, const float* xyzi, int n, void* d_feat_out, int cap, int* n_edge,
                                 int* n_surf);
int lmsf_tracker_register_aux_features_dev(lmsf_ctx* c, const void* d_feat, int n_edge, int n_surf, double pose[7],
                                           lmsf_reg_stats* st);
/* = LidarTrackerLocalMap::GetLocalMap (:184-192) */
int lmsf_get_map(lmsf_ctx* c, int kind, float* xyzi, int cap, int* n);

/* ---- loop-closure descriptors (ScanContext), device resident ------- */
/* A descriptor is 20 x 60 fp32, row-major (ring, sector); a ring key is 20 fp32.  The database lives in the
 * context's HBM (keys[n][20], descs[n][1200]); ids are insertion order, as polarcontexts_sc_ /
 * polarcontext_ringkeys_vec_ (LoopDetection/SceneRecognitionScanContext.hpp:61-72). */
#define LMSF_SC_RINGS 20
#define LMSF_SC_SECTORS 60
#define LMSF_SC_CELLS 1200
#define LMSF_SC_CANDIDATES 10 /* NUM_CANDIDATES_FROM_TREE_ */
/* one ring-key candidate of one query, as exchanged between database shards (24 bytes) */
typedef struct {
  double sc_dist;  /* DistanceBtnScanContext(query, candidate) */
  float key_dist;  /* squared ring-key distance (nanoflann L2_Adaptor, fp32) */
  int32_t id;      /* global keyframe id, -1 = empty slot */
  int32_t shift;   /* column shift of the best alignment */
  int32_t pad;
} lmsf_sc_cand;
/* = ScanContext::MakeScanContext + MakeRingkeyFromScanContext
 * (Algorithm/PointClouds/processing/GlobalDescriptor/scanContext/Scancontext.hpp:59-104, :112-126) */
int lmsf_sc_make(lmsf_ctx* c, const float* xyzi, int n, float* desc1200, float* key20);
/* = ScanContext::DistanceBtnScanContext (Scancontext.hpp:133-172) for n_pairs explicit pairs
 * (a = _sc1, the query; b = _sc2, the candidate that is shifted) */
int lmsf_sc_distance(lmsf_ctx* c, const float* desc_a, const float* desc_b, int n_pairs, double* dist,
                     int32_t* shift);
int lmsf_scdb_reserve(lmsf_ctx* c, int capacity);
int lmsf_scdb_clear(lmsf_ctx* c);
int lmsf_scdb_size(lmsf_ctx* c, int* n);
/* append n precomputed descriptors (SceneRecognitionScanContext::Load :165-231) */
int lmsf_scdb_add(lmsf_ctx* c, const float* descs, const float* keys, int n);
/* = the descriptor part of AddKeyFramePoints (:61-72): describe the cloud on the device and append it */
int lmsf_scdb_add_cloud(lmsf_ctx* c, const float* xyzi, int n, int* id_out);
int lmsf_scdb_get(lmsf_ctx* c, int id, float* desc1200, float* key20);
/* test hook for the ring-key tree search of descFindSimilar (:267-279): exact 10-NN of every query among
 * keys[0, limit), ascending by (distance, id); unfilled slots hold id -1 and +inf */
int lmsf_scdb_knn(lmsf_ctx* c, const float* q_keys, int nq, int limit, int32_t* idx10, float* d10);
/* = descFindSimilar (:260-333) for nq queries against ids [0, limit): loop id (-1 = "Not loop"), the
 * smallest SC distance among the ten ring-key candidates and its column shift (yaw = shift * 6 deg).
 * limit = size of polarcontext_ringkeys_to_search_ (see lmsf_sc_tree_limit). */
int lmsf_scdb_search(lmsf_ctx* c, const float* q_keys, const float* q_descs, int nq, int limit, double thresh,
                     int32_t* loop_id, double* loop_dist, int32_t* loop_shift);
/* Two-round variant of the sharded search (the dominant per-shard cost of the one-round exchange is scoring 10
 * candidates per query on EVERY rank): round 1 — lmsf_scdb_keys_shard_dev: the shard's ring-key top-10 as unscored
 * lmsf_sc_cand records (sc_dist = 1e7), all-gathered by the caller; round 2 — lmsf_scdb_score_owned_dev: every rank
 * derives the same global ring-key top-10 from the n_ranks x nq x 10 records and computes the ScanContext distance of
 * the candidates it owns (global ids [id_base, id_base + n_local)), writing the others as empty slots; the scored
 * blocks are all-gathered again and lmsf_scdb_pick_dev selects exactly as in the one-round scheme.  Device pointers,
 * work enqueued on the context's stream. */
int lmsf_scdb_keys_shard_dev(lmsf_ctx* c, const float* d_q_keys, int nq, int limit_local, int id_base, void* d_cand);
int lmsf_scdb_score_owned_dev(lmsf_ctx* c, const void* d_cand_all, int n_ranks, int nq, const float* d_q_descs,
                              int id_base, int n_local, void* d_scored);
/* Sharded search (one database shard per GPU): ring-key top-10 and SC distances of nq DEVICE queries
 * against local ids [0, limit_local); d_cand receives nq x 10 lmsf_sc_cand with id = local id + id_base.
 * The caller all-gathers the candidate blocks of all ranks (NCCL) and calls lmsf_scdb_pick_dev on
 * [n_ranks][nq][10] candidates: global ring-key top-10 by (key_dist, id), then the first strict minimum of
 * sc_dist in that order, then the threshold — the selection of descFindSimilar :296-323.  Both calls only
 * enqueue work on the context's stream. */
int lmsf_scdb_search_shard_dev(lmsf_ctx* c, const float* d_q_keys, const float* d_q_descs, int nq,
                               int limit_local, int id_base, void* d_cand);
int lmsf_scdb_pick_dev(lmsf_ctx* c, const void* d_cand_all, int n_ranks, int nq, double thresh,
                       int32_t* d_loop_id, double* d_loop_dist, int32_t* d_loop_shift);
/* size of the searched prefix after n_keyframes AddKeyFramePoints calls: the tree is rebuilt over
 * [0, size - 50) whenever (size - 1) % 10 == 0 and size > 50 (:74-92); 0 = no tree yet */
int lmsf_sc_tree_limit(int n_keyframes);

/* ---- multi-LiDAR extrinsic initialisation (hand-eye), host side ------ */
/* = Algorithm::HandEyeCalibrationBase (Algorithm/calibration/handeye_calibration_base.hpp:36-246) as driven by
 * MultiLidarSystem::process() in calibration status 0 (System/ML_System.hpp:243-283).  Poses are per-sweep motion
 * increments {qx,qy,qz,qw,tx,ty,tz} of the primary and the auxiliary tracker (lmsf_tracker_step's delta output).
 * add_pose = AddPose (:73-108, with checkScrewMotion :208-243); *enough = 1 once three pairs are stored.
 * calibrate = CalibExRotation (:115-153) + CalibExTranslation (:155-190) + GetCalibResult (:193-202): *ok = 1 and
 * extrinsic = pose of the auxiliary LiDAR in the primary's frame when the second smallest singular value exceeds
 * 0.25; singular_values (optional) are returned in descending order.  Status 1, the per-sweep refinement of the
 * extrinsic against the primary's local map (:286-323), is lmsf_tracker_register_aux.  Pure host arithmetic. */
typedef struct lmsf_handeye lmsf_handeye;
int lmsf_handeye_create(lmsf_handeye** out);
void lmsf_handeye_destroy(lmsf_handeye* h);
int lmsf_handeye_add_pose(lmsf_handeye* h, const double delta_primary[7], const double delta_sub[7], int* enough);
int lmsf_handeye_calibrate(lmsf_handeye* h, double extrinsic[7], double singular_values[4], int* ok);
int lmsf_handeye_size(lmsf_handeye* h, int* n);

/* ---- device memory helpers for callers that keep sweeps resident ---- */
int lmsf_dev_alloc(lmsf_ctx* c, int64_t bytes, void** d_ptr);
int lmsf_dev_free(lmsf_ctx* c, void* d_ptr);
int lmsf_dev_upload(lmsf_ctx* c, void* d_dst, const void* h_src, int64_t bytes);

/* ---- per-stage device timing (CUDA events on the context's stream) -- */
#define LMSF_STAGE_EXTRACT 0
#define LMSF_STAGE_MATCH 1    /* k_knn: exact 5-NN over the local-map grid (dominant kernel) */
#define LMSF_STAGE_SOLVE 2    /* k_lm_eval: LM candidate evaluation + reduction + 6x6 step */
#define LMSF_STAGE_MAP 3      /* local-map update + index rebuild */
#define LMSF_STAGE_VOXEL 4
#define LMSF_STAGE_ASSOC 5    /* k_assoc + query sort */
#define LMSF_STAGE_FIT 6      /* k_fit: PCA/plane fit + residual/Jacobian + reduction + first 6x6 step */
#define LMSF_N_STAGES 7
/* enable = 1 records events around every stage (adds host syncs at query time only) */
int lmsf_profile_enable(lmsf_ctx* c, int enable);
/* accumulated milliseconds and launch counts per stage since the last reset;
 * bytes = algorithmic bytes of the k_knn launches (16*(F+M)+216 each) */
int lmsf_profile_read(lmsf_ctx* c, double ms[LMSF_N_STAGES], int64_t launches[LMSF_N_STAGES],
                      double* match_alg_bytes, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LMSF_B200_H_ */
