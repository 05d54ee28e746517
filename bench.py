#!/usr/bin/env python
"""bench.py — scan-to-map registrations/sec on a synthetic HDL-64 sequence (BASELINE.json configs[1]).

One step = one sweep through the hot path: LOAM feature extraction -> Huber-LM scan-to-map
registration against the sliding-window local map -> keyframe test -> (on keyframes) local-map
update + kNN index rebuild, i.e. lmsf_tracker_step (LidarTrackerLocalMap::Solve fed by
LOAMFeatureProcessorBase::Process in the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework, on B200
    python bench.py --impl reference [...]                          # the reference's CPU path (oracle port)

N > 1: one process per GPU (torchrun), each rank tracks its own independent sequence (sequence id =
rank, BASELINE.json configs[3]); no data-path collective; value = all ranks' sweeps / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "scan-to-map registrations/sec (HDL-64 synthetic)"   # renamed for --sensor vlp16 in main()
UNIT = "scans/s"
SENSOR = "hdl64"   # --sensor vlp16 switches to BASELINE.json configs[0] (not the headline line)
SENSORS = {"hdl64": (64, "HDL-64 synthetic 64x2048 sweep (~131k pts)"), "vlp16": (16, "VLP-16 synthetic 16x1800 sweep (~29k pts)")}
# Untimed initialisation sweeps in front of every pass (ours and the reference arm alike): a tracker that has just been
# reset is not the workload — its sliding window fills over the first ten keyframes (~20 sweeps at 1.5 m/s) and the
# Huber-LM outer-iteration budget decays 9, 8, ..., 2 over the first eight solves (ceres_edgeSurfFeatureRegistration.hpp:
# 100-101).  After SETTLE sweeps the window is full and the budget sits at its floor, whatever --warmup is.
SETTLE = 24
FLUSH_BYTES = 192 << 20   # written between timed steps: 1.5x the 126 MB L2
STATE_D2H_BYTES = 880 + 8  # sizeof(SolveState) + the two feature counts, read back once per sweep
MAP_LEAF = {}              # --map-leaf E,S: voxel-filtered local map (lmsf_params.map_leaf_edge / map_leaf_surf)
MAP_DESC = "raw 10-keyframe sliding-window map"


def _gen(args):
    name, k, seq = args
    pkg = entry.load_package()
    return pkg.synth.make_sweep(pkg.synth.sensor_by_name(name), k, seq=seq)


def make_sequence(n_sweeps: int, seq: int):
    """Seeded synthetic sweeps 0..n_sweeps-1 of sequence `seq` (generated on the host cores, before CUDA init)."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp

    jobs = [(SENSOR, k, seq) for k in range(n_sweeps)]
    workers = max(1, min(len(jobs), (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if workers == 1:
        return [_gen(j) for j in jobs]
    with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork")) as ex:
        return list(ex.map(_gen, jobs))


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region, polled in-process through NVML (no fork:
    spawning nvidia-smi from a process with CUDA loaded stalls the launching thread for tens of ms)."""

    def __init__(self, index: int, period: float = 0.005):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self.active = threading.Event()
        self.stop_flag = threading.Event()
        self.err = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            # poll all the time (NVML's first calls take tens of ms and stall CUDA launches from other
            # threads meanwhile: that must happen during warm-up), record only inside the timed region
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                if self.active.is_set():
                    self.samples.append((sm, rs))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
            self.stop_flag.wait(self.period)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: " + str(self.err)]}
        nv = self.nv
        sm = sorted(s[0] for s in self.samples)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = [n for n, b in bits.items() if any(s[1] & b for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_ours(args):
    import torch

    rank, world, local = dist_env()
    n_gpus = args.gpus
    use_dist = world > 1
    W, K = args.warmup, args.steps
    # independent sequences are dealt round-robin to ranks (shard.assign): with one sequence per GPU rank r tracks sequence r
    seq_id = rank + int(os.environ.get("LMSF_BENCH_SEQ", "0"))   # env: single-GPU experiments on another sequence
    sweeps = make_sequence(SETTLE + W + K + 2, seq=seq_id)     # before any CUDA call (fork-safe); +1: the last step prefetches
    torch.cuda.set_device(local)
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = entry.load_package()
    ctx = pkg.context(local, n_scans=SENSORS[SENSOR][0], max_points=1 << 18, **MAP_LEAF)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2
    n_pts = [int(s.shape[0]) for s in sweeps]

    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    poses = {}

    def run_sequence(step_fn, first_fn, profile, K=K, flush_l2=True):
        """W untimed + K timed steps; returns (device ms for K steps, per-step wall ms, launches, prof, ...).
        first_fn(k) prefetches sweep k (the ticket the first step consumes); every step consumes the ticket of its
        sweep and prefetches the next one."""
        ctx.tracker_reset()
        state = {"ticket": first_fn(0)}
        for k in range(0, SETTLE + 1):          # initialisation (see SETTLE), then the W warm-up steps
            poses[k] = step_fn(k, 0.1 * k, state)[0]
        for k in range(SETTLE + 1, SETTLE + W + 1):
            poses[k] = step_fn(k, 0.1 * k, state)[0]
        barrier()
        ctx.profile_enable(profile)
        ctx.profile_read(reset=True)
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = []
        stats = []
        flush_host = 0.0
        marks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        sampler.active.set()
        t_wall0 = time.perf_counter()
        e0.record(stream)
        for i, k in enumerate(range(SETTLE + W + 1, SETTLE + W + K + 1)):
            tf = time.perf_counter()
            if flush_l2:
                with torch.cuda.stream(stream):
                    flush.zero_()                   # L2 flush between timed steps, inside the timed region
            t0 = time.perf_counter()
            flush_host += t0 - tf
            marks[i][0].record(stream)              # behind the flush: the step's own work starts here
            pose, st = step_fn(k, 0.1 * k, state)
            marks[i][1].record(stream)
            poses[k] = pose
            stats.append(st)
            per.append((time.perf_counter() - t0) * 1e3)
        e1.record(stream)
        barrier()
        sampler.active.clear()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        step_ms = sum(a.elapsed_time(b) for a, b in marks)   # the same K steps without the K memsets
        launches = ctx.launch_count() - l0
        prof = ctx.profile_read(reset=True) if profile else None
        ctx.profile_enable(False)
        ctx.tracker_prefetch_cancel(state["ticket"])
        return dev_ms, wall_ms, per, launches, prof, stats, flush_host * 1e3 / K, step_ms

    # ---- resident pass: sweeps already in HBM when the timed region starts
    d_ptrs = [ctx.dev_upload_new(s) for s in sweeps]
    # dress rehearsal (discarded): the very first timed loop of a process pays one-off lazy initialisation
    # (CUDA event pools, NVML) of ~50-90 ms inside its first iteration
    # Every step first hands the NEXT sweep to the front end (lmsf_tracker_prefetch*: upload + feature extraction
    # on the context's front-end stream), then registers the current one — the reference's own two-thread pipeline
    # (sensor thread: Process; estimate_thread_: Solve).  Each timed step therefore still contains exactly one
    # upload (e2e pass), one feature extraction and one registration.
    # The step is issued in its two halves (lmsf_tracker_submit / lmsf_tracker_wait == lmsf_tracker_step) so that the
    # host enqueues the next sweep's front end while the GPU registers the current one.
    def first_dev(k):
        return ctx.tracker_prefetch_dev(d_ptrs[k], n_pts[k])

    def first_host(k):
        return ctx.tracker_prefetch(sweeps[k])

    def step_dev(k, t, state):
        ctx.tracker_submit_ticket(state["ticket"], t)
        state["ticket"] = ctx.tracker_prefetch_dev(d_ptrs[k + 1], n_pts[k + 1])
        r = ctx.tracker_wait()
        return r[0], r[2]

    def step_host(k, t, state):
        ctx.tracker_submit_ticket(state["ticket"], t)
        state["ticket"] = ctx.tracker_prefetch(sweeps[k + 1])     # H2D of the next sweep, inside the step
        r = ctx.tracker_wait()                                       # D2H of this sweep's pose
        return r[0], r[2]

    run_sequence(step_dev, first_dev, False, K=2)
    dev_ms, wall_ms, per, launches, _, stats, fl1, step_ms = run_sequence(step_dev, first_dev, False)
    gpu_poses = dict(poses)
    # ---- end-to-end pass: host buffers through lmsf_tracker_prefetch + lmsf_tracker_submit_ticket + lmsf_tracker_wait
    # (H2D of a sweep + D2H of the pose inside every step)
    e_dev_ms, e_wall_ms, e_per, _, _, _, fl2, _ = run_sequence(step_host, first_host, False)
    # ---- the same resident pass without the L2 flush (reported beside the flushed figure, never as `value`)
    nf_dev_ms, _, _, _, _, _, _, _ = run_sequence(step_dev, first_dev, False, flush_l2=False)
    # ---- instrumented pass (not used for value): per-stage CUDA-event times on the context's streams
    p_dev_ms, _, _, _, prof, _, fl3, _ = run_sequence(step_dev, first_dev, True)
    sampler.stop_flag.set()
    sampler.join()

    dev_ms_max, e2e_ms_max = pkg.shard.max_over_ranks([dev_ms, e_wall_ms], device=f"cuda:{local}")
    per_rank = [dev_ms / K]
    if use_dist:   # every rank's own device time per step (evidence for the weak-scaling figure; rank r = sequence r)
        t = torch.tensor([dev_ms / K, float(sum(1 for s in stats if s["keyframe"]))], dtype=torch.float64, device=f"cuda:{local}")
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[round(float(x[0]), 4), int(x[1])] for x in allt]

    if rank == 0:
        ms, ln, alg_bytes = prof
        peak, peak_src = peaks()
        match_launches = max(1, ln["match"])
        match_ms = ms["match"] / match_launches
        bytes_per_launch = alg_bytes / match_launches
        achieved = bytes_per_launch / (match_ms * 1e-3) / 1e9 if match_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "match_traffic_bytes.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = tj.get("source", "ncu --set full capture (profiles/), not measured in this run")
            except Exception:
                traffic = None
        kf = sum(1 for s in stats if s["keyframe"])
        line = {
            "metric": METRIC, "value": n_gpus * K / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": n_gpus,
            "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": SENSORS[SENSOR][1] + " scan-to-map edge/surf registration, "
                                   "Huber-LM solver, " + MAP_DESC + ", one sequence per GPU",
                       "points_per_sweep": int(np.mean(n_pts)), "sequences": n_gpus,
                       "l2": f"{FLUSH_BYTES >> 20} MiB memset between timed steps (L2 flush: 1.5x the 126 MB L2), inside the timed region",
                       "timing": "CUDA events on the context stream around the K steps, max over ranks",
                       "initialisation": f"{SETTLE} untimed sweeps before the warm-up: full 10-keyframe window, LM budget at its floor",
                       "pipeline": "front end (extraction of sweep k+1) overlaps the registration of sweep k",
                       "p50_ms_per_scan": float(np.median(per)), "p90_ms_per_scan": float(np.percentile(per, 90)),
                       "max_ms_per_scan": float(np.max(per)), "wall_ms_per_step": wall_ms / K,
                       "flush_host_ms_per_step": [fl1, fl2, fl3], "instrumented_pass_ms_per_step": p_dev_ms / K,
                       "ms_per_step_between_flushes": step_ms / K,
                       "ms_per_step_no_flush_pass": nf_dev_ms / K,
                       "keyframes_in_timed_region": kf, "per_rank_ms_per_step_and_keyframes": per_rank,
                       "features_per_sweep": int(np.mean([s["n_edge"] + s["n_surf"] for s in stats])),
                       "map_points_end": int(stats[-1]["map_edge"] + stats[-1]["map_surf"]),
                       "stage_ms_per_step": {k: v / K for k, v in ms.items()},
                       "stage_launches_per_step": {k: v / K for k, v in ln.items()}},
            "e2e": {"value": n_gpus * K / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(np.mean(n_pts)) * 16, "d2h_bytes_per_step": STATE_D2H_BYTES,
                    "p50_ms_per_scan": float(np.median(e_per)),
                    "api": "lmsf_tracker_submit_ticket(sweep k) + lmsf_tracker_prefetch(host sweep k+1: H2D + extraction) + "
                           "lmsf_tracker_wait(pose out), wall clock around K steps, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_knn + k_knn_sparse (exact 5-NN of every scan feature over the local-map grid: one "
                                                    "correspondence pass = the per-thread search and the warp-per-query "
                                                    "search of the cases it defers; timed together)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": match_ms,
                         "launches_per_step": ln["match"] / K,
                         "share_of_step": ms["match"] / max(1e-9, p_dev_ms)},
        }
        if n_gpus == 1 and not args.no_cpu:
            line["cpu_baseline"], line["parity"] = cpu_baseline(sweeps, W, gpu_poses)
        print(json.dumps(line), flush=True)
    for p in d_ptrs:
        ctx.dev_free(p)
    ctx.close()
    if use_dist:
        dist.destroy_process_group()


def _pose_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    dt = float(np.linalg.norm(a[4:] - b[4:]))
    qa, qb = a[:4] / np.linalg.norm(a[:4]), b[:4] / np.linalg.norm(b[:4])
    return dt, 2.0 * float(np.arccos(min(1.0, abs(float(np.dot(qa, qb))))))


def cpu_baseline(sweeps, W, gpu_poses, sample=4):
    """The oracle port of the reference's CPU path on a bounded sample of the same workload: the
    tracker is brought to the steady state with all host threads (untimed), then `sample` sweeps are
    timed with ONE thread — the reference runs one thread per LiDAR (System/ML_System.hpp:137,248).
    Every pose the oracle produces on the way (initialisation, warm-up and the timed sample) is compared with the
    pose the CUDA path gave for the same sweep in the resident pass -> the line's `parity` object."""
    lib = entry.load_oracle()
    threads = os.cpu_count() or 1
    o = lib.context(0, n_scans=SENSORS[SENSOR][0], oracle_knn_mode=0, oracle_threads=threads, **MAP_LEAF)
    warm = min(SETTLE + W, len(sweeps) - sample - 1)
    max_dt = max_dr = 0.0
    kf_equal = True
    n_cmp = 0

    def compare(k, r):
        nonlocal max_dt, max_dr, n_cmp
        if k in gpu_poses:
            dt, dr = _pose_err(gpu_poses[k], r[0])
            max_dt, max_dr, n_cmp = max(max_dt, dt), max(max_dr, dr), n_cmp + 1

    for k in range(0, warm + 1):
        compare(k, o.tracker_step(sweeps[k], 0.1 * k))
    lib.fn("set_threads")(o._h, 1)
    t0 = time.perf_counter()
    res = []
    for k in range(warm + 1, warm + 1 + sample):
        res.append((k, o.tracker_step(sweeps[k], 0.1 * k)))
    dt = time.perf_counter() - t0
    for k, r in res:
        compare(k, r)
    o.close()
    base = {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample} consecutive {SENSOR} sweeps after a {warm}-sweep warm-up of the same sequence, "
                      f"oracle tracker (kd-tree kNN), 1 thread; host has {threads} cores",
            "seconds": dt}
    parity = {"max_dt_m": max_dt, "max_dr_rad": max_dr, "sweeps_compared": n_cmp, "tol_m": 1e-4, "tol_rad": 1e-5,
              "ok": bool(n_cmp > 0 and max_dt < 1e-4 and max_dr < 1e-5),
              "what": "CUDA tracker poses (resident pass) vs oracle tracker poses on the same sweeps: sweeps 0.."
                      f"{warm + sample} of the bench sequence, including {sample} sweeps of the timed region"}
    return base, parity


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be
    built here (no PCL/Eigen/Ceres; its tree lacks Map/), so this times the oracle port with all the host
    threads it can use.  --gpus N: N independent sequences (seeds 0..N-1, the workload of the N-GPU arm), tracked
    concurrently by N oracle trackers that share the host cores (cores // N OpenMP threads each — BASELINE.md: "one
    sequence per core" for the batched configuration); value = N*K sweeps / wall time of the slowest tracker."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    W, K = args.warmup, args.steps
    n_seq = max(1, args.gpus)
    seqs = [make_sequence(SETTLE + W + K + 1, seq=s) for s in range(n_seq)]
    lib = entry.load_oracle()
    cores = os.cpu_count() or 1
    threads = max(1, cores // n_seq)
    ctxs = [lib.context(0, n_scans=SENSORS[SENSOR][0], oracle_knn_mode=0, oracle_threads=threads, **MAP_LEAF)
            for _ in range(n_seq)]
    per = [[] for _ in range(n_seq)]
    go = threading.Barrier(n_seq + 1)
    done = threading.Barrier(n_seq + 1)

    def track(i):
        o, sweeps = ctxs[i], seqs[i]
        for k in range(0, SETTLE + W + 1):
            o.tracker_step(sweeps[k], 0.1 * k)
        go.wait()
        for k in range(SETTLE + W + 1, SETTLE + W + K + 1):
            t1 = time.perf_counter()
            o.tracker_step(sweeps[k], 0.1 * k)
            per[i].append((time.perf_counter() - t1) * 1e3)
        done.wait()

    ths = [threading.Thread(target=track, args=(i,), daemon=True) for i in range(n_seq)]
    for t in ths:
        t.start()
    go.wait()                      # every tracker has finished its untimed sweeps
    t0 = time.perf_counter()
    done.wait()                    # the slowest tracker has finished its K timed sweeps
    dt = time.perf_counter() - t0
    for t in ths:
        t.join()
    for o in ctxs:
        o.close()
    v = n_seq * K / dt
    allper = [x for p in per for x in p]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": SENSORS[SENSOR][1] + " scan-to-map edge/surf registration, "
                               "Huber-LM solver, " + MAP_DESC + ", one sequence per GPU",
                   "sequences": n_seq, "p50_ms_per_scan": float(np.median(allper))},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads * n_seq, "kind": "port",
                         "sample": f"{K} consecutive sweeps of each of {n_seq} sequence(s) after {SETTLE} initialisation + "
                                   f"{W} warm-up sweeps, oracle tracker, {n_seq} tracker(s) x {threads} OpenMP threads "
                                   f"on {cores} host cores"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# --workload loopdb: BASELINE.json configs[4] — loop-closure descriptor database search over 100 000 keyframes,
# database sharded over the ranks, per-rank candidates merged with one NCCL all_gather (shard.loop_search_sharded).
# Not the driver's default line; run explicitly:  python bench.py --workload loopdb [--gpus N]
LOOP_METRIC = "loop-closure descriptor searches/sec (ScanContext, 100k-keyframe database)"
LOOP_DB = int(os.environ.get("LMSF_LOOP_DB", "100000"))   # env: profiling one shard's kernels on one GPU
LOOP_NQ = 1024


def make_loop_data(rank_lo, rank_hi, seed=20260500):
    """Seeded synthetic database shard [lo, hi) + the replicated query batch.  Descriptors are derived from real
    ScanContext descriptors of 64 synthetic VLP-16 sweeps (made on the GPU) by scaling, noise and column shifts."""
    pkg = entry.load_package()
    rng = np.random.default_rng(seed)
    base_desc = make_loop_data.base
    nb = len(base_desc)
    src = rng.integers(0, nb, size=LOOP_DB)
    scale = rng.uniform(0.6, 1.4, size=LOOP_DB).astype(np.float32)
    shift = rng.integers(0, 60, size=LOOP_DB)
    q_ids = rng.integers(0, LOOP_DB - 50, size=LOOP_NQ)
    q_shift = rng.integers(0, 60, size=LOOP_NQ)

    def desc_of(ids):
        out = np.empty((len(ids), 20, 60), np.float32)
        for j, i in enumerate(ids):
            r = np.random.default_rng(seed + 1 + int(i))
            d = np.roll(base_desc[src[i]] * scale[i], int(shift[i]), axis=1)
            out[j] = np.where(d > 0, d + r.normal(0, 0.15, size=d.shape), 0).astype(np.float32)
        return out

    def keys_of(d):
        return (np.cumsum(d.astype(np.float64), axis=2)[:, :, -1] / 60).astype(np.float32)

    shard_desc = desc_of(range(rank_lo, rank_hi))
    qd = desc_of(q_ids)
    qd = np.stack([np.roll(qd[j], int(q_shift[j]), axis=1) for j in range(LOOP_NQ)])
    r = np.random.default_rng(seed + 7)
    qd = np.where(qd > 0, qd + r.normal(0, 0.03, size=qd.shape), 0).astype(np.float32)
    return shard_desc, keys_of(shard_desc), qd, keys_of(qd), q_ids, q_shift


def loopdb_cpu_baseline(keys, descs, qk, qd, limit, gpu_ids, nqs=64):
    """descFindSimilar on the host cores for a bounded sample of the query batch: the ring-key stage is the
    REFERENCE'S OWN vendored nanoflann k-d tree (oracle/_ref/libref_nanoflann.so: KDTreeVectorOfVectorsAdaptor, leaf 10,
    SceneRecognitionScanContext.hpp:86-91, 272-279; built outside the timed region as the reference rebuilds it only
    every tenth keyframe), the ScanContext distance of the ten candidates and the selection are the oracle port."""
    import ctypes as C

    lib = C.CDLL(entry.ORACLE_LIB)
    f32p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
    kk = np.ascontiguousarray(keys[:limit])
    dd = np.ascontiguousarray(descs.reshape(-1, 1200))
    q_k = np.ascontiguousarray(qk[:nqs])
    q_d = np.ascontiguousarray(qd[:nqs].reshape(nqs, 1200))
    o_id = np.full(nqs, -1, np.int32)
    if os.path.exists(entry.REF_NANOFLANN_LIB):
        ref = C.CDLL(entry.REF_NANOFLANN_LIB)
        ref.ref_ringkey_tree_create.restype = C.c_void_p
        ref.ref_ringkey_tree_create.argtypes = [f32p, C.c_int]
        ref.ref_ringkey_tree_knn10.argtypes = [C.c_void_p, f32p, C.c_int, i32p, f32p]
        ref.ref_ringkey_tree_destroy.argtypes = [C.c_void_p]
        tree = ref.ref_ringkey_tree_create(kk.ctypes.data_as(f32p), C.c_int(limit))
        idx = np.empty((nqs, 10), np.int32)
        d10 = np.empty((nqs, 10), np.float32)
        dist, shift = C.c_double(0), C.c_int(0)
        t0 = time.perf_counter()
        ref.ref_ringkey_tree_knn10(tree, q_k.ctypes.data_as(f32p), nqs, idx.ctypes.data_as(i32p), d10.ctypes.data_as(f32p))
        for q in range(nqs):
            best, best_id = 1e300, -1
            for c in idx[q]:
                if c < 0:
                    continue
                lib.lmsf_oracle_sc_distance(q_d[q].ctypes.data_as(f32p), dd[c].ctypes.data_as(f32p), C.byref(dist),
                                            C.byref(shift))
                if dist.value < best:
                    best, best_id = dist.value, int(c)
            o_id[q] = best_id if best < 0.2 else -1
        dt = time.perf_counter() - t0
        ref.ref_ringkey_tree_destroy(tree)
        how = "ring-key 10-NN by the reference's own nanoflann k-d tree (kind 'reference' for that stage), SC distance + selection by the oracle port"
    else:
        o_d, o_s = np.empty(nqs, np.float64), np.empty(nqs, np.int32)
        t0 = time.perf_counter()
        lib.lmsf_oracle_sc_search(kk.ctypes.data_as(f32p), dd.ctypes.data_as(f32p), C.c_int(limit), q_k.ctypes.data_as(f32p),
                                  q_d.ctypes.data_as(f32p), C.c_int(nqs), C.c_double(0.2), o_id.ctypes.data_as(i32p),
                                  o_d.ctypes.data_as(f64p), o_s.ctypes.data_as(i32p))
        dt = time.perf_counter() - t0
        how = "oracle brute-force ring-key scan + SC distance (oracle/_ref/libref_nanoflann.so absent)"
    return {"value": nqs / dt, "unit": "queries/s", "cores": 1, "kind": "port",
            "sample": f"{nqs} of the {LOOP_NQ} queries against the same database, {how}, 1 thread",
            "agrees_with_gpu": bool(np.array_equal(o_id, gpu_ids[:nqs]))}


def run_loopdb(args):
    import torch

    rank, world, local = dist_env()
    use_dist = world > 1
    W, K = args.warmup, args.steps
    torch.cuda.set_device(local)
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    ctx = pkg.context(local, n_scans=16, max_points=1 << 16, max_map_points=1 << 17)
    sensor = pkg.synth.vlp16()
    make_loop_data.base = np.stack([ctx.sc_make(pkg.synth.make_sweep(sensor, 3 * k))[0] for k in range(64)])
    lo, hi = pkg.shard.shard_bounds(LOOP_DB, world, rank)
    descs, keys, qd, qk, q_ids, q_shift = make_loop_data(lo, hi)
    ctx.scdb_reserve(hi - lo)
    for a in range(0, hi - lo, 8192):
        ctx.scdb_add(descs[a:a + 8192], keys[a:a + 8192])
    limit = pkg.library().fn("sc_tree_limit")(LOOP_DB)   # the searched prefix after LOOP_DB keyframes
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)
    h_qk = torch.from_numpy(qk).pin_memory()
    h_qd = torch.from_numpy(qd.reshape(LOOP_NQ, 1200)).pin_memory()
    d_qk, d_qd = h_qk.to(dev), h_qd.to(dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = None

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    if args.loop_exchange == 1:
        def search(c, a, b, lim, l, h, th):
            return pkg.shard.loop_search_sharded(c, a, b, lim, l, h, th)
    else:
        searcher = pkg.shard.ShardedLoopSearch(ctx, LOOP_NQ, limit, lo, hi, 0.2, device=dev)

        def search(c, a, b, lim, l, h, th):
            return searcher.search(a, b)
    host_ms = [0.0, 0]

    def step_resident():
        t = time.perf_counter()
        r = search(ctx, d_qk, d_qd, limit, lo, hi, 0.2)
        host_ms[0] += (time.perf_counter() - t) * 1e3
        host_ms[1] += 1
        return r

    def step_e2e():
        with torch.cuda.stream(stream):
            a = h_qk.to(dev, non_blocking=True)
            b = h_qd.to(dev, non_blocking=True)
        ids, dd, sh = search(ctx, a, b, limit, lo, hi, 0.2)
        with torch.cuda.stream(stream):
            out = (ids.cpu(), dd.cpu(), sh.cpu())
        stream.synchronize()
        return out

    def timed(fn):
        for _ in range(W):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active.set()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(K):
            with torch.cuda.stream(stream):
                flush.zero_()
            r = fn()
        e1.record(stream)
        barrier()
        sampler.active.clear()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, r

    timed(step_resident)   # dress rehearsal (discarded): clocks, allocator pools and NCCL channels settle
    l0 = ctx.launch_count()
    host_ms[0], host_ms[1] = 0.0, 0
    dev_ms, _, res = timed(step_resident)
    launches = ctx.launch_count() - l0
    _, e_wall_ms, res_e = timed(step_e2e)
    sampler.stop_flag.set()
    sampler.join()
    ids = res[0].cpu().numpy()
    hit = float(np.mean(ids == q_ids))
    assert np.array_equal(ids, res_e[0].numpy())
    dev_ms_max, e2e_ms_max = pkg.shard.max_over_ranks([dev_ms, e_wall_ms], device=f"cuda:{local}")
    if rank == 0:
        peak, peak_src = peaks()
        line = {
            "metric": LOOP_METRIC, "value": LOOP_NQ * K / (dev_ms_max * 1e-3), "unit": "queries/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ScanContext loop search: 100 000-keyframe database sharded over the ranks, "
                                   f"{LOOP_NQ} replicated queries per step, exact ring-key top-10 per shard, " +
                                   ("SC distance of the 10 local candidates on every rank, one all_gather of 24-byte "
                                    "candidates" if args.loop_exchange == 1 else
                                    "all_gather of the unscored candidates, SC distance of the global top-10 by their "
                                    "owners only, second all_gather") + ", identical selection on every rank",
                       "exchange_rounds": args.loop_exchange,
                       "host_enqueue_ms_per_step": host_ms[0] / max(1, host_ms[1]),
                       "database": LOOP_DB, "searched_prefix": int(limit), "queries_per_step": LOOP_NQ,
                       "shard": [int(lo), int(hi)], "recall_of_planted_revisits": hit,
                       "l2": f"{FLUSH_BYTES >> 20} MiB memset between timed steps (L2 flush), inside the timed region",
                       "bound_note": "the ring-key scan is fp32-ALU bound (60 non-fused flops per key pair), not HBM bound"},
            "e2e": {"value": LOOP_NQ * K / (e2e_ms_max * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": LOOP_NQ * (80 + 4800), "d2h_bytes_per_step": LOOP_NQ * 16,
                    "api": "pinned host queries -> lmsf_scdb_search_shard_dev + all_gather + lmsf_scdb_pick_dev -> host results"},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_sc_scan (exact ring-key 10-NN scan of the shard)",
                         "achieved": (80.0 * (hi - lo) + 84.0 * LOOP_NQ) / (dev_ms_max / K * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": (80.0 * (hi - lo) + 84.0 * LOOP_NQ) / (dev_ms_max / K * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src, "traffic": None,
                         "note": "whole step time used as the kernel time (upper bound); algorithmic bytes = keys once + queries"},
        }
        if args.gpus == 1 and not args.no_cpu:
            line["cpu_baseline"] = loopdb_cpu_baseline(keys, descs, qk, qd, int(limit), ids)
        print(json.dumps(line), flush=True)
    # Everything torch allocated or copied on the context's (external) stream must be released while that stream
    # still exists: freeing a pinned block records an event on the streams it was used with.
    import gc
    torch.cuda.synchronize()
    del res, res_e, d_qk, d_qd, h_qk, h_qd, flush, ids
    gc.collect()
    torch.cuda.synchronize()
    ctx.close()
    if use_dist:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 3: multi-LiDAR rig, one VLP-16 stream per GPU (rig.DistributedRig).  A step = one synchronised set of
# sweeps in calibration status 1 (System/ML_System.hpp:296-310): the primary tracks, every auxiliary GPU extracts and
# ships its features over NCCL, the primary registers them against its local map and re-derives the extrinsics.
ML_METRIC = "multi-LiDAR sweep sets/sec (N x VLP-16, one stream per GPU, online extrinsic refinement)"


def ml_extrinsics(n):
    def rz(deg):
        a = np.radians(deg)
        return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    ring = [(np.eye(3), np.zeros(3))]
    for i in range(1, n):
        ang = (40.0 if i % 2 else -40.0) * (1 + (i - 1) // 2)
        ring.append((rz(ang), np.array([0.03, -0.54 if i % 2 else 0.54, -0.14])))
    return ring, rz


def run_multilidar(args):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    if world < 2:
        print(json.dumps({"metric": ML_METRIC, "unavailable": "the multi-LiDAR workload needs one rank per LiDAR (>= 2)"}))
        return
    W, K = args.warmup, args.steps
    S0 = 12                                     # status-0 sweeps (independent odometry, hand-eye fed)
    pkg = entry.load_package()
    synth = pkg.synth
    sensor = synth.vlp16()
    ext, rz = ml_extrinsics(world)
    n_sw = S0 + 2 * (W + K) + 2
    sweeps = [np.ascontiguousarray(synth.make_sweep(sensor, k, extrinsic=(None if rank == 0 else ext[rank])))
              for k in range(n_sw)]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = pkg.context(local, n_scans=16, max_points=1 << 16)
    rig = pkg.rig.DistributedRig(pkg.library(), ctx, rank, world, device=dev)
    stream = rig.stream
    sampler = ClockSampler(local)
    sampler.start()
    k = 0
    for _ in range(S0):
        rig.process(sweeps[k], 0.1 * k)
        k += 1
    gate_opened = rig.status == 1
    if not gate_opened:
        # the synthetic trajectory (yaw + 2 deg roll / pitch) does not open the hand-eye gate (rot_cov[2] > 0.25) within
        # S0 sweeps: the refinement starts from the simulated extrinsics perturbed by 1 deg / 5 cm
        rig.status = 1
        for i in range(1, world):
            Re, te = ext[i]
            rig.extrinsic[i] = synth.pose_to_qt(Re @ rz(1.0), te + np.array([0.04, -0.03, 0.02]))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        nonlocal k
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(n):
            rig.process(sweeps[k], 0.1 * k)
            k += 1
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3

    timed(W)
    l0 = ctx.launch_count()
    sampler.active.set()
    dev_ms, wall_ms = timed(K)
    sampler.active.clear()
    launches = ctx.launch_count() - l0
    sampler.stop_flag.set()
    sampler.join()
    dev_max, wall_max = pkg.shard.max_over_ranks([dev_ms, wall_ms], device=f"cuda:{local}")
    launch_sum = pkg.shard.sum_over_ranks([float(launches)], device=f"cuda:{local}")[0]
    # extrinsics the refinement ends at, against the simulated ones
    errs = []
    if rank == 0:
        for i in range(1, world):
            Re, te = ext[i]
            E = synth.qt_to_mat(rig.extrinsic[i])
            errs.append([float(np.linalg.norm(E[:3, 3] - te)),
                         float(np.degrees(np.arccos(np.clip((np.trace(E[:3, :3].T @ Re) - 1) / 2, -1, 1))))])
        n_pts = len(sweeps[0])
        line = {
            "metric": ML_METRIC, "value": K / (wall_max * 1e-3), "unit": "sweep sets/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": wall_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{world} x VLP-16 (28 800 pts/sweep) on one rig, one LiDAR per GPU, calibration status 1: "
                                   "primary tracks, auxiliary GPUs extract and ship features (NCCL send/recv), primary "
                                   "registers them against its local map (Huber-LM) and refines the extrinsics",
                       "lidars": world, "points_per_sweep": n_pts, "status0_sweeps": S0,
                       "handeye_gate_opened": bool(gate_opened),
                       "timing": "wall clock around K synchronised steps between barriers, max over ranks (a step mixes "
                                 "host control flow, NCCL and kernels on several streams); device time of rank 0's "
                                 "context stream beside it",
                       "device_ms_per_step_rank0_stream": dev_max / K,
                       "l2": "inputs change every step (new sweeps); no flush: the local maps (<= 10 x 28 800 points) fit L2 "
                             "by construction of this configuration",
                       "extrinsic_error_m_deg": errs},
            "e2e": {"value": K / (wall_max * 1e-3), "unit": "sweep sets/s",
                    "h2d_bytes_per_step": 16 * n_pts * world, "d2h_bytes_per_step": 888,
                    "api": "rig.DistributedRig.process: host sweeps in, poses and extrinsics out, every step"},
            "gpu_launches": int(launch_sum), "clocks": sampler.summary(),
        }
        line["config"]["feature_bytes_received_per_step"] = rig.shipped_bytes / max(1, W + K)
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    rig.close()
    del rig
    import gc
    gc.collect()
    ctx.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sensor", default="hdl64", choices=sorted(SENSORS), help="hdl64 = the headline configuration")
    ap.add_argument("--map-leaf", default="", help="E,S: voxel-filter the local map at these leaf sizes (LIO-SAM-style "
                    "0.2,0.4); default: the shipped tracker's raw window")
    ap.add_argument("--loop-exchange", type=int, default=2, choices=[1, 2],
                    help="loopdb: 1 = every rank scores its 10 local candidates, one all_gather; 2 = two rounds, only the "
                         "global top-10 are scored, by their owners")
    ap.add_argument("--workload", default="registration", choices=["registration", "loopdb", "multilidar"],
                    help="registration = the headline metric (default); loopdb = sharded loop-closure descriptor search")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global SENSOR, METRIC, MAP_LEAF, MAP_DESC
    SENSOR = args.sensor
    if args.map_leaf:
        le, ls = (float(x) for x in args.map_leaf.split(","))
        MAP_LEAF = {"map_leaf_edge": le, "map_leaf_surf": ls}
        MAP_DESC = f"10-keyframe sliding-window map voxel-filtered at {le} m (edge) / {ls} m (surf)"
    if SENSOR != "hdl64":
        METRIC = METRIC.replace("HDL-64", "VLP-16")
    if args.workload == "loopdb":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "loopdb workload: the CPU port is timed inside the "
                                                               "line's cpu_baseline"}))
            return
        run_loopdb(args)
    elif args.workload == "multilidar":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "multilidar workload: parity against three oracle "
                                                               "contexts is a test (tests/test_gpu_parity.py), not a bench arm"}))
            return
        run_multilidar(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
