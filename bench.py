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
    ctx = pkg.context(local, n_scans=SENSORS[SENSOR][0], max_points=1 << 18)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2
    n_pts = [int(s.shape[0]) for s in sweeps]

    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def run_sequence(step_fn, profile, K=K):
        """W untimed + K timed steps; returns (device ms for K steps, per-step wall ms, launches, prof)."""
        ctx.tracker_reset()
        for k in range(0, SETTLE + 1):          # initialisation (see SETTLE), then the W warm-up steps
            step_fn(k, 0.1 * k)
        for k in range(SETTLE + 1, SETTLE + W + 1):
            step_fn(k, 0.1 * k)
        barrier()
        ctx.profile_enable(profile)
        ctx.profile_read(reset=True)
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = []
        stats = []
        flush_host = 0.0
        sampler.active.set()
        t_wall0 = time.perf_counter()
        e0.record(stream)
        for k in range(SETTLE + W + 1, SETTLE + W + K + 1):
            tf = time.perf_counter()
            with torch.cuda.stream(stream):
                flush.zero_()                       # L2 flush between timed steps, inside the timed region
            t0 = time.perf_counter()
            flush_host += t0 - tf
            if os.environ.get('LMSF_BENCH_DEBUG'):
                sys.stderr.write(f'flush {k} {(t0 - tf) * 1e3:.3f} ms\n')
            stats.append(step_fn(k, 0.1 * k))
            per.append((time.perf_counter() - t0) * 1e3)
        e1.record(stream)
        barrier()
        sampler.active.clear()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        launches = ctx.launch_count() - l0
        prof = ctx.profile_read(reset=True) if profile else None
        ctx.profile_enable(False)
        return dev_ms, wall_ms, per, launches, prof, stats, flush_host * 1e3 / K

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
    def step_dev(k, t):
        ctx.tracker_submit_dev(d_ptrs[k], n_pts[k], t)
        ctx.tracker_prefetch_dev(d_ptrs[k + 1], n_pts[k + 1])
        return ctx.tracker_wait()[2]

    def step_host(k, t):
        ctx.tracker_submit(sweeps[k], t)
        ctx.tracker_prefetch(sweeps[k + 1])
        return ctx.tracker_wait()[2]

    run_sequence(step_dev, False, K=2)
    dev_ms, wall_ms, per, launches, _, stats, fl1 = run_sequence(step_dev, False)
    # ---- end-to-end pass: host buffers through lmsf_tracker_prefetch + lmsf_tracker_step (H2D of a sweep + D2H of
    # the pose inside every step)
    e_dev_ms, e_wall_ms, e_per, _, _, _, fl2 = run_sequence(step_host, False)
    # ---- instrumented pass (not used for value): per-stage CUDA-event times on the context's streams
    p_dev_ms, _, _, _, prof, _, fl3 = run_sequence(step_dev, True)
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
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "match_traffic_bytes.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        kf = sum(1 for s in stats if s["keyframe"])
        line = {
            "metric": METRIC, "value": n_gpus * K / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": n_gpus,
            "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": SENSORS[SENSOR][1] + " scan-to-map edge/surf registration, "
                                   "Huber-LM solver, raw 10-keyframe sliding-window map, one sequence per GPU",
                       "points_per_sweep": int(np.mean(n_pts)), "sequences": n_gpus,
                       "l2": f"{FLUSH_BYTES >> 20} MiB memset between timed steps (L2 flush: 1.5x the 126 MB L2), inside the timed region",
                       "timing": "CUDA events on the context stream around the K steps, max over ranks",
                       "initialisation": f"{SETTLE} untimed sweeps before the warm-up: full 10-keyframe window, LM budget at its floor",
                       "pipeline": "front end (extraction of sweep k+1) overlaps the registration of sweep k",
                       "p50_ms_per_scan": float(np.median(per)), "p90_ms_per_scan": float(np.percentile(per, 90)),
                       "max_ms_per_scan": float(np.max(per)), "wall_ms_per_step": wall_ms / K,
                       "flush_host_ms_per_step": [fl1, fl2, fl3], "instrumented_pass_ms_per_step": p_dev_ms / K,
                       "keyframes_in_timed_region": kf, "per_rank_ms_per_step_and_keyframes": per_rank,
                       "features_per_sweep": int(np.mean([s["n_edge"] + s["n_surf"] for s in stats])),
                       "map_points_end": int(stats[-1]["map_edge"] + stats[-1]["map_surf"]),
                       "stage_ms_per_step": {k: v / K for k, v in ms.items()},
                       "stage_launches_per_step": {k: v / K for k, v in ln.items()}},
            "e2e": {"value": n_gpus * K / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(np.mean(n_pts)) * 16, "d2h_bytes_per_step": STATE_D2H_BYTES,
                    "p50_ms_per_scan": float(np.median(e_per)),
                    "api": "lmsf_tracker_submit(host sweep) + lmsf_tracker_prefetch(next host sweep) + lmsf_tracker_wait(pose out), "
                           "wall clock around K steps, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_knn (exact 5-NN of every scan feature over the local-map grid)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": match_ms,
                         "launches_per_step": ln["match"] / K,
                         "share_of_step": ms["match"] / max(1e-9, p_dev_ms)},
        }
        if n_gpus == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(sweeps, W)
        print(json.dumps(line), flush=True)
    for p in d_ptrs:
        ctx.dev_free(p)
    ctx.close()
    if use_dist:
        dist.destroy_process_group()


def cpu_baseline(sweeps, W, sample=4):
    """The oracle port of the reference's CPU path on a bounded sample of the same workload: the
    tracker is brought to the steady state with all host threads (untimed), then `sample` sweeps are
    timed with ONE thread — the reference runs one thread per LiDAR (System/ML_System.hpp:137,248)."""
    lib = entry.load_oracle()
    threads = os.cpu_count() or 1
    o = lib.context(0, n_scans=SENSORS[SENSOR][0], oracle_knn_mode=0, oracle_threads=threads)
    warm = min(SETTLE + W, len(sweeps) - sample - 1)
    for k in range(0, warm + 1):
        o.tracker_step(sweeps[k], 0.1 * k)
    lib.fn("set_threads")(o._h, 1)
    t0 = time.perf_counter()
    for k in range(warm + 1, warm + 1 + sample):
        o.tracker_step(sweeps[k], 0.1 * k)
    dt = time.perf_counter() - t0
    o.close()
    return {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample} consecutive {SENSOR} sweeps after a {warm}-sweep warm-up of the same sequence, "
                      f"oracle tracker (kd-tree kNN), 1 thread; host has {threads} cores",
            "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be
    built here (no PCL/Eigen/Ceres; its tree lacks Map/), so this times the oracle port with all the host
    threads it can use (OpenMP over features, as the reference's unbuilt LIO-SAM copy does)."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    W, K = args.warmup, args.steps
    sweeps = make_sequence(SETTLE + W + K + 1, seq=0)
    lib = entry.load_oracle()
    threads = os.cpu_count() or 1
    o = lib.context(0, n_scans=SENSORS[SENSOR][0], oracle_knn_mode=0, oracle_threads=threads)
    for k in range(0, SETTLE + W + 1):
        o.tracker_step(sweeps[k], 0.1 * k)
    per = []
    t0 = time.perf_counter()
    for k in range(SETTLE + W + 1, SETTLE + W + K + 1):
        t1 = time.perf_counter()
        o.tracker_step(sweeps[k], 0.1 * k)
        per.append((time.perf_counter() - t1) * 1e3)
    dt = time.perf_counter() - t0
    o.close()
    v = K / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": SENSORS[SENSOR][1] + " scan-to-map edge/surf registration, "
                               "Huber-LM solver, raw 10-keyframe sliding-window map, one sequence",
                   "p50_ms_per_scan": float(np.median(per))},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{K} consecutive sweeps after {SETTLE} initialisation + {W} warm-up sweeps, oracle tracker, "
                                   f"{threads} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# --workload loopdb: BASELINE.json configs[4] — loop-closure descriptor database search over 100 000 keyframes,
# database sharded over the ranks, per-rank candidates merged with one NCCL all_gather (shard.loop_search_sharded).
# Not the driver's default line; run explicitly:  python bench.py --workload loopdb [--gpus N]
LOOP_METRIC = "loop-closure descriptor searches/sec (ScanContext, 100k-keyframe database)"
LOOP_DB = 100_000
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

    def step_resident():
        return pkg.shard.loop_search_sharded(ctx, d_qk, d_qd, limit, lo, hi, 0.2)

    def step_e2e():
        with torch.cuda.stream(stream):
            a = h_qk.to(dev, non_blocking=True)
            b = h_qd.to(dev, non_blocking=True)
        ids, dd, sh = pkg.shard.loop_search_sharded(ctx, a, b, limit, lo, hi, 0.2)
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
                                   f"{LOOP_NQ} replicated queries per step, exact ring-key top-10 + SC distance per shard, "
                                   "one all_gather of 24-byte candidates, identical selection on every rank",
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
            import ctypes as C
            lib = C.CDLL(entry.ORACLE_LIB)
            nqs = 16
            f32p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
            o_id = np.empty(nqs, np.int32)
            o_d = np.empty(nqs, np.float64)
            o_s = np.empty(nqs, np.int32)
            kk = np.ascontiguousarray(keys)
            dd = np.ascontiguousarray(descs.reshape(-1, 1200))
            t0 = time.perf_counter()
            lib.lmsf_oracle_sc_search(kk.ctypes.data_as(f32p), dd.ctypes.data_as(f32p), C.c_int(int(limit)),
                                      qk[:nqs].ctypes.data_as(f32p), qd[:nqs].reshape(nqs, 1200).ctypes.data_as(f32p),
                                      C.c_int(nqs), C.c_double(0.2), o_id.ctypes.data_as(i32p), o_d.ctypes.data_as(f64p),
                                      o_s.ctypes.data_as(i32p))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nqs / dt, "unit": "queries/s", "cores": 1, "kind": "port",
                                    "sample": f"{nqs} of the {LOOP_NQ} queries against the same database, oracle "
                                              "brute-force ring-key scan + SC distance, 1 thread",
                                    "agrees_with_gpu": bool(np.array_equal(o_id, ids[:nqs]))}
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sensor", default="hdl64", choices=sorted(SENSORS), help="hdl64 = the headline configuration")
    ap.add_argument("--workload", default="registration", choices=["registration", "loopdb"],
                    help="registration = the headline metric (default); loopdb = sharded loop-closure descriptor search")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global SENSOR, METRIC
    SENSOR = args.sensor
    if SENSOR != "hdl64":
        METRIC = METRIC.replace("HDL-64", "VLP-16")
    if args.workload == "loopdb":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "loopdb workload: the CPU port is timed inside the "
                                                               "line's cpu_baseline"}))
            return
        run_loopdb(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
