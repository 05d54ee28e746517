"""Tuning aid: host-side time per phase of lmsf_tracker_submit_ticket / _prefetch_dev / _wait in steady state (HDL-64), next to the device step time."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
synth = pkg.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
sensor = synth.hdl64()
WARM = 34  # bench.py's SETTLE + default warm-up: full keyframe window, LM budget at its floor
sweeps = [synth.make_sweep(sensor, k) for k in range(n + WARM + 2)]
ctx = pkg.context(0, n_scans=64, max_points=1 << 18)
d = [ctx.dev_upload_new(s) for s in sweeps]
ticket = ctx.tracker_prefetch_dev(d[0], len(sweeps[0]))  # the calls bench.py times: every step consumes the ticket of its
for k in range(WARM):                                      # sweep and prefetches the next one
    ctx.tracker_submit_ticket(ticket, 0.1 * k)
    ticket = ctx.tracker_prefetch_dev(d[k + 1], len(sweeps[k + 1]))
    ctx.tracker_wait()
us = (C.c_double * 4)()
cnt = (C.c_int64 * 4)()
ctx.lib.dll.lmsf_debug_host_times(ctx._h, us, cnt, 1)
per = []
kf = []
t00 = time.perf_counter()
for k in range(WARM, WARM + n):
    t0 = time.perf_counter()
    ctx.tracker_submit_ticket(ticket, 0.1 * k)
    ticket = ctx.tracker_prefetch_dev(d[k + 1], len(sweeps[k + 1]))
    t1 = time.perf_counter()
    _, _, st = ctx.tracker_wait()
    t2 = time.perf_counter()
    per.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6))
    kf.append(st["keyframe"])
tot = (time.perf_counter() - t00) * 1e6 / n
ctx.lib.dll.lmsf_debug_host_times(ctx._h, us, cnt, 1)
per = np.array(per)
kf = np.array(kf) > 0
print(f"steps {n}, keyframes {kf.sum()}, wall per step {tot:.0f} us")
print(f"submit + prefetch calls: {per[:, 0].mean():.0f} us; wait call: {per[:, 1].mean():.0f} us (KF steps {per[kf, 1].mean():.0f}, others {per[~kf, 1].mean():.0f})")
print(f"step call following a KF step: {per[1:][kf[:-1], 1].mean():.0f} us, following a non-KF step: {per[1:][~kf[:-1], 1].mean():.0f} us")
for i, name in enumerate(["solve (enqueue + wait)", "map update enqueue", "solve enqueue only"]):
    if cnt[i]:
        print(f"  {name}: {us[i] / cnt[i]:.0f} us x {cnt[i]}")
