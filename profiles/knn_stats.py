"""Tuning aid: work counters of the unseeded kNN search (csrc/debug_stats.cu, lmsf_debug_knn_stats) on a realistic
tracker state (HDL-64, raw 10-keyframe window): sweeps started per state, cells and candidates per query."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
synth = pkg.synth
nsw = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sensor = synth.hdl64()
ctx = pkg.context(0, n_scans=64)
for k in range(nsw):
    p, d, st = ctx.tracker_step(synth.make_sweep(sensor, k), 0.1 * k)
_, e, s = ctx.extract_features(synth.make_sweep(sensor, nsw))
T = synth.qt_to_mat(p)
names = ["START", "BALL", "GROW", "SHELL", "FIND", "CELL", "LAST"]
for kind, f in ((0, e), (1, s)):
    q = np.ascontiguousarray((f[:, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    out = (C.c_ulonglong * 28)()
    rc = ctx.lib.dll.lmsf_debug_knn_stats(ctx._h, kind, q.ctypes.data_as(C.c_void_p), len(q), out)
    o = np.array(out[:], dtype=np.int64)
    n = max(1, int(o[16]))
    print(f"kind {kind}: map {st['map_edge'] if kind == 0 else st['map_surf']} pts, {len(q)} queries, rc={rc}, "
          f"with 5 neighbours {o[17]}")
    print(f"  per query: L0 lookups {o[0] / n:.2f}, L1 cells {o[1] / n:.2f}, segments {o[3] / n:.2f}, candidates {o[2] / n:.1f}")
    print("  sweeps per query by state: " + ", ".join(f"{nm} {o[8 + i] / n:.3f}" for i, nm in enumerate(names)))
