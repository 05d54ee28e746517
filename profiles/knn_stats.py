"""Tuning aid: distribution of kNN work per exit pass on a realistic tracker state (HDL-64, raw map)."""
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
for kind, f in ((0, e), (1, s)):
    q = np.ascontiguousarray((f[:, :3] @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
    out = (C.c_ulonglong * 28)()
    rc = ctx.lib.dll.lmsf_debug_knn_stats(ctx._h, kind, q.ctypes.data_as(C.c_void_p), len(q), out)
    o = np.array(out[:], dtype=np.int64)
    print(f"kind {kind}: map {st['map_edge'] if kind == 0 else st['map_surf']} pts, {len(q)} queries, rc={rc}")
    for l, name in ((0, "no map / out of range"), (1, "A (27 L2 cells)"), (2, "ball sweep"), (3, "B (27 L1 cells)"), (4, "C (27 L0 cells)")):
        n = o[l * 4]
        if n:
            print(f"  exit {name:24s}: {n:7d} queries ({100.0 * n / len(q):5.1f}%), full {o[20 + l]:7d}, per query: "
                  f"cand {o[l * 4 + 1] / n:8.1f}  box tests {o[l * 4 + 2] / n:7.1f}  lookups {o[l * 4 + 3] / n:6.1f}")
