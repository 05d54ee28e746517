"""Tuning aid: population of the 1 m cells of the steady-state surf map (HDL-64 bench sequence, sweep 30): 3 446 occupied
cells, 369 points on average, 53 % of the points in the 307 cells that hold more than 1 024 points."""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as entry
pkg = entry.load_package(); synth = pkg.synth
sensor = synth.hdl64()
ctx = pkg.context(0, n_scans=64, max_points=1 << 18)
for k in range(30):
    ctx.tracker_step(np.ascontiguousarray(synth.make_sweep(sensor, k)), 0.1 * k)
m = ctx.get_map(1)
c = np.floor(m[:, :3]).astype(np.int64)
key = (c[:, 0] + 2048) | ((c[:, 1] + 2048) << 16) | ((c[:, 2] + 2048) << 32)
_, cnt = np.unique(key, return_counts=True)
print("points", len(m), "cells", len(cnt), "mean", cnt.mean())
for lo, hi in ((1, 32), (33, 64), (65, 128), (129, 256), (257, 512), (513, 1024), (1025, 100000)):
    sel = (cnt >= lo) & (cnt <= hi)
    print(f"{lo:5d}-{hi:6d}: cells {sel.sum():6d} ({100*sel.mean():5.1f}%), points {cnt[sel].sum():8d} ({100*cnt[sel].sum()/len(m):5.1f}%)")
