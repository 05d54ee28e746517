"""Bounds check of the map index as the registration kernels walk it (library built with `make EXTRA=-DLMSF_KNN_CHECK`,
LMSF_B200_LIB=build_variants/check/liblmsf_b200.so): tracks N sweeps — keyframes, a window that fills and evicts — and
prints the first violation the searches recorded (segment starts / ends, L1 record and L2 start offsets) or `ok`."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry

pkg = entry.load_package()
name, n = sys.argv[1], int(sys.argv[2])
sensor = getattr(pkg.synth, name)()
ctx = pkg.library().context(0, n_scans=64 if name == "hdl64" else 16, window=3)
out = (C.c_int * 8)()
kf = 0
for k in range(n):
    pose, delta, st = ctx.tracker_step(np.ascontiguousarray(pkg.synth.make_sweep(sensor, k)), 0.1 * k)
    kf += 1 if st["keyframe"] else 0
    ctx.lib.dll.lmsf_debug_knn_check(ctx._h, out)
    if out[0]:
        print("violation at sweep", k, list(out))
        sys.exit(1)
print("ok:", n, "sweeps,", kf, "keyframes, matched", st["reg"]["n_edge_matched"], st["reg"]["n_surf_matched"])
