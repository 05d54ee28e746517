"""Tuning aid: per-query time of the warp-per-query kNN search (k_knn_sparse) on the bench state, from a library built
with -DLMSF_TIMING (LMSF_B200_LIB=build_variants/timing/liblmsf_b200.so)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry

pkg = entry.load_package()
synth = pkg.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sensor = synth.hdl64()
sweeps = [synth.make_sweep(sensor, k) for k in range(n + 30)]
ctx = pkg.context(0, n_scans=64, max_points=1 << 18)
d = [ctx.dev_upload_new(s) for s in sweeps]
for k in range(28):
    ctx.tracker_step_dev(d[k], len(sweeps[k]), 0.1 * k)
out = (C.c_ulonglong * 32)()
ctx.lib.dll.lmsf_debug_kernel_times(out, 1)
for k in range(28, 28 + n):
    ctx.tracker_step_dev(d[k], len(sweeps[k]), 0.1 * k)
ctx.lib.dll.lmsf_debug_kernel_times(out, 1)
o = list(out)
c = max(1, o[10])
print(f"k_knn_sparse: {o[10]} queries over {n} sweeps ({o[10] / n / 2:.0f} per launch); per query mean {o[9] / c / 1e3:.1f} us, "
      f"max {o[8] / 1e3:.1f} us; histogram < 5 / 10 / 20 / 40 / >= 40 us: {[o[11 + i] for i in range(5)]}")
