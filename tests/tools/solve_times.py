"""Tuning aid: in-kernel phase times of k_solve as block 0 sees them (library built with -DLMSF_TIMING,
LMSF_B200_LIB=build_variants/timing/liblmsf_b200.so): fit loop, per barrier the time from arrival to the observed release
(= the slowest block + reduction + 6x6 step of the last block + release), the evaluation loops, whole kernel."""
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
c = max(1, o[17])
print(f"k_solve: {o[17]} launches, {o[28]} evaluations; block 0: fit loop {o[16] / c / 1e3:.1f} us, whole kernel {o[29] / c / 1e3:.1f} us")
print("  barrier k (arrival -> release seen), summed over launches / launches:", [round(o[18 + k] / c / 1e3, 1) for k in range(5)])
print("  evaluation loop k, summed over launches / launches:", [round(o[23 + k] / c / 1e3, 1) for k in range(5)])
c = max(1, o[10])
print(f"  last block, barriers k >= 1 ({o[10]}): arrival -> totals ready {o[8] / c / 1e3:.1f} us, totals -> release {o[9] / c / 1e3:.1f} us")
print(f"    of which state load {o[4] / c / 1e3:.1f}, 6x6 update {o[5] / c / 1e3:.1f}, state store {o[6] / c / 1e3:.1f}, fence + release {o[7] / c / 1e3:.1f} us")
