"""Tuning aid: in-kernel phase times of k_fit / k_lm_eval (last block: own loop, reduction + wait for the other
blocks, 6x6 tail), from a library built with -DLMSF_TIMING (LMSF_B200_LIB=build/timing/liblmsf_b200.so)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
synth = pkg.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sensor = synth.hdl64()
sweeps = [synth.make_sweep(sensor, k) for k in range(n + 12)]
ctx = pkg.context(0, n_scans=64, max_points=1 << 18)
d = [ctx.dev_upload_new(s) for s in sweeps]
for k in range(10):
    ctx.tracker_step_dev(d[k], len(sweeps[k]), 0.1 * k)
out = (C.c_ulonglong * 32)()
ctx.lib.dll.lmsf_debug_kernel_times(out, 1)
for k in range(10, 10 + n):
    ctx.tracker_step_dev(d[k], len(sweeps[k]), 0.1 * k)
ctx.lib.dll.lmsf_debug_kernel_times(out, 1)
o = list(out)
for name, b in (("k_fit", 0), ("k_lm_eval (active)", 4)):
    c = max(1, o[b + 3])
    print(f"{name}: {o[b + 3]} launches; last block: loop {o[b] / c / 1e3:.1f} us, reduce+wait {o[b + 1] / c / 1e3:.1f} us, "
          f"tail {o[b + 2] / c / 1e3:.1f} us")
c = max(1, o[10])
print(f"k_knn lanes: {o[10]} queries; time from chunk start to the lane's finish: mean {o[9] / c / 1e3:.1f} us, max {o[8] / 1e3:.1f} us; "
      f"histogram <20/<40/<80/<160/>=160 us: {[o[11 + i] for i in range(5)]}")
