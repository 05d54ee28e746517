"""Helper of tests/test_gpu_parity.py::test_fused_solve_kernel_matches_split_launches: tracks the first N sweeps of a
synthetic sequence with the CUDA library under the caller's environment and prints the poses as JSON."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry

pkg = entry.load_package()
name, n = sys.argv[1], int(sys.argv[2])
sensor = getattr(pkg.synth, name)()
ctx = pkg.library().context(0, n_scans=64 if name == "hdl64" else 16)
out = []
for k in range(n):
    pose, delta, st = ctx.tracker_step(np.ascontiguousarray(pkg.synth.make_sweep(sensor, k)), 0.1 * k)
    reg = st["reg"]
    out.append({"pose": [float(v) for v in pose],
                "stats": [reg["n_edge_matched"], reg["n_surf_matched"], reg["lm_steps_total"], reg["lm_steps_accepted"],
                          st["keyframe"]]})
ctx.close()
print(json.dumps(out))
