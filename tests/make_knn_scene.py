"""Record a realistic kNN scene for csrc/test_knn_model (work profile of the search, CPU only):
the bench's steady state — map = ten keyframes (every other sweep) of the synthetic HDL-64 sequence in the world
frame, queries = the next sweep in ring order.   python tests/make_knn_scene.py /tmp/scene.bin [sensor]"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "lmsf-slam_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
sys.modules["synth"] = synth
spec.loader.exec_module(synth)

out = sys.argv[1]
S = synth.sensor_by_name(sys.argv[2] if len(sys.argv) > 2 else "hdl64")


def world(k):
    p = synth.make_sweep(S, k)[:, :3].astype(np.float64)
    R, t = synth.trajectory_pose(k)
    return (p @ R.T + t).astype(np.float32)


m = np.concatenate([world(k) for k in range(20, 40, 2)])
q = world(40)
q = q.reshape(S.n_az, S.n_scans, 3).transpose(1, 0, 2).reshape(-1, 3)  # ring order: consecutive queries are neighbours
m4 = np.concatenate([m, np.zeros((len(m), 1), np.float32)], 1)
with open(out, "wb") as f:
    np.array([len(m4), len(q)], np.int32).tofile(f)
    np.ascontiguousarray(m4, np.float32).tofile(f)
    np.ascontiguousarray(q, np.float32).tofile(f)
print(out, len(m4), len(q))
