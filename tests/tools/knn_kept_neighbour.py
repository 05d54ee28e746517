"""CPU experiment behind DESIGN.md section 8 (kNN): would a kept 6th (7th, 8th) neighbour from the first outer
iteration prove the 5-NN set of the second one without a search?

A query moves by delta between the two outer iterations (prediction pose -> solved pose).  By the triangle
inequality every map point's distance changes by at most delta, so the five nearest of the moved query are among
the first m of the old one whenever d5 + delta < d(m+1) - delta (m = 5: the set is unchanged).  This script runs
the oracle tracker (test infrastructure: that is why it lives under tests/) to the steady state of bench.py's
HDL-64 sequence and counts, for every surf feature of the next sweeps, how often that holds.

Result (HDL-64, sweeps 34-39, 1.27 M-point surf map, 126.8 k surf queries per sweep):
  delta p50 4.6-6.0 mm, p90 13-15 mm; d5 p50 38-39 mm; d6-d5 p50 2.5 mm, d7-d5 6 mm, d8-d5 9 mm
  unchanged 5-set provable from d6:   6-15 % of the queries
  5-set inside the first 6 (from d7): 17-34 %
  5-set inside the first 7 (from d8): 30-50 %
VLP-16 (281 k-point map, 28.1 k queries): delta p50 9-15 mm, d5 p50 89-118 mm, d6-d5 p50 6-8 mm: 6-21 % / 17-46 % / 31-64 %.
The raw ten-keyframe window is ten overlapping samplings of the same surfaces: neighbour gaps are millimetres, the
first LM step moves the queries by as much or more, and the shortcut would serve a small minority of the queries
(and leave the launch time to the rest, which is what bounds it).  Not built.

usage: python tests/tools/knn_kept_neighbour.py [hdl64|vlp16]
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

SETTLE, SWEEPS = 34, 6


def rigid(p):
    return Rotation.from_quat(p[:4]).as_matrix(), np.asarray(p[4:])


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hdl64"
    synth = entry.load_package().synth
    sensor, n_scans = (synth.hdl64(), 64) if name == "hdl64" else (synth.vlp16(), 16)
    o = entry.load_oracle().context(0, n_scans=n_scans, oracle_knn_mode=0, oracle_threads=os.cpu_count() or 1)
    poses = []
    for k in range(SETTLE + SWEEPS):
        sweep = synth.make_sweep(sensor, k)
        if k >= SETTLE:
            surf_map = o.get_map(1)
            # the tracker's constant-velocity prediction: curr = prev * (prev2^-1 * prev)  (LidarTrackerLocalMap.hpp:107-125)
            (R1, t1), (R0, t0) = rigid(poses[-1]), rigid(poses[-2])
            Rp, tp = R1 @ (R0.T @ R1), R1 @ (R0.T @ (t1 - t0)) + t1
        pose, _, st = o.tracker_step(sweep, 0.1 * k)
        poses.append(pose.copy())
        if k < SETTLE:
            continue
        _, _, surf = o.extract_features(sweep)
        Rf, tf = rigid(pose)
        xyz = surf[:, :3].astype(np.float64)
        q_pred = (xyz @ Rp.T + tp).astype(np.float32).astype(np.float64)   # pointAssociateToMap stores fp32
        q_solved = (xyz @ Rf.T + tf).astype(np.float32).astype(np.float64)
        delta = np.linalg.norm(q_pred - q_solved, axis=1)
        d, _ = cKDTree(surf_map[:, :3].astype(np.float64)).query(q_pred, k=8, workers=-1)
        ok = d[:, 4] ** 2 < 1.0                                           # the matchers' accept test
        line = (f"sweep {k}: {len(surf)} surf queries, map {len(surf_map)}, keyframe {st.get('keyframe')}, "
                f"delta p50 {np.median(delta) * 1e3:.2f} mm p90 {np.percentile(delta, 90) * 1e3:.2f} mm, "
                f"d5 p50 {np.median(d[ok, 4]) * 1e3:.1f} mm;")
        for m in (5, 6, 7):
            hit = ok & (d[:, 4] + delta < d[:, m] - delta)
            line += (f" d{m + 1}-d5 p50 {np.median(d[ok, m] - d[ok, 4]) * 1e3:.2f} mm -> 5-set inside first {m}: "
                     f"{100.0 * hit.sum() / ok.sum():.1f} %;")
        print(line)
    o.close()


if __name__ == "__main__":
    main()
