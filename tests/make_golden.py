"""Regenerate tests/golden/*.npz — small fixtures that pin the ORACLE's current outputs.

The reference holds no golden vectors for this path (parity unpinned) and cannot be
built or imported here, so these fixtures are produced by the oracle itself from the seeded synthetic
generator; they guard the oracle (and through it the CUDA path) against silent drift between rounds.
Run:  python tests/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_sweep(synth, k):
    """VLP-16 sweep thinned to 16 x 450 rays (7 200 points) to keep the fixture small."""
    sw = synth.make_sweep(synth.vlp16(), k)
    az = (np.arange(len(sw)) // 16)
    return np.ascontiguousarray(sw[az % 4 == 0])


def sc_golden_inputs(synth, sco):
    """Seeded inputs of the loop-closure fixture: descriptors of six small sweeps, a 300-entry database derived from
    them (scaled, column-shifted), six revisit queries (column-shifted database entries)."""
    import sc_helpers as sch
    made = [sco.make(small_sweep(synth, 2 * k)) for k in range(6)]
    descs = np.stack([m[0] for m in made])
    keys0 = np.stack([m[1] for m in made])
    rng = np.random.default_rng(20260600)
    src = rng.integers(0, 6, size=300)
    scale = rng.uniform(0.7, 1.3, size=300)
    shift = rng.integers(0, 60, size=300)
    db = np.stack([np.roll(descs[i] * np.float32(s), int(sh), axis=1) for i, s, sh in zip(src, scale, shift)]).astype(np.float32)
    dbk = sch.keys_of_fast(db)
    qid = np.array([5, 40, 111, 170, 222, 249])
    qd = np.stack([np.roll(db[i], 7 * j, axis=1) for j, i in enumerate(qid)]).astype(np.float32)
    return descs, keys0, db, dbk, qid, qd, sch.keys_of_fast(qd)


def main():
    pkg = entry.load_package()
    synth = pkg.synth
    lib = entry.load_oracle()
    o = lib.context(0, n_scans=16, oracle_knn_mode=1)
    s0, s1 = small_sweep(synth, 0), small_sweep(synth, 3)
    lab, e0, f0 = o.extract_features(s0)
    vox, mem = o.voxel_downsample(f0, 0.4)
    o.map_set(0, e0)
    o.map_set(1, vox)
    _, e1, f1 = o.extract_features(s1)
    q = np.ascontiguousarray(f1[::5, :3])
    idx, d2 = o.knn5(1, q)
    ok, out = o.match(1, q)
    oke, oute = o.match(0, np.ascontiguousarray(e1[:, :3]))
    o.set_lm_outer(10)
    p_lm, st_lm = o.register(e1, f1, solver=1)
    p_gn, st_gn = o.register(e1, f1, solver=0)
    np.savez_compressed(
        os.path.join(OUT, "vlp16_small.npz"), sweep0=s0, sweep1=s1, label0=lab, n_edge0=len(e0), n_surf0=len(f0),
        edge0=e0, surf0_head=f0[:64], vox=vox, mem=mem, q=q, knn_idx=idx, knn_d2=d2, match_ok=ok, match_out=out,
        ematch_ok=oke, ematch_out=oute, pose_lm=p_lm, pose_gn=p_gn,
        lm_stats=np.array([st_lm["outer_iters"], st_lm["n_edge_matched"], st_lm["n_surf_matched"],
                           st_lm["lm_steps_total"], st_lm["lm_steps_accepted"]]),
        gn_stats=np.array([st_gn["outer_iters"], st_gn["n_edge_matched"], st_gn["n_surf_matched"],
                           st_gn["converged"], st_gn["degenerate"]]))
    # tracker trajectory over 8 small sweeps
    t = lib.context(0, n_scans=16, map_leaf_edge=0.2, map_leaf_surf=0.4, oracle_knn_mode=0)
    poses, kfs = [], []
    for k in range(8):
        p, d, st = t.tracker_step(small_sweep(synth, k), 0.1 * k)
        poses.append(p)
        kfs.append(st["keyframe"])
    np.savez_compressed(os.path.join(OUT, "vlp16_small_track.npz"), poses=np.array(poses), keyframes=np.array(kfs))
    # loop-closure descriptors + alignment score (rows f1 / f2)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sc_helpers as sch
    sco = sch.ScOracle()
    descs, keys0, db, dbk, qid, qd, qk = sc_golden_inputs(synth, sco)
    kidx, kd = sco.knn(dbk, 250, qk)
    lid, ldist, lsh = sco.search(dbk, db, 250, qk, qd)
    pd, ps = sco.distance(qd, db[qid])
    a_sc, a_ov, a_n = o.align_score(1, f1[::4], np.eye(4), 1.0, 0.3)
    np.savez_compressed(os.path.join(OUT, "sc_small.npz"), desc=descs, key=keys0, knn_idx=kidx, knn_d=kd, loop_id=lid,
                        loop_dist=ldist, loop_shift=lsh, pair_dist=pd, pair_shift=ps, align=np.array([a_sc, a_ov, a_n]))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
