"""GPU parity tests of the loop-closure descriptor path (scancontext.cu) against the CPU oracle:
descriptors and ring keys bit-exact, ring-key 10-NN identical (distances bit for bit), ScanContext
distance bit-exact in fp64 (sqrt and division only), loop ids / shifts identical; plus size-independent
round trips at BASELINE config 5's database size and the sharded selection kernel."""
import numpy as np
import pytest

import sc_helpers as sch

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def sco(oracle_lib):
    return sch.ScOracle()


@pytest.fixture()
def gctx(gpu_lib):
    c = gpu_lib.context(0, n_scans=16, max_points=1 << 15, max_map_points=1 << 16)
    yield c
    c.close()


@pytest.mark.parametrize("name,k", [("vlp16", 0), ("vlp16", 9), ("hdl64", 2), ("hdl64", 6)])
def test_sc_make_bit_exact(gctx, sco, sweeps, name, k):
    sw = sweeps(name, k)
    d_g, k_g = gctx.sc_make(sw)
    d_o, k_o = sco.make(sw)
    assert np.array_equal(u32(d_g), u32(d_o))
    assert np.array_equal(u32(k_g), u32(k_o))


def test_sc_make_edge_cases(gctx, sco, sweeps):
    cases = [
        np.zeros((0, 4), np.float32),
        np.array([[1.0, 0.0, 0.5, 0], [0.0, 0.0, 9.0, 0], [79.9, 0.0, 1.0, 0], [80.5, 0.0, 5.0, 0],
                  [-3.0, -3.0, -1.5, 0], [np.nan, 1.0, 3.0, 0], [2.0, 2.0, np.nan, 0], [5.0, -5.0, -1003.0, 0],
                  [np.inf, 1.0, 2.0, 0], [1.0, -np.inf, 2.0, 0]], np.float32),
        sweeps("vlp16", 4)[:37],
        sweeps("vlp16", 4, dropout=0.5),
    ]
    for pts in cases:
        d_g, k_g = gctx.sc_make(pts)
        d_o, k_o = sco.make(pts)
        assert np.array_equal(u32(d_g), u32(d_o)) and np.array_equal(u32(k_g), u32(k_o))


def test_scdb_add_cloud_and_get(gctx, sco, sweeps):
    ids = [gctx.scdb_add_cloud(sweeps("vlp16", k)) for k in range(3)]
    assert ids == [0, 1, 2] and gctx.scdb_size() == 3
    for k in range(3):
        d_g, k_g = gctx.scdb_get(k)
        d_o, k_o = sco.make(sweeps("vlp16", k))
        assert np.array_equal(u32(d_g), u32(d_o)) and np.array_equal(u32(k_g), u32(k_o))
    gctx.scdb_clear()
    assert gctx.scdb_size() == 0


@pytest.mark.parametrize("n,limit", [(5000, 5000), (5000, 4999), (777, 300), (64, 4), (300, 256), (300, 257),
                                     (20000, 20000), (20000, 16384), (20000, 16383)])   # >= 16384: thresholded two-pass scan
def test_ringkey_knn_identical(gctx, sco, n, limit):
    rng = np.random.default_rng(n + limit)
    keys = rng.uniform(0, 5, size=(n, 20)).astype(np.float32)
    keys[n // 2] = keys[n // 3]                                  # an exact duplicate: tie broken by id
    q = np.concatenate([keys[:: max(1, n // 60)] + rng.normal(0, 0.03, size=(len(keys[:: max(1, n // 60)]), 20)).astype(np.float32),
                        keys[n // 3][None], rng.uniform(0, 5, size=(33, 20)).astype(np.float32)])
    gctx.scdb_add(np.zeros((n, 1200), np.float32), keys)
    i_g, d_g = gctx.scdb_knn(q, limit)
    i_o, d_o = sco.knn(keys, limit, q)
    assert np.array_equal(d_g.view(np.uint32), d_o.view(np.uint32))
    assert np.array_equal(i_g, i_o)


def test_sc_distance_bit_exact(gctx, sco):
    a = sch.random_descs(40, seed=21)
    b = sch.random_descs(40, seed=22)
    b[:10] = np.stack([np.roll(a[i], 5 * i, axis=1) for i in range(10)])   # true revisits
    a[10] = 0.0                                                            # no effective column
    d_g, s_g = gctx.sc_distance(a, b)
    d_o, s_o = sco.distance(a, b)
    assert np.array_equal(d_g.view(np.uint64), d_o.view(np.uint64))
    assert np.array_equal(s_g, s_o)
    assert np.all(np.abs(d_g[:10]) < 1e-12)


def _revisit_queries(descs, ids, shifts, noise, seed):
    rng = np.random.default_rng(seed)
    qd = np.stack([np.roll(descs[i], s, axis=1) for i, s in zip(ids, shifts)]).astype(np.float32)
    if noise > 0:
        qd = np.where(qd > 0, qd + rng.normal(0, noise, size=qd.shape).astype(np.float32), qd).astype(np.float32)
    return qd, sch.keys_of_fast(qd)


def test_search_identical_to_oracle(gctx, sco):
    n = 3000
    descs = sch.random_descs(n, seed=31)
    keys = sch.keys_of_fast(descs)
    rng = np.random.default_rng(32)
    ids = rng.integers(0, n, size=120)
    shifts = rng.integers(0, 60, size=120)
    qd, qk = _revisit_queries(descs, ids, shifts, 0.05, 33)
    far = sch.random_descs(20, seed=34)
    qd = np.concatenate([qd, far])
    qk = np.concatenate([qk, sch.keys_of_fast(far)])
    gctx.scdb_add(descs, keys)
    for limit in (n, n - 50, 1234, 7):
        l_g, d_g, s_g = gctx.scdb_search(qk, qd, limit)
        l_o, d_o, s_o = sco.search(keys, descs, limit, qk, qd)
        assert np.array_equal(l_g, l_o), limit
        assert np.array_equal(d_g.view(np.uint64), d_o.view(np.uint64)), limit
        assert np.array_equal(s_g, s_o), limit
    l_g, _, _ = gctx.scdb_search(qk, qd, n)
    assert (l_g[:120] == ids).mean() > 0.95 and np.all(l_g[120:] == -1)


def test_sharded_selection_equals_unsharded(gpu_lib, sco):
    """Three database shards (three contexts on one GPU), candidate blocks concatenated as an all_gather
    would, then lmsf_scdb_pick_dev: identical to the unsharded oracle search."""
    import torch

    n, nq = 2000, 64
    descs = sch.random_descs(n, seed=41)
    keys = sch.keys_of_fast(descs)
    rng = np.random.default_rng(42)
    ids = rng.integers(0, n, size=nq)
    qd, qk = _revisit_queries(descs, ids, rng.integers(0, 60, size=nq), 0.05, 43)
    limit = 1900
    bounds = [(0, 700), (700, 1300), (1300, 2000)]
    ctxs = [gpu_lib.context(0, n_scans=16, max_points=1 << 15, max_map_points=1 << 16) for _ in bounds]
    try:
        tq_k = torch.from_numpy(qk).cuda()
        tq_d = torch.from_numpy(qd.reshape(nq, 1200)).cuda()
        torch.cuda.synchronize()
        blocks = []
        for c, (lo, hi) in zip(ctxs, bounds):
            c.scdb_add(descs[lo:hi], keys[lo:hi])
            cand = torch.empty((nq, 10, 24), dtype=torch.uint8, device="cuda")
            c.scdb_search_shard_dev(tq_k.data_ptr(), tq_d.data_ptr(), nq, max(0, min(limit, hi) - lo), lo, cand.data_ptr())
            torch.cuda.synchronize()
            blocks.append(cand)
        allc = torch.stack(blocks).contiguous()
        out_id = torch.empty(nq, dtype=torch.int32, device="cuda")
        out_d = torch.empty(nq, dtype=torch.float64, device="cuda")
        out_s = torch.empty(nq, dtype=torch.int32, device="cuda")
        ctxs[0].scdb_pick_dev(allc.data_ptr(), 3, nq, 0.2, out_id.data_ptr(), out_d.data_ptr(), out_s.data_ptr())
        torch.cuda.synchronize()
        l_o, d_o, s_o = sco.search(keys, descs, limit, qk, qd)
        assert np.array_equal(out_id.cpu().numpy(), l_o)
        assert np.array_equal(out_d.cpu().numpy().view(np.uint64), d_o.view(np.uint64))
        assert np.array_equal(out_s.cpu().numpy(), s_o)
        # the two-round exchange on the same shards: unscored ring-key records first, then every shard scores what it
        # owns of the GLOBAL top-10, then the same pick — identical results, 10 scored candidates per query in total
        kblocks = []
        for c, (lo, hi) in zip(ctxs, bounds):
            cand = torch.empty((nq, 10, 24), dtype=torch.uint8, device="cuda")
            c.scdb_keys_shard_dev(tq_k.data_ptr(), nq, max(0, min(limit, hi) - lo), lo, cand.data_ptr())
            torch.cuda.synchronize()
            kblocks.append(cand)
        allk = torch.stack(kblocks).contiguous()
        sblocks = []
        for c, (lo, hi) in zip(ctxs, bounds):
            sc = torch.empty((nq, 10, 24), dtype=torch.uint8, device="cuda")
            c.scdb_score_owned_dev(allk.data_ptr(), 3, nq, tq_d.data_ptr(), lo, max(0, min(limit, hi) - lo), sc.data_ptr())
            torch.cuda.synchronize()
            sblocks.append(sc)
        alls = torch.stack(sblocks).contiguous()
        o2_id = torch.empty(nq, dtype=torch.int32, device="cuda")
        o2_d = torch.empty(nq, dtype=torch.float64, device="cuda")
        o2_s = torch.empty(nq, dtype=torch.int32, device="cuda")
        ctxs[1].scdb_pick_dev(alls.data_ptr(), 3, nq, 0.2, o2_id.data_ptr(), o2_d.data_ptr(), o2_s.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(o2_id, out_id) and torch.equal(o2_s, out_s)
        assert np.array_equal(o2_d.cpu().numpy().view(np.uint64), d_o.view(np.uint64))
        srec = alls.cpu().numpy().view(sch.CAND_DTYPE).reshape(3, nq, 10)
        assert int((srec["id"] >= 0).sum()) == nq * 10
        # the numpy selection used by the gloo test agrees with the kernel on the same records
        rec = allc.cpu().numpy().view(sch.CAND_DTYPE).reshape(3, nq, 10)
        l_p, d_p, s_p = sch.py_pick(rec)
        assert np.array_equal(l_p, l_o) and np.array_equal(s_p, s_o)
    finally:
        for c in ctxs:
            c.close()


def test_round_trip_at_full_database_size(gpu_lib):
    """BASELINE config 5 size (100 000 keyframes): every query is a column-shifted copy of a database entry,
    so it must come back with its own id, distance ~0 and exactly its shift (size-independent property)."""
    n, nq = 100_000, 512
    base = sch.random_descs(500, seed=51)
    rng = np.random.default_rng(52)
    c = gpu_lib.context(0, n_scans=16, max_points=1 << 15, max_map_points=1 << 16)
    try:
        c.scdb_reserve(n)
        for lo in range(0, n, 10_000):
            src = rng.integers(0, 500, size=10_000)
            block = base[src] * rng.uniform(0.7, 1.3, size=(10_000, 1, 1)).astype(np.float32)
            block = (block + (block > 0) * rng.normal(0, 0.2, size=block.shape)).astype(np.float32)
            c.scdb_add(block, sch.keys_of_fast(block))
            if lo == 30_000:
                keep = block.copy()
        assert c.scdb_size() == n
        ids = rng.integers(0, 10_000, size=nq)
        shifts = rng.integers(0, 60, size=nq)
        qd = np.stack([np.roll(keep[i], s, axis=1) for i, s in zip(ids, shifts)]).astype(np.float32)
        qk = sch.keys_of_fast(qd)
        lid, dist, sh = c.scdb_search(qk, qd, n)
        assert np.array_equal(lid, 30_000 + ids)
        assert np.all(np.abs(dist) < 1e-12)
        assert np.array_equal(sh, shifts)
        lid2, _, _ = c.scdb_search(qk, qd, 30_000)          # the revisited stretch is outside the searched prefix
        assert not np.any(lid2 == 30_000 + ids)
    finally:
        c.close()


def test_sc_golden_fixture_on_gpu(gctx, sco, synth):
    """The CUDA path against the committed fixture (no oracle call in the comparison)."""
    import os
    from make_golden import sc_golden_inputs, small_sweep
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sc_small.npz"))
    _, _, db, dbk, qid, qd, qk = sc_golden_inputs(synth, sco)
    for k in range(6):
        d, key = gctx.sc_make(small_sweep(synth, 2 * k))
        assert np.array_equal(u32(d), u32(g["desc"][k])) and np.array_equal(u32(key), u32(g["key"][k]))
    gctx.scdb_add(db, dbk)
    kidx, kd = gctx.scdb_knn(qk, 250)
    assert np.array_equal(kidx, g["knn_idx"]) and np.array_equal(kd.view(np.uint32), g["knn_d"].view(np.uint32))
    lid, ldist, lsh = gctx.scdb_search(qk, qd, 250)
    assert np.array_equal(lid, g["loop_id"]) and np.array_equal(ldist, g["loop_dist"]) and np.array_equal(lsh, g["loop_shift"])
    pd, ps = gctx.sc_distance(qd, db[qid])
    assert np.array_equal(pd, g["pair_dist"]) and np.array_equal(ps, g["pair_shift"])
    _, e0, f0 = gctx.extract_features(small_sweep(synth, 0))
    vox, _ = gctx.voxel_downsample(f0, 0.4)
    gctx.map_set(1, vox)
    _, _, f1 = gctx.extract_features(small_sweep(synth, 3))
    sc, ov, ni = gctx.align_score(1, f1[::4], np.eye(4), 1.0, 0.3)
    assert ov == g["align"][1] and ni == g["align"][2] and abs(sc - g["align"][0]) <= 1e-12 * g["align"][0]
