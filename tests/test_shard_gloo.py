"""world_size-2 gloo tests (CPU) of the N > 1 host logic: unit assignment, max-over-ranks timing and the
sharded top-k merge used by the loop-closure descriptor search."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as entry


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    from lmsf_slam_b200 import shard
    out = {}
    out["assign"] = shard.assign(8, world, rank)
    out["max"] = shard.max_over_ranks([1.0 + rank, 10.0 - rank])
    out["sum"] = shard.sum_over_ranks([float(len(out["assign"]))])
    # sharded database search: every rank holds a slice of the keys, finds its local top-k, then merges
    g = torch.Generator().manual_seed(1234)
    db = torch.rand(1000, 20, generator=g)
    db[500] = db[3]                                   # an exact tie across shards
    queries = torch.cat([db[[3, 77]] + 0.0, torch.rand(6, 20, generator=g)])
    lo, hi = shard.shard_bounds(len(db), world, rank)
    d = torch.cdist(queries, db[lo:hi]) ** 2
    k = 10
    ld, li = torch.topk(d, k, dim=1, largest=False)
    md, mi = shard.merge_topk(ld, li + lo, k)
    out["merged"] = (md.numpy(), mi.numpy())
    out["bounds"] = (lo, hi)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_topk_merge():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["assign"] == [0, 2, 4, 6] and res[1]["assign"] == [1, 3, 5, 7]
    assert res[0]["max"] == res[1]["max"] == [2.0, 10.0]
    assert res[0]["sum"] == [8.0]
    assert res[0]["bounds"] == (0, 500) and res[1]["bounds"] == (500, 1000)
    # both ranks hold the same merged result, equal to a single-process brute force with (dist, id) order
    g = torch.Generator().manual_seed(1234)
    db = torch.rand(1000, 20, generator=g)
    db[500] = db[3]
    queries = torch.cat([db[[3, 77]] + 0.0, torch.rand(6, 20, generator=g)])
    d = (torch.cdist(queries, db[:500]) ** 2, torch.cdist(queries, db[500:]) ** 2)
    full = torch.cat(d, dim=1).numpy()
    order = np.lexsort((np.broadcast_to(np.arange(1000), full.shape), full), axis=1)[:, :10]
    for r in (0, 1):
        md, mi = res[r]["merged"]
        assert np.array_equal(mi, order)
        assert np.allclose(md, np.take_along_axis(full, order, 1))
    assert res[0]["merged"][1][0, 0] == 3 and res[0]["merged"][1][0, 1] == 500   # the tie: smaller id first


def test_single_process_paths():
    entry.load_package()
    from lmsf_slam_b200 import shard
    assert shard.assign(3, 1, 0) == [0, 1, 2]
    assert shard.shard_bounds(10, 3, 0) == (0, 4) and shard.shard_bounds(10, 3, 2) == (7, 10)
    assert shard.max_over_ranks([3.0]) == [3.0]
    d = torch.tensor([[0.3, 0.1, 0.2]])
    i = torch.tensor([[7, 9, 8]])
    md, mi = shard.merge_topk(d, i, 2)
    assert mi.tolist() == [[9, 8]]


# ---------------------------------------------------------------- sharded loop-closure search (config 5)
def _loop_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    entry.load_package()
    from lmsf_slam_b200 import shard
    import sc_helpers as sch

    sco = sch.ScOracle()
    n, nq, limit = 600, 24, 540
    descs = sch.random_descs(n, seed=71)
    keys = sch.keys_of_fast(descs)
    rng = np.random.default_rng(72)
    ids = rng.integers(0, n, size=nq)
    qd = np.stack([np.roll(descs[i], int(s), axis=1) for i, s in zip(ids, rng.integers(0, 60, size=nq))])
    qk = sch.keys_of_fast(qd)
    lo, hi = shard.shard_bounds(n, world, rank)
    lim_local = shard.shard_limit(limit, lo, hi)
    # what lmsf_scdb_search_shard_dev produces on this rank, computed here by the CPU oracle
    cand = np.zeros((nq, 10), sch.CAND_DTYPE)
    cand["id"] = -1
    cand["sc_dist"] = 10000000.0
    cand["key_dist"] = np.inf
    if lim_local > 0:
        idx, kd = sco.knn(keys[lo:hi], lim_local, qk)
        for qi in range(nq):
            for j in range(10):
                if idx[qi, j] < 0:
                    continue
                d, s = sco.distance(qd[qi], descs[lo + idx[qi, j]])
                cand[qi, j] = (d[0], kd[qi, j], lo + idx[qi, j], s[0], 0)
    block = torch.from_numpy(cand.view(np.uint8).reshape(nq, 10, 24).copy())
    allc = shard.gather_candidates(block)
    rec = allc.numpy().view(sch.CAND_DTYPE).reshape(world, nq, 10)
    lid, dd, sh = sch.py_pick(rec)
    # the two-round exchange (shard.loop_search_sharded_2round) with the kernels' work done by the CPU oracle: round 1
    # gathers unscored records, every rank scores the survivors it owns, round 2 gathers them, the same pick
    keys_only = cand.copy()
    keys_only["sc_dist"] = 10000000.0
    keys_only["shift"] = 0
    all_keys = shard.gather_candidates(torch.from_numpy(keys_only.view(np.uint8).reshape(nq, 10, 24).copy()))
    scored = sch.py_select_and_score(all_keys.numpy().view(sch.CAND_DTYPE).reshape(world, nq, 10), lo, lim_local,
                                     lambda qi, gid: tuple(x[0] for x in sco.distance(qd[qi], descs[gid])))
    all_scored = shard.gather_candidates(torch.from_numpy(scored.view(np.uint8).reshape(nq, 10, 24).copy()))
    lid2, dd2, sh2 = sch.py_pick(all_scored.numpy().view(sch.CAND_DTYPE).reshape(world, nq, 10))
    assert np.array_equal(lid2, lid) and np.array_equal(dd2, dd) and np.array_equal(sh2, sh)
    owned = int((scored["id"] >= 0).sum())
    q.put((rank, (lo, hi, lim_local, lid, dd, sh, owned)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_loop_search_equals_unsharded_oracle():
    import sc_helpers as sch

    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_loop_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][:3] == (0, 300, 300) and res[1][:3] == (300, 600, 240)
    sco = sch.ScOracle()
    n, nq, limit = 600, 24, 540
    descs = sch.random_descs(n, seed=71)
    keys = sch.keys_of_fast(descs)
    rng = np.random.default_rng(72)
    ids = rng.integers(0, n, size=nq)
    qd = np.stack([np.roll(descs[i], int(s), axis=1) for i, s in zip(ids, rng.integers(0, 60, size=nq))])
    qk = sch.keys_of_fast(qd)
    l_o, d_o, s_o = sco.search(keys, descs, limit, qk, qd)
    for r in (0, 1):
        _, _, _, lid, dd, sh, owned = res[r]
        assert np.array_equal(lid, l_o) and np.array_equal(dd, d_o) and np.array_equal(sh, s_o)
    # two rounds: every surviving candidate is scored exactly once, by its owner (10 per query in total, not per rank)
    assert res[0][6] + res[1][6] == nq * 10
    assert (l_o >= 0).sum() >= 15


def test_shard_limit():
    entry.load_package()
    from lmsf_slam_b200 import shard
    assert shard.shard_limit(540, 0, 300) == 300
    assert shard.shard_limit(540, 300, 600) == 240
    assert shard.shard_limit(100, 300, 600) == 0
