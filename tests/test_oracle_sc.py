"""CPU tests of the loop-closure descriptor oracle (oracle/lmsf_oracle_sc.cpp) — pins it against
(i) the reference's own ScanContext class and its own ring-key KD-tree, both compiled from the reference tree (oracle/_ref),
(ii) an independent numpy transcription of Scancontext.hpp, (iii) domain properties — and checks the
host logic of the product path that needs no GPU (tree-rebuild schedule, record layout)."""
import ctypes as C

import numpy as np
import pytest

import ref_pin
import sc_helpers as sch


@pytest.fixture(scope="module")
def sco(oracle_lib):
    return sch.ScOracle()


@pytest.mark.parametrize("name,k", [("vlp16", 0), ("vlp16", 5), ("hdl64", 3)])
def test_sc_make_vs_numpy_transcription(sco, sweeps, name, k):
    sw = sweeps(name, k)
    d_o, k_o = sco.make(sw)
    d_p, k_p = sch.py_make_sc(sw)
    assert np.array_equal(d_o.view(np.uint32), d_p.view(np.uint32))
    assert np.array_equal(k_o.view(np.uint32), k_p.view(np.uint32))
    assert (d_o > 0).sum() > 100      # the room is seen


def test_sc_make_edge_cases(sco):
    d, k = sco.make(np.zeros((0, 4), np.float32))
    assert not d.any() and not k.any()
    pts = np.array([
        [1.0, 0.0, 0.5, 0],        # ring 1, sector 1 (theta = 0 -> ceil(0) = 0 -> clamped to 1)
        [0.0, 0.0, 9.0, 0],        # x = y = 0: atan(0/0) = NaN -> sector 1, ring 1
        [79.9, 0.0, 1.0, 0],       # ring 20
        [80.5, 0.0, 5.0, 0],       # beyond PC_MAX_RADIUS_: skipped
        [-3.0, -3.0, -1.5, 0],     # third quadrant
        [np.nan, 1.0, 3.0, 0],     # NaN range: not skipped, lands in ring 1 / sector 1
        [2.0, 2.0, np.nan, 0],     # NaN height never wins the max
        [5.0, -5.0, -1003.0, 0],   # below NO_POINT: never recorded, the bin ends as 0
    ], np.float32)
    d, k = sco.make(pts)
    dp, kp = sch.py_make_sc(pts)
    assert np.array_equal(d.view(np.uint32), dp.view(np.uint32)) and np.array_equal(k.view(np.uint32), kp.view(np.uint32))
    assert d[0, 0] == np.float32(11.0)           # max(0.5, 9, 3) + 2
    assert d[19, 0] == np.float32(3.0)
    assert np.count_nonzero(d) == 3


@pytest.fixture(scope="module")
def ref_sc():
    return ref_pin.RefSc()


def test_sc_make_vs_reference_code(sco, sweeps, ref_sc):
    """Row f1 PINNED (descriptor + ring key): the oracle against the reference's own MakeScanContext /
    MakeRingkeyFromScanContext (Scancontext.hpp:59-126), bit for bit, on full sweeps of both sensors and on the
    defined edge cases (tests/ref_pin.py)."""
    ref_pin.check_sc_make_against_reference(sco.make, sweeps, ref_sc)


def test_sc_distance_vs_reference_code(sco, sweeps, ref_sc):
    """Row f1 PINNED (control flow of DistanceBtnScanContext, Scancontext.hpp:133-172, 213-263: sector keys, 60-shift
    alignment, the seven refined shifts, first strict minimum): identical shift and bit-identical distance.  Eigen's
    mean / norm / dot are sequential sums on both sides (oracle/shim/Eigen/Dense), the one documented divergence."""
    real = [sco.make(sweeps(name, k))[0] for name, k in (("vlp16", 0), ("vlp16", 30), ("hdl64", 3), ("hdl64", 60))]
    synthetic = list(sch.random_descs(12, seed=9))
    descs = real + synthetic + [np.zeros((sch.NR, sch.NS), np.float32)]
    rng = np.random.default_rng(2)
    n = 0
    for i, a in enumerate(descs):
        for j, b in enumerate(descs):
            if (i + j) % 3 == 2:
                continue
            b = np.roll(b, int(rng.integers(0, sch.NS)), axis=1)
            d_r, s_r = ref_sc.distance(a, b)
            d_o, s_o = sco.distance(a, b)
            assert s_r == int(s_o[0]), (i, j)
            assert np.array([d_r], np.float64).view(np.uint64)[0] == d_o[:1].view(np.uint64)[0], (i, j, d_r, d_o[0])
            n += 1
    assert n > 150


def test_loop_search_vs_reference_code(sco, synth, ref_sc, gpu_lib):
    """Row f1 PINNED end to end: the reference's own SceneRecognitionScanContext (AddKeyFramePoints + LoopDetect,
    SceneRecognitionScanContext.hpp:61-124, 260-333) fed 140 keyframe clouds — 70 places, then the same 70 revisited
    under random yaw with fresh noise — against the oracle's search over the prefix the product's lmsf_sc_tree_limit
    names: identical loop id (or -1) and identical yaw (= column shift) for every keyframe from the 51st on."""
    import math
    dll = ref_sc.dll
    dll.ref_scdb_create.restype = C.c_void_p
    h = C.c_void_p(dll.ref_scdb_create())
    tree_limit = gpu_lib.fn("sc_tree_limit")
    tree_limit.restype = C.c_int
    rng = np.random.default_rng(8)
    sensor = synth.vlp16()

    def yawed(cloud, deg):
        c, s_ = math.cos(math.radians(deg)), math.sin(math.radians(deg))
        out = cloud.copy()
        out[:, 0] = c * cloud[:, 0] - s_ * cloud[:, 1]
        out[:, 1] = s_ * cloud[:, 0] + c * cloud[:, 1]
        return out.astype(np.float32)

    places = [synth.make_sweep(sensor, 3 * i)[::4] for i in range(70)]
    clouds = places + [yawed(p + rng.normal(0, 0.01, p.shape).astype(np.float32), float(rng.uniform(0, 360)))
                       for p in places]
    keys, descs, n_loop = [], [], 0
    fp = C.POINTER(C.c_float)
    for i, cl in enumerate(clouds):
        cl = np.ascontiguousarray(cl, np.float32)
        assert dll.ref_scdb_add(h, cl.ctypes.data_as(fp), len(cl)) == 0
        d, k = sco.make(cl)
        descs.append(d.reshape(-1))
        keys.append(k)
        if i + 1 < 51:
            continue                                   # LoopDetect returns -1 below NUM_EXCLUDE_RECENT_ + 1 keyframes
        lid, yaw = C.c_longlong(0), C.c_double(0)
        assert dll.ref_scdb_loop_detect(h, i, C.byref(lid), C.byref(yaw)) == 0
        o_id, _, o_sh = sco.search(np.array(keys), np.array(descs), tree_limit(i + 1), keys[i], descs[i])
        # float deg2rad(float degrees) { return degrees * M_PI / 180.0; } on nn_align * PC_UNIT_SECTORANGLE_ (:325, :344)
        o_yaw = float(np.float32(float(np.float32(o_sh[0] * 6.0)) * math.pi / 180.0)) if o_id[0] >= 0 else 0.0
        assert (lid.value, yaw.value) == (int(o_id[0]), o_yaw), (i, lid.value, yaw.value, o_id[0], o_sh[0])
        n_loop += lid.value >= 0
    dll.ref_scdb_destroy(h)
    assert n_loop > 60                                 # the revisits are found, not merely "both say -1"


def test_ringkey_knn_vs_reference_kdtree(sco):
    """Oracle brute force == the reference's own KDTreeVectorOfVectorsAdaptor<.., float> (leaf 10, metric_L2)."""
    rng = np.random.default_rng(11)
    keys = rng.uniform(0, 5, size=(4000, 20)).astype(np.float32)
    q = np.concatenate([keys[::97] + rng.normal(0, 0.05, size=(len(keys[::97]), 20)).astype(np.float32),
                        rng.uniform(0, 5, size=(40, 20)).astype(np.float32)])
    ref = sch.ref_ringkey_knn10(keys, q)
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    idx_r, d_r = ref
    idx_o, d_o = sco.knn(keys, len(keys), q)
    assert np.array_equal(d_r.view(np.uint32), d_o.view(np.uint32))   # same fp32 accumulation, bit for bit
    assert np.array_equal(idx_r, idx_o)


def test_ringkey_knn_small_and_ties(sco):
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 3, size=(64, 20)).astype(np.float32)       # many exact ties
    q = keys[:8].copy()
    idx, d = sco.knn(keys, 64, q)
    for i in range(8):
        assert d[i, 0] == 0.0
        assert np.all(np.diff(d[i]) >= 0)
        for j in range(9):
            if d[i, j] == d[i, j + 1]:
                assert idx[i, j] < idx[i, j + 1]
    idx, d = sco.knn(keys, 4, q)                                       # fewer than ten keys searchable
    assert np.all(idx[:, 4:] == -1) and np.all(np.isinf(d[:, 4:]))
    assert np.all(np.sort(idx[:, :4], axis=1) == np.arange(4))


def test_sc_distance_vs_numpy_and_properties(sco):
    descs = sch.random_descs(12, seed=3)
    rng = np.random.default_rng(4)
    for i in range(6):
        s = int(rng.integers(0, 60))
        a = descs[i]
        b = np.roll(a, -s, axis=1)                  # b shifted right by s gives a back
        d, sh = sco.distance(a, b)
        assert abs(d[0]) < 1e-12 and sh[0] == s
        d2, sh2 = sco.distance(a, descs[i + 6])
        dp, sp = sch.py_sc_distance(a, descs[i + 6])
        assert d2[0] == dp and sh2[0] == sp         # same operation order -> same bits
        assert 0.0 < d2[0] <= 1.0
    z = np.zeros((20, 60), np.float32)
    d, sh = sco.distance(z, descs[0])               # no effective column: 0/0 -> NaN never beats the initial minimum
    assert d[0] == 10000000.0 and sh[0] == 0


def test_search_finds_rotated_revisits(sco):
    descs = sch.random_descs(300, seed=8)
    keys = sch.keys_of_fast(descs)
    assert np.array_equal(keys[:5].view(np.uint32), sch.keys_of(descs[:5]).view(np.uint32))
    ids = np.array([3, 77, 150, 249])
    shifts = np.array([0, 7, 31, 59])
    qd = np.stack([np.roll(descs[i], s, axis=1) for i, s in zip(ids, shifts)])
    qk = sch.keys_of_fast(qd)
    lid, dist, sh = sco.search(keys, descs, 250, qk, qd)
    assert list(lid) == [3, 77, 150, 249]
    assert np.all(dist < 1e-12)
    assert list(sh) == [0, 7, 31, 59]               # query = circshift(candidate, s)  ->  argmin shift s
    lid, _, _ = sco.search(keys, descs, 100, qk, qd)
    assert list(lid) == [3, 77, -1, -1]             # ids beyond the searched prefix are invisible
    far = sch.random_descs(3, seed=99)
    lid, dist, _ = sco.search(keys, descs, 250, sch.keys_of_fast(far), far)
    assert np.all(lid == -1) and np.all(dist >= 0.2)


def test_tree_limit_schedule(gpu_lib):
    """lmsf_sc_tree_limit == the size of polarcontext_ringkeys_to_search_ after n AddKeyFramePoints calls
    (SceneRecognitionScanContext.hpp:74-92), simulated literally."""
    f = gpu_lib.fn("sc_tree_limit")
    f.restype = C.c_int
    size, counter, limit = 0, 0, 0
    for n in range(1, 400):
        size += 1
        if counter % 10 == 0 and size > 50:
            limit = size - 50
        counter += 1
        assert f(n) == limit, n


def test_candidate_record_layout():
    assert sch.CAND_DTYPE.itemsize == 24
    assert [sch.CAND_DTYPE.fields[k][1] for k in ("sc_dist", "key_dist", "id", "shift")] == [0, 8, 12, 16]


def test_sc_golden_fixture_reproduced(sco, oracle_lib, synth):
    """tests/golden/sc_small.npz (tests/make_golden.py): descriptors, ring-key top-10, SC distances, loop decisions and
    one alignment score — guards the oracle against drift between rounds."""
    import os
    from make_golden import sc_golden_inputs, small_sweep
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sc_small.npz"))
    descs, keys0, db, dbk, qid, qd, qk = sc_golden_inputs(synth, sco)
    assert np.array_equal(descs.view(np.uint32), g["desc"].view(np.uint32))
    assert np.array_equal(keys0.view(np.uint32), g["key"].view(np.uint32))
    kidx, kd = sco.knn(dbk, 250, qk)
    assert np.array_equal(kidx, g["knn_idx"]) and np.array_equal(kd.view(np.uint32), g["knn_d"].view(np.uint32))
    lid, ldist, lsh = sco.search(dbk, db, 250, qk, qd)
    assert np.array_equal(lid, g["loop_id"]) and np.array_equal(ldist, g["loop_dist"]) and np.array_equal(lsh, g["loop_shift"])
    pd, ps = sco.distance(qd, db[qid])
    assert np.array_equal(pd, g["pair_dist"]) and np.array_equal(ps, g["pair_shift"])
    o = oracle_lib.context(0, n_scans=16, oracle_knn_mode=1)
    _, e0, f0 = o.extract_features(small_sweep(synth, 0))
    vox, _ = o.voxel_downsample(f0, 0.4)
    o.map_set(1, vox)
    _, _, f1 = o.extract_features(small_sweep(synth, 3))
    assert np.array_equal(np.array(o.align_score(1, f1[::4], np.eye(4), 1.0, 0.3)), g["align"])
    o.close()
