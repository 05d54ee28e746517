"""Shared harness of the "pinned against the reference's own code" tests: ctypes wrappers of oracle/_ref/libref_*.so
(the reference's unmodified headers compiled where they lie, oracle/Makefile) and the comparison loops, written once
and run twice — on the CPU with the oracle (tests/test_oracle*.py) and on the B200 with the CUDA library
(tests/test_zz_gpu_reference_pin.py).  `oracle/_ref` travels to the GPU box prebuilt; without it the tests skip."""
import ctypes as C
import os

import numpy as np
import pytest

import __graft_entry__ as entry

NR, NS = 20, 60


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _ref_path(name):
    path = os.path.join(os.path.dirname(entry.ORACLE_LIB), "_ref", name)
    if not os.path.exists(path):
        pytest.skip(f"oracle/_ref/{name} is not built (needs /root/reference at build time)")
    return path


class RefLoam:
    """oracle/_ref/libref_loam.so: LOAMFeatureProcessorBase / PointCloudCommonProcess compiled from the reference's
    unmodified headers where they lie (oracle/Makefile, oracle/ref_loam.cpp, PCL as a container: oracle/shim/)."""

    def __init__(self):
        self.dll = C.CDLL(_ref_path("libref_loam.so"))

    def extract(self, sw, n_scans, min_r=2.0, max_r=80.0, thresh=1.0, bad=True):
        sw = np.ascontiguousarray(sw, np.float32)
        n = len(sw)
        fp = C.POINTER(C.c_float)
        e, s = np.zeros((max(n, 1), 4), np.float32), np.zeros((max(n, 1), 4), np.float32)
        ne, ns = C.c_int(0), C.c_int(0)
        rc = self.dll.ref_loam_extract(sw.ctypes.data_as(fp), n, n_scans, C.c_float(min_r), C.c_float(max_r),
                                       C.c_float(thresh), int(bad), n, e.ctypes.data_as(fp), C.byref(ne),
                                       s.ctypes.data_as(fp), C.byref(ns))
        assert rc == 0
        return e[:ne.value], s[:ns.value]

    def common_process(self, sw, remove_nan, near, far):
        sw = np.ascontiguousarray(sw, np.float32)
        n = len(sw)
        fp = C.POINTER(C.c_float)
        out = np.zeros((max(n, 1), 4), np.float32)
        m = C.c_int(0)
        rc = self.dll.ref_common_process(sw.ctypes.data_as(fp), n, int(remove_nan), C.c_float(near), C.c_float(far), n,
                                         out.ctypes.data_as(fp), C.byref(m))
        assert rc == 0
        return out[:m.value]


class RefSc:
    """oracle/_ref/libref_sc.so: the reference's own ScanContext class compiled from its unmodified header where it
    lies (oracle/Makefile, oracle/ref_sc.cpp; PCL / Eigen as containers: oracle/shim/)."""

    def __init__(self):
        self.dll = C.CDLL(_ref_path("libref_sc.so"))

    def make(self, xyzi):
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        fp = C.POINTER(C.c_float)
        desc, key = np.zeros((NR, NS), np.float32), np.zeros(NR, np.float32)
        assert self.dll.ref_sc_make(a.ctypes.data_as(fp), len(a), desc.ctypes.data_as(fp), key.ctypes.data_as(fp)) == 0
        return desc, key

    def distance(self, a, b):
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        fp = C.POINTER(C.c_float)
        d, s = C.c_double(0), C.c_int(0)
        assert self.dll.ref_sc_distance(a.ctypes.data_as(fp), b.ctypes.data_as(fp), C.byref(d), C.byref(s)) == 0
        return d.value, s.value


def check_extract_against_reference(make_ctx, synth, ref_loam, include_32_line=True):
    """make_ctx(**params) -> a context with extract_features (oracle or CUDA library).  Edge and surf clouds must equal
    the reference's LOAMFeatureProcessorBase::Process bit for bit, order included.  Returns the features compared."""
    rng = np.random.default_rng(11)
    cases = []
    for k in (0, 1, 7, 50, 120):
        cases.append((16, synth.make_sweep(synth.vlp16(), k), {}))
    for k in (0, 1, 7, 50):
        cases.append((64, synth.make_sweep(synth.hdl64(), k), {}))
    sw = synth.make_sweep(synth.vlp16(), 3)
    cases.append((16, sw[rng.random(len(sw)) > 0.3], {}))                       # ragged rings
    cases.append((16, sw, dict(thresh=0.2)))
    cases.append((16, sw, dict(bad=False)))
    cases.append((16, sw, dict(min_r=5.0, max_r=30.0)))
    if include_32_line:
        cases.append((32, sw, {}))                                              # 32-line ring formula on the same rays
    h = synth.make_sweep(synth.hdl64(), 9)
    cases.append((64, h[rng.random(len(h)) > 0.5], dict(thresh=0.5)))
    cases.append((64, h[: 64 * 15], {}))                                        # < 20 points per ring
    cases.append((16, sw[: 16 * 40], {}))                                       # short rings: 30 candidates, sectors of 5
    cases.append((16, np.zeros((0, 4), np.float32), {}))
    n_feat = 0
    for n_scans, cloud, kw in cases:
        o = make_ctx(n_scans=n_scans, min_range=kw.get("min_r", 2.0), max_range=kw.get("max_r", 80.0),
                     edge_thresh=kw.get("thresh", 1.0), remove_bad_points=int(kw.get("bad", True)))
        _, oe, os_ = o.extract_features(cloud)
        re_, rs_ = ref_loam.extract(cloud, n_scans, **kw)
        assert (len(oe), len(os_)) == (len(re_), len(rs_)), (n_scans, len(cloud), kw)
        assert np.array_equal(bits(oe), bits(re_)) and np.array_equal(bits(os_), bits(rs_)), (n_scans, len(cloud), kw)
        n_feat += len(oe) + len(os_)
        o.close()
    return n_feat


def check_common_process_against_reference(ctx, synth, ref_loam):
    """removeNaN + DistanceFilter of PointCloudCommonProcess::Process (no VoxelGrid: PCL arithmetic is not available)."""
    sw = synth.make_sweep(synth.vlp16(), 2).copy()
    sw[::31, 0] = np.nan
    sw[5, 2] = np.inf
    sw[77, 1] = -np.inf
    for near, far in ((0.0, 0.0), (3.0, 25.5), (0.0, 10.0), (7.25, 7.5)):
        ref = ref_loam.common_process(sw, True, near, far)
        out = ctx.common_process(sw, True, 0.0, near, far)
        assert np.array_equal(bits(out), bits(ref)), (near, far)


def check_sc_make_against_reference(make, sweeps, ref_sc):
    """make(xyzi) -> (descriptor, ring key) (oracle or CUDA library) against the reference's MakeScanContext /
    MakeRingkeyFromScanContext, bit for bit (NaN x / y is undefined behaviour in the reference's xy2theta: left out)."""
    for name, k in (("vlp16", 0), ("vlp16", 5), ("vlp16", 77), ("hdl64", 3), ("hdl64", 41)):
        sw = sweeps(name, k)
        d_o, k_o = make(sw)
        d_r, k_r = ref_sc.make(sw)
        assert np.array_equal(bits(d_o), bits(d_r)), (name, k)
        assert np.array_equal(bits(k_o), bits(k_r)), (name, k)
    rng = np.random.default_rng(4)
    cloud = np.zeros((20000, 4), np.float32)
    cloud[:, :2] = rng.uniform(-95, 95, size=(20000, 2))      # all quadrants, some beyond the 80 m radius
    cloud[:, 2] = rng.uniform(-3, 6, size=20000)              # heights below the sensor too
    cloud[::7, 0] = 0.0                                       # on the y axis: atan(+-inf)
    cloud[::11, 1] = 0.0                                      # on the x axis
    cloud[::77, :2] = 0.0                                     # x = y = 0: atan(NaN)
    edge = np.array([[1.0, 0.0, 0.5, 0], [0.0, 0.0, 9.0, 0], [79.9, 0.0, 1.0, 0], [80.5, 0.0, 5.0, 0],
                     [-3.0, -3.0, -1.5, 0], [2.0, 2.0, np.nan, 0], [5.0, -5.0, -1003.0, 0]], np.float32)
    for pts in (cloud, edge, np.zeros((0, 4), np.float32)):
        d_o, k_o = make(pts)
        d_r, k_r = ref_sc.make(pts)
        assert np.array_equal(bits(d_o), bits(d_r))
        assert np.array_equal(bits(k_o), bits(k_r))


def check_lm_register_against_reference(ctx, synth, exact):
    """ctx: a context (oracle or CUDA library) with extract_features / map_set / set_lm_outer / register.  Twelve
    consecutive factory-default (Huber-LM) solves on ONE solver object against the reference's own
    CeresEdgeSurfFeatureRegistration (oracle/_ref/libref_lm.so; ceres::Solve answered by the oracle's restated loop).
    exact: translation bit-identical and rotation to the last bit (oracle); otherwise BASELINE.json's parity bar,
    1e-4 m and 1e-5 rad (CUDA library: its fp64 reduction order over the rows differs).  Returns the outer budgets."""
    from scipy.spatial.transform import Rotation
    C.CDLL(entry.ORACLE_LIB, mode=C.RTLD_GLOBAL)
    dll = C.CDLL(_ref_path("libref_lm.so"))
    dll.ref_lm_create.restype = C.c_void_p
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    sensor = synth.vlp16()
    _, me, ms = ctx.extract_features(synth.make_sweep(sensor, 0))
    ctx.map_set(0, me)
    ctx.map_set(1, ms)
    h = C.c_void_p(dll.ref_lm_create())
    for kind, m in ((0, me), (1, ms)):
        m = np.ascontiguousarray(m, np.float32)
        assert dll.ref_lm_set_map(h, kind, m.ctypes.data_as(fp), len(m)) == 0
    ctx.set_lm_outer(10)
    yaw = Rotation.from_euler("z", 0.4, degrees=True).as_matrix()
    priors = [(np.eye(3), np.zeros(3)), (np.eye(3), np.array([0.05, -0.03, 0.01])), (yaw, np.array([0.02, 0.0, 0.0]))]
    outers = []
    for k in range(12):
        R0, t0 = priors[k % 3]
        _, e, s = ctx.extract_features(synth.make_sweep(sensor, 1 + k % 3))
        e, s = np.ascontiguousarray(e), np.ascontiguousarray(s[::3])
        R, t, q0 = np.array(R0, np.float64).reshape(9).copy(), np.array(t0, np.float64).copy(), np.zeros(4)
        assert dll.ref_lm_solve(h, e.ctypes.data_as(fp), len(e), s.ctypes.data_as(fp), len(s), R.ctypes.data_as(dp),
                                t.ctypes.data_as(dp), q0.ctypes.data_as(dp)) == 0
        po, st = ctx.register(e, s, pose=np.concatenate([q0, t0]), solver=1)
        Ro = Rotation.from_quat(po[:4]).as_matrix()
        if exact:
            assert np.array_equal(t.view(np.uint64), po[4:].view(np.uint64)), (k, t, po[4:])
            assert np.abs(R.reshape(3, 3) - Ro).max() < 1e-15
        else:
            assert np.linalg.norm(t - po[4:]) < 1e-4, (k, t, po[4:])
            assert Rotation.from_matrix(R.reshape(3, 3) @ Ro.T).magnitude() < 1e-5, k
        assert np.linalg.norm(t - t0) > 1e-3                     # the solve moved the pose: the comparison is not vacuous
        outers.append(st["outer_iters"])
    dll.ref_lm_destroy(h)
    return outers


def check_tracker_against_reference(ctx, synth, exact, n_sweeps=14, window=3):
    """ctx: a context (oracle or CUDA library; created with window=`window`, raw maps) with extract_features /
    tracker_step_features / get_map.  `n_sweeps` consecutive sweeps through the reference's OWN tracker
    (LidarTracker/LidarTrackerLocalMap.hpp compiled into oracle/_ref/libref_tracker.so, instantiated as its factory does,
    with its own CeresEdgeSurfFeatureRegistration; ceres::Solve = the oracle's restated loop, the sliding window = the
    inferred stand-in of oracle/shim_inferred) and through ctx: pose, motion increment, keyframe decisions (seen through
    the local-map sizes) and the local maps themselves, every sweep.  The sequence holds the constant-motion prediction,
    a caller-supplied prediction, motion keyframes, a window that fills and evicts, and a time keyframe (a 12 s gap).
    exact: translations bit-identical, rotations to the last bit, maps bit-identical (oracle); otherwise BASELINE.json's
    bar of 1e-4 m / 1e-5 rad and maps within 1e-4 m (CUDA library).  Returns the keyframe kinds ctx reported."""
    from scipy.spatial.transform import Rotation
    C.CDLL(entry.ORACLE_LIB, mode=C.RTLD_GLOBAL)
    dll = C.CDLL(_ref_path("libref_tracker.so"))
    dll.ref_tracker_create.restype = C.c_void_p
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    h = C.c_void_p(dll.ref_tracker_create(window))
    sensor = synth.vlp16()
    kinds = []
    stamp = 0.0
    for k in range(n_sweeps):
        stamp += 12.0 if k == 9 else 0.1                    # sweep 9 arrives after TIME_INTERVAL_ = 10 s
        _, e, s = ctx.extract_features(synth.make_sweep(sensor, k))
        e, s = np.ascontiguousarray(e), np.ascontiguousarray(s[::2])     # thinned: the CPU side runs both trackers
        # sweep 5: the caller predicts the motion itself (deltaT != Identity, LidarTrackerLocalMap.hpp:125-129) — a pure
        # translation, so that both sides receive exactly the same transform (no quaternion -> matrix conversion between)
        delta = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
        if k == 5:
            delta[4:] = [0.14, 0.01, 0.0]
        dR = np.eye(3).reshape(9).copy()
        dt = delta[4:].copy()
        pR, pt, nm = np.zeros(9), np.zeros(3), np.zeros(2, np.int32)
        assert dll.ref_tracker_solve(h, e.ctypes.data_as(fp), len(e), s.ctypes.data_as(fp), len(s), C.c_double(stamp),
                                     dR.ctypes.data_as(dp), dt.ctypes.data_as(dp), pR.ctypes.data_as(dp),
                                     pt.ctypes.data_as(dp), nm.ctypes.data_as(ip)) == 0
        po, do, st = ctx.tracker_step_features(e, s, stamp, delta=delta)
        kinds.append(st["keyframe"])
        Ro = Rotation.from_quat(po[:4]).as_matrix()
        Rd = Rotation.from_quat(do[:4]).as_matrix()
        if os.environ.get("LMSF_PIN_VERBOSE"):
            print(k, "kind", st["keyframe"], "maps", st["map_edge"], st["map_surf"], list(nm), "dt", np.abs(pt - po[4:]).max(),
                  "dR", np.abs(pR.reshape(3, 3) - Ro).max(), "delta", np.abs(dt - do[4:]).max(), flush=True)
        if exact:
            assert np.array_equal(pt.view(np.uint64), po[4:].view(np.uint64)), (k, pt, po[4:])
            assert np.abs(pR.reshape(3, 3) - Ro).max() < 1e-15, k
            assert np.abs(dt - do[4:]).max() < 1e-15 and np.abs(dR.reshape(3, 3) - Rd).max() < 1e-15, k
        else:
            assert np.linalg.norm(pt - po[4:]) < 1e-4, (k, pt, po[4:])
            assert Rotation.from_matrix(pR.reshape(3, 3) @ Ro.T).magnitude() < 1e-5, k
        assert (st["map_edge"], st["map_surf"]) == (int(nm[0]), int(nm[1])), (k, st, nm)   # same keyframe decisions
        for kind in (0, 1):
            m_o = ctx.get_map(kind)
            m_r = np.zeros((max(int(nm[kind]), 1), 4), np.float32)
            n = dll.ref_tracker_map(h, kind, m_r.ctypes.data_as(fp), len(m_r))
            assert n == len(m_o), (k, kind)
            if exact:
                assert np.array_equal(m_r[:n].view(np.uint32), m_o.view(np.uint32)), (k, kind)
            else:
                assert np.abs(m_r[:n, :3] - m_o[:, :3]).max() < 1e-4, (k, kind)
    dll.ref_tracker_destroy(h)
    return kinds


def rotary_cases(synth):
    """Sweeps for the rotary preprocess: VLP-16 / HDL-64 as they arrive (a full turn starting at azimuth 0), a sweep that
    starts mid-turn and one that wraps (start / end angle logic, findStartEndAngle :80-94), NaN / inf points that
    removeNaN drops (among them the first and the last point), a partial sweep, tiny clouds."""
    cases = []
    for sensor, k in ((synth.vlp16(), 3), (synth.hdl64(), 7)):
        cases.append(synth.make_sweep(sensor, k))
    v = synth.make_sweep(synth.vlp16(), 5)
    n = len(v)
    cases.append(np.ascontiguousarray(np.roll(v, n // 3, axis=0)))            # starts a third of a turn in
    cases.append(np.ascontiguousarray(np.roll(v, -(n // 7) * 1, axis=0)))
    cases.append(np.ascontiguousarray(v[: n // 2]))                            # half a turn
    cases.append(np.ascontiguousarray(v[n // 5: n - n // 9]))
    w = v.copy()
    w[0, 0] = np.nan                                                           # the first point is dropped
    w[-1, 2] = np.inf                                                          # and the last
    w[::13, 1] = np.nan
    w[7, :3] = -np.inf
    cases.append(w)
    m = v[:, :].copy()
    m[:, 1] *= -1.0                                                            # a LiDAR spinning the other way
    cases.append(np.ascontiguousarray(m))
    cases.append(np.ascontiguousarray(v[:1]))
    cases.append(np.ascontiguousarray(v[:2]))
    cases.append(np.full((5, 4), np.nan, np.float32))
    return cases


def check_rotary_against_reference(pre, synth, periods=(0.1, 0.05)):
    """pre(xyzi, period) -> preprocessed cloud (oracle or CUDA library) against the reference's removeNaN +
    RotaryLidarPreProcess<PointXYZI>::Process (Preprocess/RotaryLidar_preprocessing.hpp:31-104, compiled into
    oracle/_ref/libref_loam.so), bit for bit, every case of rotary_cases."""
    dll = C.CDLL(_ref_path("libref_loam.so"))
    fp = C.POINTER(C.c_float)
    n_pts = 0
    for cloud in rotary_cases(synth):
        cloud = np.ascontiguousarray(cloud, np.float32)
        for period in periods:
            ref = np.zeros((max(len(cloud), 1), 4), np.float32)
            n = dll.ref_rotary_preprocess(cloud.ctypes.data_as(fp), len(cloud), C.c_double(period), len(ref),
                                          ref.ctypes.data_as(fp))
            assert n >= 0
            out = pre(cloud, period)
            assert len(out) == n, (len(cloud), period)
            assert np.array_equal(bits(out), bits(ref[:n])), (len(cloud), period)
            n_pts += n
    return n_pts
