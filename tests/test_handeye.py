"""Row f3: multi-LiDAR extrinsic initialisation (hand-eye) — the host arithmetic of calib.cu against an independent
numpy restatement of handeye_calibration_base.hpp (np.linalg.svd / lstsq play Eigen's JacobiSVD), no GPU needed —
and the calibration branch of MultiLidarSystem::process() end to end on the GPU."""
import numpy as np
import pytest


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def to_pose(R, t, synth):
    return synth.pose_to_qt(R, np.asarray(t, float))


def make_pairs(synth, n, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    Rx, tx = rot([0.2, -0.1, 1.0], np.radians(38.0)), np.array([0.03, -0.54, -0.14])
    X = np.eye(4)
    X[:3, :3], X[:3, 3] = Rx, tx
    pri, sub = [], []
    for _ in range(n):
        A = np.eye(4)
        A[:3, :3] = rot(rng.normal(size=3), rng.uniform(0.05, 0.25))
        A[:3, 3] = rng.normal(0, 0.1, size=3)
        B = np.linalg.inv(X) @ A @ X
        if noise > 0:
            B[:3, 3] += rng.normal(0, noise, size=3)
        pri.append(to_pose(A[:3, :3], A[:3, 3], synth))
        sub.append(to_pose(B[:3, :3], B[:3, 3], synth))
    return X, pri, sub


def numpy_handeye(pri, sub):
    """CalibExRotation with numpy's SVD."""
    def L(q):
        x, y, z, w = q
        return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])

    def Rm(q):
        x, y, z, w = q
        return np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])

    Q = np.zeros((1200, 4))
    for i, (p, s) in enumerate(zip(pri, sub)):
        Q[4 * i:4 * i + 4] = L(p[:4]) - Rm(s[:4])
    _, sv, Vt = np.linalg.svd(Q)
    x = Vt[3]
    if x[0] < 0:
        x = -x
    q = np.array([x[1], x[2], x[3], x[0]])
    q /= np.linalg.norm(q)
    return q, sv


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_handeye_matches_numpy_and_recovers_extrinsic(gpu_lib, synth):
    from lmsf_slam_b200 import capi
    X, pri, sub = make_pairs(synth, 40, seed=5)
    h = capi.HandEye(gpu_lib)
    enough = [h.add_pose(p, s) for p, s in zip(pri, sub)]
    assert enough[:2] == [False, False] and all(enough[2:]) and h.size() == 40
    ok, ext, sv = h.calibrate()
    q_np, sv_np = numpy_handeye(pri, sub)
    assert ok and sv[2] > 0.25
    assert np.allclose(sv, sv_np, rtol=1e-9, atol=1e-9)
    assert min(np.linalg.norm(ext[:4] - q_np), np.linalg.norm(ext[:4] + q_np)) < 1e-9
    # translation: least squares of (R_A - I) t = R_X t_B - t_A
    A = np.concatenate([quat_to_R(p[:4]) - np.eye(3) for p in pri])
    b = np.concatenate([quat_to_R(ext[:4]) @ s[4:] - p[4:] for p, s in zip(pri, sub)])
    t_np = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.allclose(ext[4:], t_np, atol=1e-9)
    # and both equal the extrinsic the pairs were generated with
    assert np.allclose(quat_to_R(ext[:4]), X[:3, :3], atol=1e-9) and np.allclose(ext[4:], X[:3, 3], atol=1e-8)
    h.close()


def test_handeye_screening_and_degeneracy(gpu_lib, synth):
    from lmsf_slam_b200 import capi
    X, pri, sub = make_pairs(synth, 6, seed=9)
    h = capi.HandEye(gpu_lib)
    bad = np.array(sub[0]).copy()
    bad[4:] += [0.5, 0.5, 0.5]                        # translation along the screw axis disagrees: rejected (EPSILON_T)
    assert not h.add_pose(pri[0], bad) and h.size() == 0
    twisted = to_pose(rot([0, 0, 1], 0.6), [0, 0, 0], synth)
    assert not h.add_pose(pri[0], twisted) and h.size() == 0   # rotation angles differ by more than EPSILON_R
    # pure yaw motions: the rotation axis is never excited in two directions -> second smallest singular value stays small
    Rx = rot([0.2, -0.1, 1.0], np.radians(38.0))
    Xm = np.eye(4)
    Xm[:3, :3] = Rx
    for k in range(5):
        A = np.eye(4)
        A[:3, :3] = rot([0, 0, 1], 0.02)
        B = np.linalg.inv(Xm) @ A @ Xm
        h.add_pose(to_pose(A[:3, :3], A[:3, 3], synth), to_pose(B[:3, :3], B[:3, 3], synth))
    ok, _, sv = h.calibrate()
    assert not ok and sv[2] <= 0.25
    h.close()


def test_handeye_window_of_300_replaces_smallest_rotation(gpu_lib, synth):
    from lmsf_slam_b200 import capi
    _, pri, sub = make_pairs(synth, 320, seed=13)
    h = capi.HandEye(gpu_lib)
    for p, s in zip(pri, sub):
        h.add_pose(p, s)
    assert h.size() == 300
    ok, ext, sv = h.calibrate()
    assert ok
    h.close()


def _ref_handeye():
    """oracle/_ref/libref_handeye.so: the reference's own HandEyeCalibrationBase compiled from its unmodified header."""
    import ctypes as C
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_handeye.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_handeye.so not built (no /root/reference at build time)")
    dll = C.CDLL(path)
    dll.ref_handeye_create.restype = C.c_void_p
    dll.ref_handeye_destroy.argtypes = [C.c_void_p]
    f64p = C.POINTER(C.c_double)
    dll.ref_handeye_add_pose.argtypes = [C.c_void_p, f64p, f64p]
    dll.ref_handeye_calibrate.argtypes = [C.c_void_p, f64p, f64p]
    return dll, f64p


def test_handeye_vs_reference_code(gpu_lib, synth):
    """Row f3 pinned: lmsf_handeye_* against the reference's OWN HandEyeCalibrationBase (handeye_calibration_base.hpp:36-246,
    compiled where it lies; Eigen's JacobiSVD / AngleAxisd answered by restatements in oracle/shim_fixed, a one-sided
    Jacobi SVD — not the product's eigen-decomposition of Q^T Q).  One long stream of motion pairs with everything the
    class reacts to: pairs rejected by the rotation gate and by the translation gate (which also drop the accumulated
    motion), sub-threshold motions, more than 300 accepted pairs (the stored pair with the smallest rotation is replaced),
    a degenerate start (planar rotations: rot_cov[2] <= 0.25 refuses) and calibrations at several points.  Same accept /
    enough flag after every pair, same refusals, extrinsics equal to 1e-9."""
    import ctypes as C
    from lmsf_slam_b200 import capi
    dll, f64p = _ref_handeye()
    rng = np.random.default_rng(21)
    Rx, tx = rot([0.2, -0.1, 1.0], np.radians(38.0)), np.array([0.03, -0.54, -0.14])
    X = np.eye(4)
    X[:3, :3], X[:3, 3] = Rx, tx

    def pair(axis, ang, t, noise=0.0):
        A = np.eye(4)
        A[:3, :3], A[:3, 3] = rot(axis, ang), t
        B = np.linalg.inv(X) @ A @ X
        B[:3, 3] += rng.normal(0, noise, size=3) if noise > 0 else 0.0
        return to_pose(A[:3, :3], A[:3, 3], synth), to_pose(B[:3, :3], B[:3, 3], synth)

    stream = []
    for k in range(6):                       # planar start: yaw only
        stream.append(pair([0, 0, 1], 0.03 + 0.01 * k, [0.1, 0.0, 0.0]))
    stream.append("calibrate")               # degenerate: refused by both
    for k in range(340):
        p, s = pair(rng.normal(size=3), rng.uniform(0.01, 0.3), rng.normal(0, 0.1, size=3), noise=0.002)
        if k % 37 == 5:                      # rotation gate
            s = np.array(s)
            s[:4] = to_pose(rot(rng.normal(size=3), 0.5), [0, 0, 0], synth)[:4]
        if k % 41 == 7:                      # translation gate
            s = np.array(s)
            s[4:] += [0.4, -0.4, 0.4]
        stream.append((p, s))
        if k in (2, 40, 299, 339):
            stream.append("calibrate")

    h = capi.HandEye(gpu_lib)
    r = dll.ref_handeye_create()
    n_cal = n_ok = n_rej = 0
    try:
        for item in stream:
            if item == "calibrate":
                ok, ext, sv = h.calibrate()
                R9, t3 = np.zeros(9), np.zeros(3)
                ok_ref = dll.ref_handeye_calibrate(r, R9.ctypes.data_as(f64p), t3.ctypes.data_as(f64p))
                assert bool(ok_ref) == ok, (n_cal, sv)
                n_cal += 1
                if ok:
                    n_ok += 1
                    assert np.allclose(quat_to_R(ext[:4]), R9.reshape(3, 3), rtol=0, atol=1e-9)
                    assert np.allclose(ext[4:], t3, rtol=0, atol=1e-9)
                continue
            p, s = (np.ascontiguousarray(v, dtype=np.float64) for v in item)
            e_ref = dll.ref_handeye_add_pose(r, p.ctypes.data_as(f64p), s.ctypes.data_as(f64p))
            e = h.add_pose(p, s)
            assert bool(e_ref) == e
            n_rej += 0 if e else 1
        assert n_cal == 5 and n_ok >= 3 and n_rej >= 15 and h.size() == 300
    finally:
        h.close()
        dll.ref_handeye_destroy(r)


@pytest.mark.gpu
def test_rig_calibrates_then_refines(gpu_lib, synth):
    """MultiLidarSystem::process() calibration branch on two HDL-64 (the 64-line sensor tracks accurately in the synthetic
    room): status 0 until the hand-eye initialisation succeeds, then status 1 keeps the extrinsic close to the truth."""
    from lmsf_slam_b200 import rig
    sensor = synth.hdl64()
    Re, te = rot([0, 0, 1], np.radians(40.0)), np.array([0.03, -0.54, -0.14])
    r = rig.MultiLidarRig(gpu_lib, devices=(0, 0), n_scans=64)
    try:
        switched, e0 = None, None
        for k in range(36):
            yaw, pitch, roll = 0.09 * k, 0.25 * np.sin(0.45 * k), 0.25 * np.sin(0.33 * k + 1.0)
            Rk = synth.rot_zyx(yaw, pitch, roll)
            tk = np.array([-10.0 + 0.12 * k, 0.6 * np.sin(0.3 * k), 0.0])
            s0 = synth.make_sweep(sensor, k, pose=(Rk, tk))
            s1 = synth.make_sweep(sensor, k, pose=(Rk, tk), extrinsic=(Re, te))
            st = r.process([s0, s1], 0.1 * k)
            if st == 1 and switched is None:
                switched = k
                e0 = r.extrinsic.copy()
        assert switched is not None and switched < 30, r.singular_values
        for e in (e0, r.extrinsic):
            R = quat_to_R(e[:4])
            ang = np.degrees(np.arccos(np.clip((np.trace(R.T @ Re) - 1) / 2, -1, 1)))
            assert ang < 1.0 and np.linalg.norm(e[4:] - te) < 0.08, (ang, e[4:] - te)
        Rf = quat_to_R(r.extrinsic[:4])
        assert np.degrees(np.arccos(np.clip((np.trace(Rf.T @ Re) - 1) / 2, -1, 1))) < 0.2   # refined against the map
        assert np.linalg.norm(r.extrinsic[4:] - te) < 0.03
    finally:
        r.close()


def test_rig_pose_algebra(synth):
    """rig.pose_mul / pose_inv (host side of the calibration loop) against 4x4 matrices."""
    from lmsf_slam_b200 import rig
    rng = np.random.default_rng(3)
    for _ in range(20):
        A, B = np.eye(4), np.eye(4)
        A[:3, :3], A[:3, 3] = rot(rng.normal(size=3), rng.uniform(0, 3.0)), rng.normal(size=3)
        B[:3, :3], B[:3, 3] = rot(rng.normal(size=3), rng.uniform(0, 3.0)), rng.normal(size=3)
        pa, pb = to_pose(A[:3, :3], A[:3, 3], synth), to_pose(B[:3, :3], B[:3, 3], synth)
        assert np.allclose(synth.qt_to_mat(rig.pose_mul(pa, pb)), A @ B, atol=1e-12)
        assert np.allclose(synth.qt_to_mat(rig.pose_inv(pa)), np.linalg.inv(A), atol=1e-12)
        assert np.allclose(synth.qt_to_mat(rig.pose_mul(rig.pose_inv(pa), pa)), np.eye(4), atol=1e-12)
