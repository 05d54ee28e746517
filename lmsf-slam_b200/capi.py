"""ctypes binding of the C ABI declared in include/lmsf_b200.h.

The binding is generic over (shared-library path, symbol prefix) because the CPU
oracle under oracle/ deliberately exports the same entry points with the prefix
``lmsf_oracle_`` (see the header under oracle/) so the parity tests can drive both through
one wrapper.  This package itself only ever loads the CUDA library
(``csrc/liblmsf_b200.so``); pointing the binding at the oracle is done by
tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() alone.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

KIND_EDGE, KIND_SURF = 0, 1
SC_RINGS, SC_SECTORS, SC_CELLS, SC_CANDIDATES = 20, 60, 1200, 10
SC_CAND_BYTES = 24       # sizeof(lmsf_sc_cand)
SC_DIST_THRES = 0.2      # SC_DIST_THRES_ (LoopDetection/SceneRecognitionScanContext.hpp)
SC_EXCLUDE_RECENT = 50   # NUM_EXCLUDE_RECENT_
SC_TREE_PERIOD = 10      # TREE_MAKING_PERIOD_
SOLVER_GN, SOLVER_HUBER_LM = 0, 1
N_STAGES = 7
STAGE_NAMES = ("extract", "match", "solve", "map", "voxel", "assoc", "fit")


class Params(C.Structure):
    """lmsf_params (include/lmsf_b200.h)."""

    _fields_ = [
        ("n_scans", C.c_int32),
        ("min_range", C.c_float),
        ("max_range", C.c_float),
        ("edge_thresh", C.c_float),
        ("remove_bad_points", C.c_int32),
        ("max_points", C.c_int32),
        ("window", C.c_int32),
        ("solver", C.c_int32),
        ("map_leaf_edge", C.c_float),
        ("map_leaf_surf", C.c_float),
        ("scan_leaf_edge", C.c_float),
        ("scan_leaf_surf", C.c_float),
        ("gn_max_iters", C.c_int32),
        ("lm_outer_start", C.c_int32),
        ("lm_inner_iters", C.c_int32),
        ("huber_delta", C.c_float),
        ("kf_trans", C.c_double),
        ("kf_rot", C.c_double),
        ("kf_time", C.c_double),
        ("max_map_points", C.c_int32),
        ("rotary_scan_period", C.c_float),
        ("reserved", C.c_int32 * 10),
    ]


class RegStats(C.Structure):
    _fields_ = [
        ("outer_iters", C.c_int32),
        ("n_edge_matched", C.c_int32),
        ("n_surf_matched", C.c_int32),
        ("converged", C.c_int32),
        ("degenerate", C.c_int32),
        ("lm_steps_total", C.c_int32),
        ("lm_steps_accepted", C.c_int32),
        ("pad", C.c_int32),
        ("final_cost", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad"}


class TrackStats(C.Structure):
    _fields_ = [
        ("n_edge", C.c_int32),
        ("n_surf", C.c_int32),
        ("keyframe", C.c_int32),
        ("map_edge", C.c_int32),
        ("map_surf", C.c_int32),
        ("first", C.c_int32),
        ("reg", RegStats),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reg"}
        d["reg"] = self.reg.as_dict()
        return d


class LmsfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lmsf error {code}: {msg}")
        self.code = code


_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_intp = C.POINTER(C.c_int)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _xyzi(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("expected an (n, 4) float32 XYZI array")
    return a


IDENTITY_POSE = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)


class Library:
    """One loaded shared library exporting the ABI under `prefix`."""

    # symbols every implementation of the ABI exports (name without prefix)
    COMMON = (
        "params_default", "ctx_create", "ctx_destroy", "strerror", "extract_features",
        "voxel_downsample", "map_set", "knn5", "match", "register", "set_lm_outer",
        "tracker_step", "tracker_step_features", "tracker_reset", "tracker_register_aux", "get_map", "align_score", "common_process",
    )
    # symbols only the CUDA library exports
    DEVICE_ONLY = (
        "last_cuda_error", "launch_count", "stream", "tracker_step_dev", "tracker_prefetch", "tracker_prefetch_dev",
        "tracker_submit", "tracker_submit_dev", "tracker_wait", "tracker_prefetch_cancel", "tracker_step_ticket",
        "tracker_submit_ticket",
        "dev_alloc", "dev_free", "extract_features_to_dev", "tracker_register_aux_features_dev",
        "dev_upload", "profile_enable", "profile_read",
        # loop-closure descriptor path (scancontext.cu)
        "sc_make", "sc_distance", "scdb_reserve", "scdb_clear", "scdb_size", "scdb_add", "scdb_add_cloud",
        "scdb_get", "scdb_knn", "scdb_search", "scdb_search_shard_dev", "scdb_pick_dev", "sc_tree_limit",
        "scdb_keys_shard_dev", "scdb_score_owned_dev",
        # multi-LiDAR extrinsic initialisation (calib.cu, host arithmetic)
        "handeye_create", "handeye_destroy", "handeye_add_pose", "handeye_calibrate", "handeye_size",
    )

    def __init__(self, path: str, prefix: str = "lmsf_", params_cls=Params):
        self.params_cls = params_cls   # a same-layout struct that names implementation-specific reserved words
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback behind this ABI.")
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL if False else C.DEFAULT_MODE)
        self.is_device = prefix == "lmsf_"
        f = self.fn
        f("strerror").restype = C.c_char_p
        f("ctx_destroy").restype = None
        if self.is_device:
            f("last_cuda_error").restype = C.c_char_p
            f("launch_count").restype = C.c_int64
            f("stream").restype = C.c_void_p
            f("tracker_step_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, _f64p, _f64p,
                                              C.POINTER(TrackStats)]
            f("tracker_prefetch").argtypes = [C.c_void_p, _f32p, C.c_int, C.POINTER(C.c_int64)]
            f("tracker_prefetch_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
            f("tracker_prefetch_cancel").argtypes = [C.c_void_p, C.c_int64]
            f("tracker_step_ticket").argtypes = [C.c_void_p, C.c_int64, C.c_double, _f64p, _f64p, C.POINTER(TrackStats)]
            f("tracker_submit_ticket").argtypes = [C.c_void_p, C.c_int64, C.c_double, _f64p]
            f("tracker_submit").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_double, _f64p]
            f("tracker_submit_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, _f64p]
            f("tracker_wait").argtypes = [C.c_void_p, _f64p, _f64p, C.POINTER(TrackStats)]
            f("extract_features_to_dev").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_void_p, C.c_int, _intp, _intp]
            f("tracker_register_aux_features_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _f64p,
                                                               C.POINTER(RegStats)]
            f("dev_alloc").argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
            f("dev_free").argtypes = [C.c_void_p, C.c_void_p]
            f("dev_upload").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
            f("launch_count").argtypes = [C.c_void_p]
            f("stream").argtypes = [C.c_void_p]
            f("last_cuda_error").argtypes = [C.c_void_p]
            f("profile_enable").argtypes = [C.c_void_p, C.c_int]
            f("profile_read").argtypes = [C.c_void_p, _f64p, C.POINTER(C.c_int64), _f64p, C.c_int]
            f("sc_make").argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, _f32p]
            f("sc_distance").argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, _f64p, _i32p]
            f("scdb_reserve").argtypes = [C.c_void_p, C.c_int]
            f("scdb_clear").argtypes = [C.c_void_p]
            f("scdb_size").argtypes = [C.c_void_p, _intp]
            f("scdb_add").argtypes = [C.c_void_p, _f32p, _f32p, C.c_int]
            f("scdb_add_cloud").argtypes = [C.c_void_p, _f32p, C.c_int, _intp]
            f("scdb_get").argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
            f("scdb_knn").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _i32p, _f32p]
            f("scdb_search").argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, C.c_int, C.c_double, _i32p, _f64p, _i32p]
            f("scdb_search_shard_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   C.c_void_p]
            f("scdb_pick_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
            f("scdb_keys_shard_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            f("scdb_score_owned_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                                  C.c_void_p]
            f("sc_tree_limit").argtypes = [C.c_int]
            f("handeye_create").argtypes = [C.POINTER(C.c_void_p)]
            f("handeye_destroy").argtypes = [C.c_void_p]
            f("handeye_destroy").restype = None
            f("handeye_add_pose").argtypes = [C.c_void_p, _f64p, _f64p, _intp]
            f("handeye_calibrate").argtypes = [C.c_void_p, _f64p, _f64p, _intp]
            f("handeye_size").argtypes = [C.c_void_p, _intp]
        f("ctx_create").argtypes = [C.c_int, C.POINTER(params_cls), C.POINTER(C.c_void_p)]
        f("ctx_destroy").argtypes = [C.c_void_p]
        f("extract_features").argtypes = [C.c_void_p, _f32p, C.c_int, _u8p, _f32p, _intp, _f32p, _intp]
        f("voxel_downsample").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, _f32p, _intp, _i32p]
        f("map_set").argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int]
        f("rotary_preprocess").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, _f32p, _intp]
        f("knn5").argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, _i32p, _f32p]
        f("match").argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, _u8p, _f64p]
        f("register").argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, C.c_int, C.c_int, _f64p, C.POINTER(RegStats)]
        f("set_lm_outer").argtypes = [C.c_void_p, C.c_int]
        f("tracker_step").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_double, _f64p, _f64p, C.POINTER(TrackStats)]
        f("tracker_step_features").argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, C.c_int, C.c_double, _f64p, _f64p,
                                               C.POINTER(TrackStats)]
        f("tracker_reset").argtypes = [C.c_void_p]
        f("tracker_register_aux").argtypes = [C.c_void_p, _f32p, C.c_int, _f64p, C.POINTER(RegStats)]
        f("get_map").argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, _intp]
        f("align_score").argtypes = [C.c_void_p, C.c_int, _f32p, C.c_int, _f32p, C.c_double, C.c_double, _f64p, _f64p,
                                     _i32p]
        f("common_process").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _f32p, _intp]

    def fn(self, name):
        return getattr(self.dll, self.prefix + name)

    def has(self, name) -> bool:
        try:
            self.fn(name)
            return True
        except AttributeError:
            return False

    def default_params(self) -> Params:
        p = self.params_cls()
        rc = self.fn("params_default")(C.byref(p))
        if rc:
            raise LmsfError(rc, "params_default")
        return p

    def context(self, device: int = 0, **overrides) -> "Context":
        return Context(self, device, **overrides)


class HandEye:
    """HandEyeCalibrationBase (handeye_calibration_base.hpp): extrinsic initialisation from paired motion increments."""

    def __init__(self, lib: "Library"):
        self.lib = lib
        self._h = C.c_void_p()
        rc = lib.fn("handeye_create")(C.byref(self._h))
        if rc:
            raise LmsfError(rc, "handeye_create")

    def close(self):
        if self._h:
            self.lib.fn("handeye_destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_pose(self, delta_primary, delta_sub) -> bool:
        a = np.array(delta_primary, dtype=np.float64)
        b = np.array(delta_sub, dtype=np.float64)
        ok = C.c_int(0)
        rc = self.lib.fn("handeye_add_pose")(self._h, a.ctypes.data_as(_f64p), b.ctypes.data_as(_f64p), C.byref(ok))
        if rc:
            raise LmsfError(rc, "handeye_add_pose")
        return bool(ok.value)

    def calibrate(self):
        """(ok, extrinsic {qx,qy,qz,qw,tx,ty,tz}, singular values descending)"""
        e = np.zeros(7, np.float64)
        sv = np.zeros(4, np.float64)
        ok = C.c_int(0)
        rc = self.lib.fn("handeye_calibrate")(self._h, e.ctypes.data_as(_f64p), sv.ctypes.data_as(_f64p), C.byref(ok))
        if rc:
            raise LmsfError(rc, "handeye_calibrate")
        return bool(ok.value), e, sv

    def size(self) -> int:
        n = C.c_int(0)
        self.lib.fn("handeye_size")(self._h, C.byref(n))
        return n.value


class Context:
    """One LiDAR's context (lmsf_ctx): stream, arenas, local-map index, tracker state."""

    def __init__(self, lib: Library, device: int = 0, **overrides):
        self.lib = lib
        self.params = lib.default_params()
        for k, v in overrides.items():
            if not hasattr(self.params, k):
                raise AttributeError(f"lmsf_params has no field {k}")
            setattr(self.params, k, v)
        self._h = C.c_void_p()
        rc = lib.fn("ctx_create")(device, C.byref(self.params), C.byref(self._h))
        if rc:
            raise LmsfError(rc, lib.fn("strerror")(rc).decode())

    def close(self):
        if self._h:
            self.lib.fn("ctx_destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc:
            msg = self.lib.fn("strerror")(rc).decode()
            if self.lib.is_device and rc == -3:
                msg += ": " + self.lib.fn("last_cuda_error")(self._h).decode()
            raise LmsfError(rc, msg)

    # -- seam 1
    def extract_features(self, xyzi):
        a = _xyzi(xyzi)
        n = a.shape[0]
        lab = np.zeros(n, np.uint8)
        e = np.empty((max(n, 1), 4), np.float32)
        s = np.empty((max(n, 1), 4), np.float32)
        ne, ns = C.c_int(0), C.c_int(0)
        self._chk(self.lib.fn("extract_features")(self._h, _fp(a), n, lab.ctypes.data_as(_u8p), _fp(e), C.byref(ne),
                                                  _fp(s), C.byref(ns)))
        return lab, e[: ne.value].copy(), s[: ns.value].copy()

    # -- seam 2
    def voxel_downsample(self, xyzi, leaf: float, want_membership: bool = True):
        a = _xyzi(xyzi)
        n = a.shape[0]
        out = np.empty((max(n, 1), 4), np.float32)
        mem = np.empty(max(n, 1), np.int32) if want_membership else None
        no = C.c_int(0)
        self._chk(self.lib.fn("voxel_downsample")(self._h, _fp(a), n, float(leaf), _fp(out), C.byref(no),
                                                  mem.ctypes.data_as(_i32p) if mem is not None else None))
        return out[: no.value].copy(), (mem[:n].copy() if mem is not None else None)

    def common_process(self, xyzi, remove_nan: bool = True, leaf: float = 0.0, dist_near: float = 0.0,
                       dist_far: float = 0.0):
        """PointCloudCommonProcess::Process: removeNaN -> VoxelGrid(leaf) -> DistanceFilter(near, far)."""
        a = _xyzi(xyzi)
        n = a.shape[0]
        out = np.empty((max(n, 1), 4), np.float32)
        no = C.c_int(0)
        self._chk(self.lib.fn("common_process")(self._h, _fp(a), n, int(bool(remove_nan)), float(leaf), float(dist_near),
                                                float(dist_far), _fp(out), C.byref(no)))
        return out[: no.value].copy()

    # -- seam 3
    def map_set(self, kind: int, xyzi):
        a = _xyzi(xyzi)
        self._chk(self.lib.fn("map_set")(self._h, kind, _fp(a), a.shape[0]))

    def knn5(self, kind: int, q_xyz):
        q = np.ascontiguousarray(q_xyz, dtype=np.float32).reshape(-1, 3)
        nq = q.shape[0]
        idx = np.empty((max(nq, 1), 5), np.int32)
        d2 = np.empty((max(nq, 1), 5), np.float32)
        self._chk(self.lib.fn("knn5")(self._h, kind, _fp(q), nq, idx.ctypes.data_as(_i32p), _fp(d2)))
        return idx[:nq], d2[:nq]

    def match(self, kind: int, q_xyz):
        q = np.ascontiguousarray(q_xyz, dtype=np.float32).reshape(-1, 3)
        nq = q.shape[0]
        ok = np.zeros(max(nq, 1), np.uint8)
        out = np.zeros((max(nq, 1), 10), np.float64)
        self._chk(self.lib.fn("match")(self._h, kind, _fp(q), nq, ok.ctypes.data_as(_u8p),
                                       out.ctypes.data_as(_f64p)))
        return ok[:nq].astype(bool), out[:nq]

    def register(self, edge, surf, pose=IDENTITY_POSE, solver: int | None = None):
        e, s = _xyzi(edge), _xyzi(surf)
        p = np.array(pose, dtype=np.float64)
        st = RegStats()
        sv = self.params.solver if solver is None else solver
        self._chk(self.lib.fn("register")(self._h, _fp(e), e.shape[0], _fp(s), s.shape[0], sv,
                                          p.ctypes.data_as(_f64p), C.byref(st)))
        return p, st.as_dict()

    def set_lm_outer(self, count: int):
        self._chk(self.lib.fn("set_lm_outer")(self._h, count))

    # -- tracker
    def tracker_step(self, xyzi, stamp: float, delta=IDENTITY_POSE):
        a = _xyzi(xyzi)
        d = np.array(delta, dtype=np.float64)
        p = np.zeros(7, np.float64)
        st = TrackStats()
        self._chk(self.lib.fn("tracker_step")(self._h, _fp(a), a.shape[0], float(stamp), d.ctypes.data_as(_f64p),
                                              p.ctypes.data_as(_f64p), C.byref(st)))
        return p, d, st.as_dict()

    def tracker_step_dev(self, d_ptr: int, n: int, stamp: float, delta=IDENTITY_POSE):
        d = np.array(delta, dtype=np.float64)
        p = np.zeros(7, np.float64)
        st = TrackStats()
        self._chk(self.lib.fn("tracker_step_dev")(self._h, C.c_void_p(d_ptr), n, float(stamp),
                                                  d.ctypes.data_as(_f64p), p.ctypes.data_as(_f64p), C.byref(st)))
        return p, d, st.as_dict()

    def tracker_prefetch(self, xyzi) -> int:
        """Enqueue upload + feature extraction of a COMING sweep (copied now); returns the ticket that
        tracker_step_ticket / tracker_submit_ticket consume."""
        a = _xyzi(xyzi)
        t = C.c_int64(0)
        self._chk(self.lib.fn("tracker_prefetch")(self._h, _fp(a), a.shape[0], C.byref(t)))
        return t.value

    def tracker_prefetch_dev(self, d_ptr: int, n: int) -> int:
        t = C.c_int64(0)
        self._chk(self.lib.fn("tracker_prefetch_dev")(self._h, C.c_void_p(d_ptr), n, C.byref(t)))
        return t.value

    def tracker_prefetch_cancel(self, ticket: int):
        self._chk(self.lib.fn("tracker_prefetch_cancel")(self._h, int(ticket)))

    def tracker_step_ticket(self, ticket: int, stamp: float, delta=IDENTITY_POSE):
        d = np.array(delta, dtype=np.float64)
        p = np.zeros(7, np.float64)
        st = TrackStats()
        self._chk(self.lib.fn("tracker_step_ticket")(self._h, int(ticket), float(stamp), d.ctypes.data_as(_f64p),
                                                     p.ctypes.data_as(_f64p), C.byref(st)))
        return p, d, st.as_dict()

    def tracker_submit_ticket(self, ticket: int, stamp: float, delta=IDENTITY_POSE):
        d = np.array(delta, dtype=np.float64)
        self._chk(self.lib.fn("tracker_submit_ticket")(self._h, int(ticket), float(stamp), d.ctypes.data_as(_f64p)))

    def tracker_submit(self, xyzi, stamp: float, delta=IDENTITY_POSE):
        """First half of tracker_step: everything enqueued, nothing waited for."""
        a = _xyzi(xyzi)
        d = np.array(delta, dtype=np.float64)
        self._chk(self.lib.fn("tracker_submit")(self._h, _fp(a), a.shape[0], float(stamp), d.ctypes.data_as(_f64p)))

    def tracker_submit_dev(self, d_ptr: int, n: int, stamp: float, delta=IDENTITY_POSE):
        d = np.array(delta, dtype=np.float64)
        self._chk(self.lib.fn("tracker_submit_dev")(self._h, C.c_void_p(d_ptr), n, float(stamp),
                                                    d.ctypes.data_as(_f64p)))

    def tracker_wait(self):
        """Second half of tracker_step: (pose, motion increment, statistics) of the submitted sweep."""
        d = np.zeros(7, np.float64)
        p = np.zeros(7, np.float64)
        st = TrackStats()
        self._chk(self.lib.fn("tracker_wait")(self._h, d.ctypes.data_as(_f64p), p.ctypes.data_as(_f64p), C.byref(st)))
        return p, d, st.as_dict()

    def tracker_step_features(self, edge, surf, stamp: float, delta=IDENTITY_POSE):
        e, s = _xyzi(edge), _xyzi(surf)
        d = np.array(delta, dtype=np.float64)
        p = np.zeros(7, np.float64)
        st = TrackStats()
        self._chk(self.lib.fn("tracker_step_features")(self._h, _fp(e), e.shape[0], _fp(s), s.shape[0], float(stamp),
                                                       d.ctypes.data_as(_f64p), p.ctypes.data_as(_f64p),
                                                       C.byref(st)))
        return p, d, st.as_dict()

    def tracker_reset(self):
        self._chk(self.lib.fn("tracker_reset")(self._h))

    def tracker_register_aux(self, xyzi, pose):
        a = _xyzi(xyzi)
        p = np.array(pose, dtype=np.float64)
        st = RegStats()
        self._chk(self.lib.fn("tracker_register_aux")(self._h, _fp(a), a.shape[0], p.ctypes.data_as(_f64p),
                                                      C.byref(st)))
        return p, st.as_dict()

    def rotary_preprocess(self, xyzi, scan_period: float = 0.1):
        """removeNaN + RotaryLidarPreProcess::Process: the sweep with intensity := relative time."""
        a = _xyzi(xyzi)
        out = np.empty((max(a.shape[0], 1), 4), np.float32)
        n = C.c_int(0)
        self._chk(self.lib.fn("rotary_preprocess")(self._h, _fp(a), a.shape[0], float(scan_period), _fp(out), C.byref(n)))
        return out[: n.value].copy()

    def extract_features_to_dev(self, xyzi, d_out: int, cap: int):
        """Extraction with the features left on the device: edges then surfs into the buffer at d_out -> (n_edge, n_surf)."""
        a = _xyzi(xyzi)
        ne, ns = C.c_int(0), C.c_int(0)
        self._chk(self.lib.fn("extract_features_to_dev")(self._h, _fp(a), a.shape[0], C.c_void_p(d_out), int(cap),
                                                         C.byref(ne), C.byref(ns)))
        return ne.value, ns.value

    def tracker_register_aux_features_dev(self, d_feat: int, n_edge: int, n_surf: int, pose):
        p = np.array(pose, dtype=np.float64)
        st = RegStats()
        self._chk(self.lib.fn("tracker_register_aux_features_dev")(self._h, C.c_void_p(d_feat), int(n_edge), int(n_surf),
                                                                   p.ctypes.data_as(_f64p), C.byref(st)))
        return p, st.as_dict()

    def get_map(self, kind: int):
        n = C.c_int(0)
        self._chk(self.lib.fn("get_map")(self._h, kind, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 4), np.float32)
        self._chk(self.lib.fn("get_map")(self._h, kind, _fp(out), out.shape[0], C.byref(n)))
        return out[: n.value].copy()

    # -- device-only helpers
    def launch_count(self) -> int:
        return int(self.lib.fn("launch_count")(self._h))

    def stream(self) -> int:
        return int(self.lib.fn("stream")(self._h) or 0)

    def dev_upload_new(self, arr) -> int:
        a = np.ascontiguousarray(arr)
        ptr = C.c_void_p()
        self._chk(self.lib.fn("dev_alloc")(self._h, a.nbytes, C.byref(ptr)))
        self._chk(self.lib.fn("dev_upload")(self._h, ptr, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return int(ptr.value)

    def dev_free(self, ptr: int):
        self._chk(self.lib.fn("dev_free")(self._h, C.c_void_p(ptr)))

    def align_score(self, kind: int, xyzi, relpose4x4, inlier_thresh: float, inlier_ratio_thresh: float):
        """AlignmentScore against the map index `kind`: (score, overlap ratio, inlier count)."""
        a = _xyzi(xyzi)
        T = np.ascontiguousarray(relpose4x4, dtype=np.float32).reshape(16)
        sc, ov = C.c_double(0), C.c_double(0)
        ni = C.c_int32(0)
        self._chk(self.lib.fn("align_score")(self._h, kind, _fp(a), a.shape[0], _fp(T), float(inlier_thresh),
                                             float(inlier_ratio_thresh), C.byref(sc), C.byref(ov), C.byref(ni)))
        return sc.value, ov.value, ni.value

    # -- loop-closure descriptors (ScanContext), device library only
    def sc_make(self, xyzi):
        """MakeScanContext + MakeRingkeyFromScanContext: (20, 60) descriptor and (20,) ring key."""
        a = _xyzi(xyzi)
        desc = np.empty((SC_RINGS, SC_SECTORS), np.float32)
        key = np.empty(SC_RINGS, np.float32)
        self._chk(self.lib.fn("sc_make")(self._h, _fp(a), a.shape[0], _fp(desc), _fp(key)))
        return desc, key

    def sc_distance(self, desc_a, desc_b):
        """DistanceBtnScanContext for explicit pairs: (n,) distances and column shifts."""
        a = np.ascontiguousarray(desc_a, dtype=np.float32).reshape(-1, SC_CELLS)
        b = np.ascontiguousarray(desc_b, dtype=np.float32).reshape(-1, SC_CELLS)
        if a.shape != b.shape:
            raise ValueError("descriptor batches differ in shape")
        n = a.shape[0]
        dist = np.empty(max(n, 1), np.float64)
        shift = np.empty(max(n, 1), np.int32)
        self._chk(self.lib.fn("sc_distance")(self._h, _fp(a), _fp(b), n, dist.ctypes.data_as(_f64p),
                                             shift.ctypes.data_as(_i32p)))
        return dist[:n], shift[:n]

    def scdb_reserve(self, capacity: int):
        self._chk(self.lib.fn("scdb_reserve")(self._h, int(capacity)))

    def scdb_clear(self):
        self._chk(self.lib.fn("scdb_clear")(self._h))

    def scdb_size(self) -> int:
        n = C.c_int(0)
        self._chk(self.lib.fn("scdb_size")(self._h, C.byref(n)))
        return n.value

    def scdb_add(self, descs, keys):
        d = np.ascontiguousarray(descs, dtype=np.float32).reshape(-1, SC_CELLS)
        k = np.ascontiguousarray(keys, dtype=np.float32).reshape(-1, SC_RINGS)
        if d.shape[0] != k.shape[0]:
            raise ValueError("descriptor / key counts differ")
        self._chk(self.lib.fn("scdb_add")(self._h, _fp(d), _fp(k), d.shape[0]))

    def scdb_add_cloud(self, xyzi) -> int:
        a = _xyzi(xyzi)
        i = C.c_int(-1)
        self._chk(self.lib.fn("scdb_add_cloud")(self._h, _fp(a), a.shape[0], C.byref(i)))
        return i.value

    def scdb_get(self, idx: int):
        desc = np.empty((SC_RINGS, SC_SECTORS), np.float32)
        key = np.empty(SC_RINGS, np.float32)
        self._chk(self.lib.fn("scdb_get")(self._h, int(idx), _fp(desc), _fp(key)))
        return desc, key

    def scdb_knn(self, q_keys, limit: int):
        q = np.ascontiguousarray(q_keys, dtype=np.float32).reshape(-1, SC_RINGS)
        nq = q.shape[0]
        idx = np.empty((max(nq, 1), SC_CANDIDATES), np.int32)
        d = np.empty((max(nq, 1), SC_CANDIDATES), np.float32)
        self._chk(self.lib.fn("scdb_knn")(self._h, _fp(q), nq, int(limit), idx.ctypes.data_as(_i32p), _fp(d)))
        return idx[:nq], d[:nq]

    def scdb_search(self, q_keys, q_descs, limit: int, thresh: float = SC_DIST_THRES):
        """descFindSimilar for a batch: loop ids (-1 = not a loop), SC distances, column shifts."""
        k = np.ascontiguousarray(q_keys, dtype=np.float32).reshape(-1, SC_RINGS)
        d = np.ascontiguousarray(q_descs, dtype=np.float32).reshape(-1, SC_CELLS)
        nq = k.shape[0]
        lid = np.empty(max(nq, 1), np.int32)
        dist = np.empty(max(nq, 1), np.float64)
        sh = np.empty(max(nq, 1), np.int32)
        self._chk(self.lib.fn("scdb_search")(self._h, _fp(k), _fp(d), nq, int(limit), float(thresh),
                                             lid.ctypes.data_as(_i32p), dist.ctypes.data_as(_f64p),
                                             sh.ctypes.data_as(_i32p)))
        return lid[:nq], dist[:nq], sh[:nq]

    def scdb_search_shard_dev(self, d_q_keys: int, d_q_descs: int, nq: int, limit_local: int, id_base: int,
                              d_cand: int):
        self._chk(self.lib.fn("scdb_search_shard_dev")(self._h, C.c_void_p(d_q_keys), C.c_void_p(d_q_descs), nq,
                                                       int(limit_local), int(id_base), C.c_void_p(d_cand)))

    def scdb_keys_shard_dev(self, d_q_keys: int, nq: int, limit_local: int, id_base: int, d_cand: int):
        self._chk(self.lib.fn("scdb_keys_shard_dev")(self._h, C.c_void_p(d_q_keys), nq, int(limit_local), int(id_base),
                                                     C.c_void_p(d_cand)))

    def scdb_score_owned_dev(self, d_cand_all: int, n_ranks: int, nq: int, d_q_descs: int, id_base: int, n_local: int,
                             d_scored: int):
        self._chk(self.lib.fn("scdb_score_owned_dev")(self._h, C.c_void_p(d_cand_all), n_ranks, nq, C.c_void_p(d_q_descs),
                                                      int(id_base), int(n_local), C.c_void_p(d_scored)))

    def scdb_pick_dev(self, d_cand_all: int, n_ranks: int, nq: int, thresh: float, d_id: int, d_dist: int,
                      d_shift: int):
        self._chk(self.lib.fn("scdb_pick_dev")(self._h, C.c_void_p(d_cand_all), n_ranks, nq, float(thresh),
                                               C.c_void_p(d_id), C.c_void_p(d_dist), C.c_void_p(d_shift)))

    def profile_enable(self, on: bool = True):
        self._chk(self.lib.fn("profile_enable")(self._h, int(on)))

    def profile_read(self, reset: bool = True):
        ms = (C.c_double * N_STAGES)()
        ln = (C.c_int64 * N_STAGES)()
        by = C.c_double(0)
        self._chk(self.lib.fn("profile_read")(self._h, ms, ln, C.byref(by), int(reset)))
        return ({STAGE_NAMES[i]: ms[i] for i in range(N_STAGES)}, {STAGE_NAMES[i]: ln[i] for i in range(N_STAGES)},
                by.value)
