"""Test-side helpers for the loop-closure descriptor path: ctypes access to the CPU oracle
(oracle/lmsf_oracle_sc.cpp), to the reference's own ring-key KD-tree (oracle/_ref, built from the
reference's nanoflann.hpp + KDTreeVectorOfVectorsAdaptor.h), an independent numpy transcription of
Scancontext.hpp, and seeded descriptor generators.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

import __graft_entry__ as entry

NR, NS, CELLS, K = 20, 60, 1200, 10
CAND_DTYPE = np.dtype([("sc_dist", "<f8"), ("key_dist", "<f4"), ("id", "<i4"), ("shift", "<i4"), ("pad", "<i4")])
assert CAND_DTYPE.itemsize == 24

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)


def _p(a, t):
    return a.ctypes.data_as(t)


class ScOracle:
    def __init__(self):
        self.dll = C.CDLL(entry.ORACLE_LIB)

    def make(self, xyzi):
        a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        desc = np.empty((NR, NS), np.float32)
        key = np.empty(NR, np.float32)
        assert self.dll.lmsf_oracle_sc_make(_p(a, _f32p), C.c_int(a.shape[0]), _p(desc, _f32p), _p(key, _f32p)) == 0
        return desc, key

    def distance(self, a, b):
        a = np.ascontiguousarray(a, np.float32).reshape(-1, CELLS)
        b = np.ascontiguousarray(b, np.float32).reshape(-1, CELLS)
        dist = np.empty(len(a), np.float64)
        shift = np.empty(len(a), np.int32)
        d1, s1 = C.c_double(0), C.c_int(0)
        for i in range(len(a)):
            self.dll.lmsf_oracle_sc_distance(_p(a[i], _f32p), _p(b[i], _f32p), C.byref(d1), C.byref(s1))
            dist[i], shift[i] = d1.value, s1.value
        return dist, shift

    def knn(self, keys, limit, q_keys):
        keys = np.ascontiguousarray(keys, np.float32).reshape(-1, NR)
        q = np.ascontiguousarray(q_keys, np.float32).reshape(-1, NR)
        idx = np.empty((len(q), K), np.int32)
        d = np.empty((len(q), K), np.float32)
        assert self.dll.lmsf_oracle_sc_knn(_p(keys, _f32p), C.c_int(limit), _p(q, _f32p), C.c_int(len(q)),
                                           _p(idx, _i32p), _p(d, _f32p)) == 0
        return idx, d

    def search(self, keys, descs, limit, q_keys, q_descs, thresh=0.2):
        keys = np.ascontiguousarray(keys, np.float32).reshape(-1, NR)
        descs = np.ascontiguousarray(descs, np.float32).reshape(-1, CELLS)
        qk = np.ascontiguousarray(q_keys, np.float32).reshape(-1, NR)
        qd = np.ascontiguousarray(q_descs, np.float32).reshape(-1, CELLS)
        lid = np.empty(len(qk), np.int32)
        dist = np.empty(len(qk), np.float64)
        sh = np.empty(len(qk), np.int32)
        rc = self.dll.lmsf_oracle_sc_search(_p(keys, _f32p), _p(descs, _f32p), C.c_int(limit), _p(qk, _f32p),
                                            _p(qd, _f32p), C.c_int(len(qk)), C.c_double(thresh), _p(lid, _i32p),
                                            _p(dist, _f64p), _p(sh, _i32p))
        assert rc == 0
        return lid, dist, sh


def ref_ringkey_knn10(keys, q_keys):
    """The reference's own ring-key tree (KDTreeVectorOfVectorsAdaptor over nanoflann, leaf 10, metric_L2)."""
    if not os.path.exists(entry.REF_NANOFLANN_LIB):
        return None
    dll = C.CDLL(entry.REF_NANOFLANN_LIB)
    if not hasattr(dll, "ref_nanoflann_ringkey_knn10"):
        return None
    keys = np.ascontiguousarray(keys, np.float32).reshape(-1, NR)
    q = np.ascontiguousarray(q_keys, np.float32).reshape(-1, NR)
    idx = np.empty((len(q), K), np.int32)
    d = np.empty((len(q), K), np.float32)
    dll.ref_nanoflann_ringkey_knn10(_p(keys, _f32p), C.c_int(len(keys)), _p(q, _f32p), C.c_int(len(q)),
                                    _p(idx, _i32p), _p(d, _f32p))
    return idx, d


# ---------------------------------------------------------------- independent numpy transcription
_libm = C.CDLL("libm.so.6")
_libm.atanf.restype = C.c_float
_libm.atanf.argtypes = [C.c_float]


def _atanf(q):
    """std::atan(float) of the reference's translation unit = the C library's atanf, element by element, as float64."""
    return np.array([_libm.atanf(float(v)) for v in np.asarray(q, np.float32)], np.float64)


def py_make_sc(xyzi):
    """Scancontext.hpp:59-104 + :112-126, written independently of the C++ oracle (vectorised numpy)."""
    p = np.asarray(xyzi, np.float32).reshape(-1, 4)
    x, y = p[:, 0], p[:, 1]
    z = (p[:, 2].astype(np.float64) + 2.0).astype(np.float32)
    rng = np.sqrt((x * x + y * y).astype(np.float64)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        k = 180 / np.pi
        ang = np.full(len(p), np.nan, np.float64)
        q1 = (x >= 0) & (y >= 0)
        q2 = (x < 0) & (y >= 0)
        q3 = (x < 0) & (y < 0)
        q4 = (x >= 0) & (y < 0)
        ang[q1] = k * _atanf(y[q1] / x[q1])
        ang[q2] = 180 - k * _atanf(y[q2] / (-x[q2]))
        ang[q3] = 180 + k * _atanf(y[q3] / x[q3])
        ang[q4] = 360 - k * _atanf((-y[q4]) / x[q4])
    ang = ang.astype(np.float32)
    keep = ~(rng.astype(np.float64) > 80.0)

    def cidx(v, n):
        c = np.ceil(v)
        bad = ~((c > -2147483648.0) & (c < 2147483648.0))
        c = np.where(bad, -2147483648.0, c).astype(np.int64)
        return np.maximum(np.minimum(n, c), 1)

    ring = cidx(rng.astype(np.float64) / 80.0 * NR, NR)
    sec = cidx(ang.astype(np.float64) / 360.0 * NS, NS)
    desc = np.full((NR, NS), -1000.0, np.float32)
    ok = keep & ~np.isnan(z)
    np.maximum.at(desc, (ring[ok] - 1, sec[ok] - 1), z[ok])
    desc[desc == -1000.0] = 0.0
    key = np.array([np.float32(sum(float(v) for v in desc[r]) / NS) for r in range(NR)], np.float32)
    return desc, key


def py_sc_distance(a, b):
    """Scancontext.hpp:133-172 with sequential sums (the oracle's convention), independent transcription."""
    a = np.asarray(a, np.float64).reshape(NR, NS)
    b = np.asarray(b, np.float64).reshape(NR, NS)

    def colmean(m):
        out = np.zeros(NS)
        for c in range(NS):
            s = 0.0
            for r in range(NR):
                s += m[r, c]
            out[c] = s / NR
        return out

    v1, v2 = colmean(a), colmean(b)
    best, bestn = 0, 10000000.0
    for s in range(NS):
        sh = np.roll(v2, s)
        q = 0.0
        for c in range(NS):
            d = v1[c] - sh[c]
            q += d * d
        n = np.sqrt(q)
        if n < bestn:
            best, bestn = s, n
    space = sorted([best] + [(best + i + NS) % NS for i in (1, 2, 3)] + [(best - i + NS) % NS for i in (1, 2, 3)])
    arg, mind = 0, 10000000.0
    for s in space:
        bs = np.roll(b, s, axis=1)
        eff, tot = 0, 0.0
        for c in range(NS):
            na = nb = dot = 0.0
            for r in range(NR):
                na += a[r, c] * a[r, c]
                nb += bs[r, c] * bs[r, c]
                dot += a[r, c] * bs[r, c]
            na, nb = np.sqrt(na), np.sqrt(nb)
            if na == 0 or nb == 0:
                continue
            tot = tot + dot / (na * nb)
            eff += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            d = 1.0 - np.float64(tot) / np.float64(eff)
        if d < mind:
            arg, mind = s, d
    return mind, arg


def py_pick(all_cand, thresh=0.2):
    """Selection of descFindSimilar over gathered candidate records (world, nq, 10) -> ids, dists, shifts."""
    w, nq, _ = all_cand.shape
    lid = np.empty(nq, np.int32)
    dist = np.empty(nq, np.float64)
    sh = np.empty(nq, np.int32)
    for q in range(nq):
        c = all_cand[:, q, :].reshape(-1)
        c = c[c["id"] >= 0]
        order = np.lexsort((c["id"], c["key_dist"]))[:K]
        mind, align, nn = 10000000.0, 0, 0
        for j in order:
            if c["sc_dist"][j] < mind:
                mind, align, nn = c["sc_dist"][j], c["shift"][j], c["id"][j]
        dist[q], sh[q], lid[q] = mind, align, (nn if mind < thresh else -1)
    return lid, dist, sh


def py_select_and_score(all_keys, lo, n_local, score):
    """Round 2 of the two-round exchange on one rank, in numpy: the global ring-key top-10 of every query out of the
    gathered unscored records (world, nq, 10); the candidates this rank owns (ids [lo, lo + n_local)) scored with
    score(q, id) -> (dist, shift), the others written as empty slots.  Returns (nq, 10) records."""
    w, nq, _ = all_keys.shape
    out = np.zeros((nq, K), CAND_DTYPE)
    out["id"] = -1
    out["sc_dist"] = 10000000.0
    out["key_dist"] = np.inf
    for q in range(nq):
        c = all_keys[:, q, :].reshape(-1)
        c = c[c["id"] >= 0]
        order = np.lexsort((c["id"], c["key_dist"]))[:K]
        for k, j in enumerate(order):
            gid = int(c["id"][j])
            if lo <= gid < lo + n_local:
                d, s = score(q, gid)
                out[q, k] = (d, c["key_dist"][j], gid, s, 0)
            else:
                out[q, k]["key_dist"] = c["key_dist"][j]
    return out


# ---------------------------------------------------------------- seeded descriptor data
def random_descs(n, seed, zero_cols=0.15):
    """Height-map-like descriptors: smooth-ish positive values, some empty bins and empty sectors."""
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.2, 6.0, size=(n, NR, NS)).astype(np.float32)
    d[rng.random((n, NR, NS)) < 0.3] = 0.0
    cols = rng.random((n, NS)) < zero_cols
    d = np.where(cols[:, None, :], np.float32(0), d).astype(np.float32)
    return d


def keys_of(descs):
    d = np.asarray(descs, np.float32).reshape(-1, NR, NS)
    out = np.empty((len(d), NR), np.float32)
    for i in range(len(d)):
        for r in range(NR):
            s = 0.0
            for v in d[i, r]:
                s += float(v)
            out[i, r] = np.float32(s / NS)
    return out


def keys_of_fast(descs):
    """Row means via cumulative (sequential) fp64 summation — same order as the reference restatement."""
    d = np.asarray(descs, np.float32).reshape(-1, NR, NS).astype(np.float64)
    s = np.cumsum(d, axis=2)[:, :, -1]  # cumsum adds left to right
    return (s / NS).astype(np.float32)
