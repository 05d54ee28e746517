"""One-off (too slow for the suite, 2.5 min): the oracle's feature extraction against the reference's own
LOAMFeatureProcessorBase (oracle/_ref/libref_loam.so) on sweeps 0-119 of both bench sequences.
Result (round 1): HDL-64 15 338 957 features, VLP-16 3 422 411 features, 0 mismatching sweeps (bit for bit, order included).
"""
import sys, ctypes as C, numpy as np, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
synth = entry.load_package().synth
lib = C.CDLL(os.path.join(ROOT, 'oracle', '_ref', 'libref_loam.so'))
fp = C.POINTER(C.c_float)
def ref_extract(sw, n_scans):
    sw = np.ascontiguousarray(sw, np.float32); n = len(sw)
    e = np.zeros((n, 4), np.float32); s = np.zeros((n, 4), np.float32)
    ne, ns = C.c_int(), C.c_int()
    assert lib.ref_loam_extract(sw.ctypes.data_as(fp), n, n_scans, C.c_float(2.0), C.c_float(80.0), C.c_float(1.0), 1, n, e.ctypes.data_as(fp), C.byref(ne), s.ctypes.data_as(fp), C.byref(ns)) == 0
    return e[:ne.value], s[:ns.value]
ol = entry.load_oracle()
for name, sensor, ns_, seq in (("hdl64", synth.hdl64(), 64, 0), ("vlp16", synth.vlp16(), 16, 0)):
    o = ol.context(0, n_scans=ns_)
    bad = 0; t = time.time(); tot = 0
    for k in range(0, 120):
        sw = synth.make_sweep(sensor, k)
        re_, rs_ = ref_extract(sw, ns_)
        _, oe, os_ = o.extract_features(sw)
        ok = re_.shape == oe.shape and rs_.shape == os_.shape and np.array_equal(re_.view(np.uint32), oe.view(np.uint32)) and np.array_equal(rs_.view(np.uint32), os_.view(np.uint32))
        bad += (not ok); tot += len(oe) + len(os_)
    print(name, "sweeps 0-119, features", tot, "mismatching sweeps", bad, f"{time.time()-t:.0f}s")
