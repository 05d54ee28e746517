import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count() -> int:
    """CUDA devices visible to this process, asked of the driver library (present exactly where a GPU is)."""
    import ctypes

    try:
        drv = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    n = ctypes.c_int(0)
    if drv.cuInit(0) != 0 or drv.cuDeviceGetCount(ctypes.byref(n)) != 0:
        return 0
    return n.value


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a device: skipped (not errored) on a box without one."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu tests run on the B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _ensure_built():
    if not os.path.exists(entry.ORACLE_LIB) or not os.path.exists(pkg.LIB_PATH):
        entry.build()


@pytest.fixture(scope="session")
def oracle_lib():
    _ensure_built()
    return entry.load_oracle()


@pytest.fixture(scope="session")
def gpu_lib():
    _ensure_built()
    return pkg.library()


@pytest.fixture(scope="session")
def synth():
    return pkg.synth


@pytest.fixture(scope="session")
def sweeps(synth):
    """A few cached synthetic sweeps: dict[(sensor, k)] -> (n,4) float32."""
    cache = {}

    def get(name, k, **kw):
        key = (name, k, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = synth.make_sweep(synth.sensor_by_name(name), k, **kw)
        return cache[key]

    return get


def pose_err(a, b):
    """(metres, radians) between two {qx,qy,qz,qw,tx,ty,tz} poses."""
    a, b = np.asarray(a), np.asarray(b)
    dt = float(np.linalg.norm(a[4:] - b[4:]))
    qa, qb = a[:4] / np.linalg.norm(a[:4]), b[:4] / np.linalg.norm(b[:4])
    d = abs(float(np.dot(qa, qb)))
    return dt, 2.0 * float(np.arccos(min(1.0, d)))
