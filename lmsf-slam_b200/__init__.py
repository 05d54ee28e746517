"""lmsf-slam_b200 — B200-native LiDAR scan-to-map registration hot path.

Host-side Python mirror of the C ABI in include/lmsf_b200.h (used by the tests
and bench.py; the product boundary itself is the C ABI plus the C++ adapters in
include/lmsf_b200_adapters.hpp).  The directory name contains a hyphen, so it is
imported under the module name ``lmsf_slam_b200`` via ``__graft_entry__.load_package()``.

The CUDA library is mandatory: ``library()`` raises when csrc/liblmsf_b200.so has
not been built, and every ABI call fails with LMSF_ERR_NO_DEVICE when no CUDA
device is usable.  There is no CPU fallback.
"""
from __future__ import annotations

import os

from . import capi, rig, shard, synth  # noqa: F401
from .capi import (IDENTITY_POSE, KIND_EDGE, KIND_SURF, SOLVER_GN, SOLVER_HUBER_LM, Context, Library,  # noqa: F401
                   LmsfError, Params, RegStats, TrackStats)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LMSF_B200_LIB") or os.path.join(PKG_DIR, "csrc", "liblmsf_b200.so")  # env: tuning builds only

_lib = None


def library() -> Library:
    """The CUDA implementation of the ABI (csrc/liblmsf_b200.so)."""
    global _lib
    if _lib is None:
        _lib = Library(LIB_PATH, "lmsf_")
    return _lib


def context(device: int = 0, **params) -> Context:
    return library().context(device, **params)
