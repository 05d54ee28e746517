"""CPU tests of the C-ABI boundary: the CUDA library loads, exports every symbol include/lmsf_b200.h
declares, agrees with the oracle on struct layout and defaults, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import __graft_entry__ as entry

ROOT = entry.ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lmsf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lmsf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(gpu_lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(gpu_lib.dll, s), f"liblmsf_b200.so does not export {s}"


def test_oracle_mirrors_the_abi(gpu_lib, oracle_lib):
    pkg = entry.load_package()
    for name in pkg.capi.Library.COMMON:
        assert oracle_lib.has(name), name
    a, b = gpu_lib.default_params(), oracle_lib.default_params()
    assert C.sizeof(a) == 136 and C.sizeof(b) == 136
    common = type(a).reserved.offset        # the oracle names two of the public struct's reserved words (knn mode, threads)
    assert bytes(a)[:common] == bytes(b)[:common]
    assert not any(bytes(a)[common:]), "lmsf_params.reserved must default to zero"
    assert not hasattr(a, "oracle_knn_mode") and not hasattr(a, "oracle_threads")   # test infrastructure stays out of the ABI
    assert (a.n_scans, a.min_range, a.max_range, a.edge_thresh, a.window) == (16, 2.0, 80.0, 1.0, 10)
    assert (a.gn_max_iters, a.lm_outer_start, a.lm_inner_iters) == (10, 10, 4)
    assert abs(a.huber_delta - 0.1) < 1e-7 and (a.kf_trans, a.kf_rot, a.kf_time) == (0.3, 0.1, 10.0)


def test_product_library_has_no_oracle_dependency():
    pkg = entry.load_package()
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for root, _, files in os.walk(entry.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "liblmsf_oracle" not in src and "lmsf_oracle.h" not in src, os.path.join(root, f)


def test_no_cpu_fallback_without_a_device(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pkg = entry.load_package()
    with pytest.raises(pkg.LmsfError) as ei:
        gpu_lib.context(0)
    assert ei.value.code == -2          # LMSF_ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)


def test_strerror_and_null_handling(gpu_lib):
    f = gpu_lib.fn("strerror")
    assert f(0) == b"ok" and b"CUDA" in f(-3) and b"capacity" in f(-4)
    assert gpu_lib.fn("params_default")(None) == -1
    assert gpu_lib.fn("ctx_create")(0, None, None) == -1
    gpu_lib.fn("ctx_destroy")(None)


def test_cpp_adapter_binary_builds_and_fails_loudly_without_gpu():
    import torch
    exe = os.path.join(ROOT, "tests", "cpp", "adapter_smoke")
    if not os.path.exists(exe):
        entry.build()
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "/dev/null", "0", "0", "16"], capture_output=True, text=True)
    assert r.returncode == 4 and "no CPU path" in r.stderr
    # the same program over the reference's own seam headers (LMSF_WITH_REFERENCE), where /root/reference exists
    if os.path.isdir("/root/reference"):
        ref_exe = exe + "_ref"
        assert os.path.exists(ref_exe)
        r = subprocess.run([ref_exe, "/dev/null", "0", "0", "16"], capture_output=True, text=True)
        assert r.returncode == 4 and "no CPU path" in r.stderr


def test_fmath_atan2f_has_the_c_librarys_bits(tmp_path):
    """csrc/fmath.cuh (the atan2f / atanf the extraction and ScanContext kernels use) compiled for the host against
    the C library's atan2f — std::atan2(float, float) of the reference's translation unit: 9 M arguments, every bit.
    (The device side of the same check is csrc/test_dmath, run by the GPU tests.)"""
    src = os.path.join(entry.CSRC, "test_fmath.cpp")
    exe = str(tmp_path / "test_fmath")
    subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-I", entry.CSRC, "-o", exe, src], check=True)
    r = subprocess.run([exe, "3000000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit mismatches 0" in r.stdout
