"""The reference arm of bench.py runs without a GPU: check the JSON contract of its line (keys the driver reads) on the
small VLP-16 configuration, and the argument handling of the other modes that need no device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line_contract():
    d = _run("--impl", "reference", "--sensor", "vlp16", "--steps", "2", "--warmup", "3")
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000.0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sweeps" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "VLP-16" in d["metric"] and "workload" in d["config"]


def test_reference_arm_tracks_one_sequence_per_gpu():
    """--impl reference --gpus N tracks N sequences on the host cores (the N-GPU arm's workload), rank 0 only."""
    d = _run("--impl", "reference", "--sensor", "vlp16", "--steps", "2", "--warmup", "3", "--gpus", "2")
    assert d["n_gpus"] == 2 and d["config"]["sequences"] == 2
    assert abs(d["ms_per_step"] * d["value"] - 2000.0) < 1e-6 * 2000.0
    assert "2 tracker(s)" in d["cpu_baseline"]["sample"]


def test_reference_arm_of_the_loop_search_is_declared_unavailable():
    d = _run("--impl", "reference", "--workload", "loopdb")
    assert d["impl"] == "reference" and "unavailable" in d
