"""The cooperative 5-NN search of csrc/knn.cuh executed on the CPU (csrc/test_knn_model.cu): the same template source
the device runs, with one lane per query and as an emulated 32-lane warp (four groups of eight lanes in lockstep),
unseeded and seeded, against brute force — ids and distance bits identical.  Replaces what pcl::KdTreeFLANN::
nearestKSearch(point, 5, ...) answers for registration/FeatureMatch/EdgeFeatureMatch.hpp:38 and surfFeatureMatch.hpp:37."""
import os
import subprocess

import pytest

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lmsf-slam_b200", "csrc")


@pytest.fixture(scope="module")
def model_bin():
    exe = os.path.join(CSRC, "test_knn_model")
    subprocess.run(["make", "-C", CSRC, "test_knn_model"], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("n_map,n_q,seed", [(60000, 600, 1), (3000, 400, 7), (40, 100, 9)])
def test_knn_model_matches_brute_force(model_bin, n_map, n_q, seed):
    r = subprocess.run([model_bin, str(n_map), str(n_q), str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout
