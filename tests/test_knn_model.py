"""The 5-NN search of csrc/knn.cuh executed on the CPU (csrc/test_knn_model.cu): the same per-thread source the
device runs, unseeded and seeded, against brute force — ids and distance bits identical.  Replaces what pcl::KdTreeFLANN::
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


@pytest.mark.parametrize("n_map,n_q,seed,grow0", [(60000, 4000, 1, 0), (120000, 4000, 2, 0), (3000, 2000, 7, 0), (40, 300, 9, 0),
                                                   (3000, 2000, 7, 8), (60000, 2000, 3, 4), (500, 2000, 5, 12)])
def test_knn_model_matches_brute_force(model_bin, n_map, n_q, seed, grow0):
    """grow0 = MapDev::grow0, the density hint of voxel-filtered maps (the first grown box of a sparse query)."""
    r = subprocess.run([model_bin, str(n_map), str(n_q), str(seed), str(grow0)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout
