"""CPU: the simple-knn oracle (oracle/knn_oracle.py) against the golden outputs of the unmodified reference extension and
against an independent exact search."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_knn_golden import knn_cases  # noqa: E402
from oracle import knn_oracle as ko  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "knn_golden.npz")


@pytest.mark.parametrize("name", list(knn_cases().keys()))
def test_oracle_matches_reference_extension(name):
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/knn_golden.npz not generated yet")
    g = np.load(GOLDEN)
    got = ko.mean_dist2(knn_cases()[name])
    np.testing.assert_allclose(got, g[name], rtol=2e-6, atol=1e-12)


def test_oracle_matches_kdtree():
    from scipy.spatial import cKDTree
    pts = knn_cases()["ball5000"]
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=4)
    ref = (d[:, 1:] ** 2).mean(1)
    np.testing.assert_allclose(ko.mean_dist2(pts), ref, rtol=2e-6)


def test_oracle_edge_cases():
    pts = knn_cases()["tiny7"]
    assert np.isinf(ko.mean_dist2(pts[:1])).all() and np.isinf(ko.mean_dist2(pts[:2])).all()       # FLT_MAX + FLT_MAX overflows
    three = ko.mean_dist2(pts[:3])
    assert np.all(three == np.float32(3.4028234663852886e38) / np.float32(3.0))                    # d0 + d1 + FLT_MAX
    dup = ko.mean_dist2(knn_cases()["dups3000"])
    assert np.all(dup[:1500] < ko.mean_dist2(knn_cases()["dups3000"][:1500]))                      # the twin counts, at distance 0
    assert ko.mean_dist2(np.zeros((0, 3), np.float32)).shape == (0,)
