"""Host K-means port (bnmtf_b200/kmeans.py, used by init_FG='kmeans') against tests/golden/kmeans.json, produced by the
reference's own KMeans (tests/golden/make_golden_kmeans.py): same clusters, same overwritten data points (the
reference's centroid-aliases-X quirk) and the same position in both host random streams afterwards."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_kmeans as gen  # noqa: E402  (only its run_case / CASES; the reference is not imported here)

from bnmtf_b200.kmeans import KMeans  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "kmeans.json")))


def test_fixture_covers_the_singleton_quirk():
    assert len(GOLDEN) == len(gen.CASES)
    assert sum(c["overwritten_points"] > 0 for c in GOLDEN) >= 5


@pytest.mark.parametrize("want", GOLDEN, ids=lambda c: "K%d-%s-seed%d" % (c["K"], c["side"], c["seed"]))
def test_kmeans_matches_the_reference(want):
    toy = np.load(os.path.join(HERE, "golden", "toy_bnmtf_vb.npz"))
    got = gen.run_case(KMeans, toy["R"], toy["M"], want["K"], want["side"], want["seed"])
    assert got == want


def test_constructor_messages_and_unobserved_columns():
    X = np.arange(12.).reshape(4, 3)
    with pytest.raises(AssertionError, match="Fully unobserved row in X, row 1."):
        KMeans(X, np.array([[1, 1, 1], [0, 0, 0], [1, 0, 1], [1, 1, 0]]), 2)
    with pytest.raises(AssertionError, match="K should be greater than 0."):
        KMeans(X, np.ones((4, 3)), 0)
    km = KMeans(X, np.array([[1, 0, 1], [1, 0, 0], [1, 0, 1], [1, 0, 1]]), 2)       # column 1 never observed: dropped
    assert km.no_coordinates == 2
    km.initialise(seed=0)
    km.cluster()
    assert km.clustering_results.shape == (4, 2) and (km.clustering_results.sum(axis=1) == 1).all()
