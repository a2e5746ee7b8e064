"""Device assignment step of K-means with missing values (csrc/kmeans.cu, bnmtf_kmeans_distances_f64) against the host
statement of code/models/kmeans/kmeans.py:105-133 (bnmtf_b200/kmeans.py::_all_distances, itself pinned against the
reference's goldens on CPU): the SAME BITS -- near-ties between centroids decide the clustering -- and, through the device
path, the reference's clusterings of tests/golden/kmeans.json."""
import functools
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_kmeans as gen  # noqa: E402

from bnmtf_b200.kmeans import KMeans  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "kmeans.json")))


# numpy's pairwise summation: < 8 terms, one block (<= 128), two blocks, uneven halves, the toy / GDSC widths, many blocks
@pytest.mark.parametrize("n,d,K", [(9, 5, 2), (40, 8, 3), (64, 80, 5), (30, 100, 10), (33, 129, 4), (50, 138, 10),
                                   (17, 257, 7), (25, 622, 10), (12, 1100, 33), (6, 4099, 3)])
def test_device_distances_have_the_host_bits(n, d, K):
    rng = np.random.RandomState(n + d + K)
    X = rng.normal(size=(n, d)) * np.exp(rng.normal(size=(n, 1)))
    M = (rng.rand(n, d) < 0.7).astype(float)
    M[np.arange(n), rng.randint(0, d, n)] = 1.0                       # no fully unobserved row
    M[:, M.sum(axis=0) == 0] = 1.0
    host, dev = KMeans(X, M, K), KMeans(X, M, K, device="cuda:0")
    for km in (host, dev):
        km.initialise(seed=3)
    dev.mask_centroids[0, : d // 2] = host.mask_centroids[0, : d // 2] = 0.0          # a centroid with unobserved coordinates
    if n > 10:
        # a point and a centroid without any common coordinate: infinite distance on both sides
        M2 = M.copy()
        M2[1] = 0.0
        M2[1, 0] = 1.0
        host, dev = KMeans(X, M2, K), KMeans(X, M2, K, device="cuda:0")
        for km in (host, dev):
            km.initialise(seed=3)
            km.mask_centroids[0, 0] = 0.0
    a, b = host._all_distances(), dev._all_distances()
    assert a.shape == b.shape == (n, K)
    assert np.array_equal(a, b)                                       # bit for bit, inf included
    if n > 10:
        assert np.isinf(a[1, 0])


@pytest.mark.parametrize("want", GOLDEN[::3], ids=lambda c: "K%d-%s-seed%d" % (c["K"], c["side"], c["seed"]))
def test_device_kmeans_matches_the_reference(want):
    toy = np.load(os.path.join(HERE, "golden", "toy_bnmtf_vb.npz"))
    got = gen.run_case(functools.partial(KMeans, device="cuda:0"), toy["R"], toy["M"], want["K"], want["side"], want["seed"])
    assert got == want


def test_model_classes_use_the_device_step(monkeypatch):
    import bnmtf_b200
    from bnmtf_b200 import _lib
    toy = np.load(os.path.join(HERE, "golden", "toy_bnmtf_vb.npz"))
    calls = []
    real = _lib.call
    monkeypatch.setattr(_lib, "call", lambda name, *a: (calls.append(name), real(name, *a))[1])
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaF": 1.0, "lambdaS": 1.0, "lambdaG": 1.0}
    m = bnmtf_b200.bnmtf_vb_optimised(toy["R"], toy["M"], 5, 5, pri)
    np.random.seed(0)
    import random
    random.seed(0)
    m.initialise(init_S="random", init_FG="kmeans")
    assert "bnmtf_kmeans_distances_f64" in calls
