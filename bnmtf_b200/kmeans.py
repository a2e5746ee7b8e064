"""K-means with missing values, used for init_FG='kmeans' (host-side port of the behaviour of
code/models/kmeans/kmeans.py:6-205; one-off initialisation, outside the update sweep -- SURVEY.md section 8f #2).

Semantics kept: centroids drawn with python's `random.uniform` between the per-coordinate min and max of the
observed values (same call order, so a seeded run picks the same centroids); distance = mean squared difference
over the coordinates observed in both the point and the centroid (None / infinite when there is no overlap, ties go
to the lowest cluster index); centroid = per-coordinate mean of the observed values of its points (coordinate masked
out when none is observed); an empty cluster takes the point currently furthest from its centroid ('singleton');
stop when no assignment changes or after 200 iterations; result = one-hot assignment matrix.

One reference quirk is kept on purpose, because it changes the clustering in about a third of seeded runs on the toy
data: the 'singleton' rule binds the empty cluster's centroid to the ROW OF X ITSELF (kmeans.py:149,
`self.centroids[c] = self.X[index_furthest_away]`, a numpy view of the model's private copy of X), so every later
centroid update of that cluster (kmeans.py:170) also overwrites that data point.  Here the centroids are a list of row
arrays as well, and the singleton rule stores the view, so the same write-through happens.

The assignment step -- the no_points x K x no_coordinates part, everything else is O(K x no_coordinates) bookkeeping -- runs
on the GPU when a device is given (csrc/kmeans.cu through bnmtf_kmeans_distances_f64: the same bits as the numpy
expression in _all_distances, so the clustering is the same); the model classes always pass their device.  Without a
device the class is the plain host port (tests/test_kmeans.py pins it against the reference's goldens on CPU).
"""
import random

import numpy as np

MAX_ITERATIONS = 200


class KMeans(object):
    def __init__(self, X, M, K, resolve_empty='singleton', device=None):
        self.device = device
        self._dev = None                # (X, M, centroid, mask, distance) device tensors, made on first use
        self._aliased = set()           # rows of X a centroid is a view of (they change under update_cluster)
        self.X = np.array(X, dtype=float)
        self.M = np.array(M, dtype=float)
        self.K = K
        self.resolve_empty = resolve_empty
        assert len(self.X.shape) == 2, "Input matrix X is not a two-dimensional array, but instead %s-dimensional." % len(self.X.shape)
        assert self.X.shape == self.M.shape, "Input matrix X is not of the same size as the indicator matrix M: %s and %s respectively." % (self.X.shape, self.M.shape)
        assert self.K > 0, "K should be greater than 0."
        (self.no_points, self.no_coordinates) = self.X.shape
        self.no_unique_points = len(set(tuple(row) for row in self.X.tolist()))
        for i, c in enumerate(self.M.sum(axis=1)):
            assert c != 0, "Fully unobserved row in X, row %s." % i
        keep = self.M.sum(axis=0) != 0
        if not keep.all():
            self.X, self.M = self.X[:, keep], self.M[:, keep]
            self.no_coordinates = self.X.shape[1]
        self.distances = np.zeros(self.no_points)

    def initialise(self, seed=None):
        if seed is not None:
            random.seed(seed)
        obs = self.M != 0
        self.mins = [self.X[obs[:, j], j].min() for j in range(self.no_coordinates)]
        self.maxs = [self.X[obs[:, j], j].max() for j in range(self.no_coordinates)]
        self.centroids = [np.array(self.random_cluster_centroid(), dtype=float) for _ in range(self.K)]
        self.cluster_assignments = np.full(self.no_points, -1, dtype=int)
        self.mask_centroids = np.ones((self.K, self.no_coordinates))

    def random_cluster_centroid(self):
        return [random.uniform(self.mins[c], self.maxs[c]) for c in range(self.no_coordinates)]

    def cluster(self):
        iteration = 1
        change = True
        while change:
            iteration += 1
            change = self.assignment()
            self.update()
            if iteration >= MAX_ITERATIONS:
                break
        self.create_matrix()

    def _all_distances_device(self):
        import torch
        from . import _lib
        from .engine import _ptr, _stream
        n, d, K = self.no_points, self.no_coordinates, self.K
        if self._dev is None:
            dev = torch.device(self.device)
            f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
            self._dev = {"X": f(self.X), "M": f(self.M), "C": torch.empty((K, d), dtype=torch.float64, device=dev),
                         "MC": torch.empty((K, d), dtype=torch.float64, device=dev),
                         "D": torch.empty((n, K), dtype=torch.float64, device=dev)}
        t = self._dev
        for i in sorted(self._aliased):                 # data points overwritten through a centroid view (see top)
            t["X"][i].copy_(torch.from_numpy(np.ascontiguousarray(self.X[i])))
        t["C"].copy_(torch.from_numpy(np.stack(self.centroids)))
        t["MC"].copy_(torch.from_numpy(np.ascontiguousarray(self.mask_centroids, dtype=np.float64)))
        with torch.cuda.device(t["X"].device):
            _lib.call("bnmtf_kmeans_distances_f64", _ptr(t["X"]), _ptr(t["M"]), n, d, _ptr(t["C"]), _ptr(t["MC"]), K, _ptr(t["D"]),
                      _stream())
        return t["D"].cpu().numpy()

    def _all_distances(self):
        """no_points x K matrix of masked mean squared differences (inf where nothing overlaps)."""
        if self.device is not None:
            return self._all_distances_device()
        both = self.M[:, None, :] * self.mask_centroids[None, :, :]
        overlap = both.sum(axis=2)
        sq = (both * (self.X[:, None, :] - np.stack(self.centroids)[None, :, :]) ** 2).sum(axis=2)
        with np.errstate(all='ignore'):
            return np.where(overlap > 0, sq / overlap, np.inf)

    def assignment(self):
        dist = self._all_distances()
        new = dist.argmin(axis=1)                       # first minimum = lowest index on ties
        best = dist[np.arange(self.no_points), new]
        self.distances = np.where(np.isfinite(best), best, 0.0)
        change = bool((new != self.cluster_assignments).any())
        self.cluster_assignments = new
        self.data_point_assignments = [list(np.nonzero(new == c)[0]) for c in range(self.K)]
        return change

    def update(self):
        for c in range(self.K):
            self.update_cluster(c)

    def update_cluster(self, c):
        members = self.data_point_assignments[c]
        if len(members) == 0:
            if self.no_unique_points >= self.K:
                if self.resolve_empty == 'singleton':
                    far = int(self.distances.argmax())
                    old = int(self.cluster_assignments[far])
                    self.centroids[c] = self.X[far]          # a VIEW: later updates of c write through into X (see top)
                    self._aliased.add(far)
                    self.mask_centroids[c] = self.M[far]
                    self.distances[far] = 0.0
                    self.cluster_assignments[far] = c
                    self.data_point_assignments[c] = [far]
                    self.data_point_assignments[old].remove(far)
                    self.update_cluster(old)
                else:
                    self.centroids[c] = np.array(self.random_cluster_centroid(), dtype=float)
                    self.mask_centroids[c] = np.ones(self.no_coordinates)
            return
        Xc, Mc = self.X[members], self.M[members]
        counts = Mc.sum(axis=0)
        with np.errstate(all='ignore'):
            means = np.where(counts > 0, (Mc * Xc).sum(axis=0) / counts, 0.0)
        self.centroids[c][:] = means                         # in place, like the reference's per-coordinate writes
        self.mask_centroids[c] = (counts > 0).astype(float)

    def create_matrix(self):
        self.clustering_results = np.zeros((self.no_points, self.K))
        self.clustering_results[np.arange(self.no_points), self.cluster_assignments] = 1
