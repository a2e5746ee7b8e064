"""A deterministic CPU stand-in for the NP model protocol (`cls(X, train, **parameters).train(**cfg).predict(test)`),
importable (hence picklable) so that the reference's multiprocessing.Pool can use it too: a rank-`rank` truncated SVD
of the training entries (missing entries filled with the training mean), `shrink` pulling it towards that mean."""
import numpy as np


class SVDModel:
    def __init__(self, X, M, rank, shrink=0.0):
        self.X, self.M, self.rank, self.shrink = np.array(X, dtype=float), np.array(M, dtype=float), rank, shrink

    def train(self, iterations=1):
        mean = (self.X * self.M).sum() / self.M.sum()
        filled = np.where(self.M != 0, self.X, mean)
        for _ in range(iterations):
            u, s, vt = np.linalg.svd(filled, full_matrices=False)
            low = (u[:, :self.rank] * s[:self.rank]) @ vt[:self.rank]
            filled = np.where(self.M != 0, self.X, low)
        self.P = (1.0 - self.shrink) * low + self.shrink * mean

    def predict(self, M_test):
        M_test = np.asarray(M_test, dtype=float)
        n = M_test.sum()
        err = M_test * (self.X - self.P)
        mse = (err ** 2).sum() / n
        mean = (M_test * self.X).sum() / n
        r2 = 1.0 - (err ** 2).sum() / (M_test * (self.X - mean) ** 2).sum()
        return {'MSE': float(mse), 'R^2': float(r2), 'Rp': float(np.corrcoef(self.X[M_test != 0], self.P[M_test != 0])[0, 1])}
