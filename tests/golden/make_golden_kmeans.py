"""Generate tests/golden/kmeans.json from the REAL reference K-means (build container only).

    python tests/golden/make_golden_kmeans.py

Runs code/models/kmeans/kmeans.py (through oracle/ref_shim.py) on the reference's toy BNMTF matrix and its transpose
-- the two calls `initialise(init_FG='kmeans')` makes (bnmtf_vb_optimised.py:128-139) -- for several K and python
`random` seeds, and stores the final assignments, the number of data points the reference's 'singleton' rule
overwrote (its centroid-aliases-X quirk, kmeans.py:149,170) and the next draw of both host random streams.
"""
import contextlib
import importlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_shim  # noqa: E402

CASES = [(K, side, seed) for K in (4, 5, 6, 7) for side in ("rows", "columns") for seed in (0, 1, 2, 3, 26, 27)]


def run_case(cls, R, M, K, side, seed):
    X, Mx = (R, M) if side == "rows" else (R.T, M.T)
    np.random.seed(seed), random.seed(seed)
    km = cls(X, Mx, K)
    with contextlib.redirect_stdout(io.StringIO()):
        km.initialise()
        km.cluster()
    return {"K": K, "side": side, "seed": seed,
            "assignments": [int(c) for c in km.clustering_results.argmax(axis=1)],
            "overwritten_points": int((np.asarray(km.X) != X).any(axis=1).sum()),
            "next_random": random.random(), "next_numpy": float(np.random.rand())}


def main():
    ref_shim.load()
    cls = importlib.import_module("BNMTF.code.models.kmeans.kmeans").KMeans
    toy = np.load(os.path.join(HERE, "toy_bnmtf_vb.npz"))
    out = [run_case(cls, toy["R"], toy["M"], *c) for c in CASES]
    json.dump(out, open(os.path.join(HERE, "kmeans.json"), "w"), indent=0)
    print("wrote %d cases, %d of them with overwritten data points" % (len(out), sum(c["overwritten_points"] > 0 for c in out)))


if __name__ == "__main__":
    main()
