"""Time the row-solver kernels against each other at the row counts of 1/2/4/8-way shards (one GPU; the statistics are
real but the matrix is narrow, the solver does not care): ms per launch, Gibbs and VB, mean of 10 after 3 warm-ups."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnmtf_b200 import bnmf

K = 20
pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
rng = np.random.RandomState(0)
print("rows    mode   " + "  ".join("%7s" % s for s in ("auto", "warp", "lane", "sub4", "sub8", "sub16", "sub32")))
for rows in (4096, 8192, 16384, 32768, 65536):
    cols = 1024
    R = rng.exponential(1.0, (rows, K)) @ rng.exponential(1.0, (cols, K)).T + rng.normal(size=(rows, cols))
    M = (rng.rand(rows, cols) >= 0.2).astype(float)
    for mode, cls in (("gibbs", bnmf.bnmf_gibbs_optimised), ("vb", bnmf.bnmf_vb_optimised)):
        m = cls(R, M, K, pri, seed=1)
        m.initialise("exp")
        m.run(1)
        eng = m._push()
        eng.stats(0)
        res = []
        for solver in ("", "warp", "lane", "sub4", "sub8", "sub16", "sub32"):
            if solver:
                os.environ["BNMTF_SOLVE"] = solver
            else:
                os.environ.pop("BNMTF_SOLVE", None)
            for _ in range(3):
                eng.solve(0, gather=False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.solve(0, gather=False)
            e1.record()
            e1.synchronize()
            res.append(e0.elapsed_time(e1) / 10)
        os.environ.pop("BNMTF_SOLVE", None)
        print("%-7d %-6s " % (rows, mode) + "  ".join("%7.3f" % t for t in res), flush=True)
