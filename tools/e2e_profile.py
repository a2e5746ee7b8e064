"""Where does the host time of model.run(1) go at the benchmark shape?  cProfile over 10 calls per model."""
import cProfile, io, os, pstats, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from bnmtf_b200 import bnmf, engine
I, J, K = 65536, 32768, 20
dev = torch.device("cuda", 0)
R, bits, n_obs = bench.make_synthetic(I, J, K, dev)
ds = engine.Dataset.from_device(R, bits, I, J, n_obs=n_obs)
for cls in (bnmf.bnmf_gibbs_optimised, bnmf.bnmf_vb_optimised):
    m = cls.from_dataset(ds, K, bench.PRIORS, seed=1)
    m.initialise("exp")
    m.run(1), m.run(1)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        m.run(1)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(cls.__name__)
    print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000])
