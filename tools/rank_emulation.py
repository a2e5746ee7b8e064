"""One rank's share of an N-way sharded sweep on ONE GPU (no peers: the exchange is skipped, the other ranks' factor rows
simply keep their start values): the launch sequence, grids and per-kernel work are those of a real rank, so a launch
list taken under ncu shows which kernels do not shrink with N.

    python tools/rank_emulation.py [world] [sweeps]          # prints ms per sweep (CUDA events, graph replay)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/rank_emulation.py 8 2
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bnmtf_b200 import engine, parallel

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
I, J, K = 65536, 32768, 20
dev = torch.device("cuda", 0)
R, bits, RT, bitsT, n_obs = parallel.make_synthetic_shards(I, J, K, dev, 0, world)
ds = engine.Dataset.from_device(R, bits, I, J, RT, bitsT, n_obs=n_obs, world=world, rank=0)
for mode in ("gibbs", "vb"):
    eng = engine.BNMFEngine(ds, K, mode, 1.0, 1.0, seed=1, comm=engine.Comm(1, 0))
    rng = np.random.RandomState(1)
    eng.U.fac[:I] = torch.from_numpy(rng.exponential(1.0, (I, K))).to(dev)
    eng.V.fac[:J] = torch.from_numpy(rng.exponential(1.0, (J, K))).to(dev)
    if eng.vb:
        eng.U.var[:I] = 0.01
        eng.V.var[:J] = 0.01
    eng.scalars[0] = 1.0
    eng.alloc_trace(sweeps + 8)
    for _ in range(3):
        eng.sweep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(sweeps):
        eng.sweep()
    e1.record()
    e1.synchronize()
    print("world %d rank 0 %s: %.3f ms per sweep (graph: %s)" % (world, mode, e0.elapsed_time(e1) / sweeps, eng._graph is not None), flush=True)
    if os.environ.get("PROFILE_KERNELS"):
        print("   per-kernel (isolated / in sweep):", {k: round(v, 3) for k, v in eng.profile_sweep(reps=2).items() if k != "_meta"}, flush=True)
