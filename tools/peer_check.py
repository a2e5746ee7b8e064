"""2+ GPU check of the fused (in-kernel, peer-store) factor exchange against the NCCL all-gather path: same sharded
Gibbs / VB runs with BNMTF_PEER=1 and =0 must agree exactly (same kernels, same values; only the transport differs).
    torchrun --nproc-per-node 2 tools/peer_check.py"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnmtf_b200 import parallel, bnmf

rank, world = parallel.init_process_group("nccl")
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
rng = np.random.RandomState(3)
pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
out = {}
for shape, K in (((300, 170), 7), ((40000, 96), 20), ((130, 36000), 12)):
    I, J = shape
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J))
    M = (rng.rand(I, J) >= 0.25).astype(float)
    for peer in ("1", "0"):
        os.environ["BNMTF_PEER"] = peer
        for name, cls in (("gibbs", bnmf.bnmf_gibbs_optimised), ("vb", bnmf.bnmf_vb_optimised)):
            m = cls(R, M, K, pri, seed=3, distributed=True)
            m.initialise("exp")
            t0 = time.time()
            m.run(4)
            fused = m._engine().U.peer is not None
            out[(shape, name, peer)] = (m.U.copy() if name == "gibbs" else m.expU.copy(), list(m.all_performances["MSE"]), fused, time.time() - t0)
    for name in ("gibbs", "vb"):
        a, b = out[(shape, name, "1")], out[(shape, name, "0")]
        if rank == 0:
            print(shape, name, "fused path active:", a[2], "| nccl run fused:", b[2], "| max |dU|", float(np.abs(a[0] - b[0]).max()),
                  "| MSE equal:", a[1] == b[1], "| t %.3f vs %.3f s" % (a[3], b[3]), flush=True)
        assert a[2] and not b[2]
        assert np.array_equal(a[0], b[0]) and a[1] == b[1]
torch.distributed.barrier()
if rank == 0:
    print("PEER CHECK OK")
torch.distributed.destroy_process_group()
