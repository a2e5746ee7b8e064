"""One run(50) of BNMF VB and of BNMTF VB on the GDSC matrix through the single-kernel sweeps (target for ncu)."""
import os, sys, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bnmtf_b200
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
d = np.load(os.path.join(G, "gdsc_bnmf_vb.npz"))
np.random.seed(0), random.seed(0)
m = bnmtf_b200.bnmf_vb_optimised(d["R"], d["M"], 10, {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1})
m.initialise("random")
m.run(50)
m3 = bnmtf_b200.bnmtf_vb_optimised(d["R"], d["M"], 5, 5, {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1})
m3.initialise("random", "random")
m3.run(50)
torch.cuda.synchronize()
print(m.all_performances["MSE"][-1], m3.all_performances["MSE"][-1])
