"""Device time per sweep of the single-kernel path (csrc/small.cu) on the reference's own matrices, next to the wall time
of the run() call around it and the multi-kernel path (BNMTF_SMALL=0)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bnmtf_b200

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
for name, f, K in (("toy 100x80", "toy_bnmf_vb.npz", 10), ("GDSC 622x138", "gdsc_bnmf_vb.npz", 10)):
    d = np.load(os.path.join(G, f))
    R, M = d["R"], d["M"]
    for cls in (bnmtf_b200.bnmf_gibbs_optimised, bnmtf_b200.bnmf_vb_optimised, bnmtf_b200.nmf_icm):
        for small in ("1", "0"):
            os.environ["BNMTF_SMALL"] = small
            np.random.seed(0)
            m = cls(R, M, K, pri, seed=1)
            m.initialise("random")
            m.run(20)
            torch.cuda.synchronize()
            t0 = time.time()
            m.run(1000)
            torch.cuda.synchronize()
            wall = time.time() - t0
            print("%-13s %-22s small=%s  device %.1f us/sweep   wall %.1f us/sweep   MSE %.6f" % (
                name, cls.__name__, small, m.all_times[-1] / 1000 * 1e6, wall / 1000 * 1e6, m.all_performances["MSE"][-1]), flush=True)
            if small == "1":
                full = m._engine()._small_partial[256:269].cpu().numpy()
                print("      U phase: staging the other factor %.1f us, row statistics %.1f us, update chains %.1f us" % (
                    (full[11] - full[0]) / 1e3, (full[12] - full[11]) / 1e3, (full[1] - full[12]) / 1e3))
                st = full[:9]
                print("      SM clock during the last sweep: %.0f MHz" % ((full[10] - full[9]) / (st[8] - st[0]) * 1e3))
                print("      stages of the last sweep (us): U phase %.1f | barrier %.1f | V phase %.1f | barrier %.1f | metrics %.1f | "
                      "barrier %.1f | end of sweep %.1f | barrier %.1f" % tuple((st[1:] - st[:-1]) / 1e3), flush=True)

import random
pri3 = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}
for name, f, K, L in (("toy 100x80", "toy_bnmtf_vb.npz", 5, 5), ("GDSC 622x138", "gdsc_bnmtf_vb.npz", 5, 5), ("GDSC 622x138", "gdsc_bnmtf_vb.npz", 7, 7),
                      ("GDSC 622x138", "gdsc_bnmtf_vb.npz", 8, 6), ("GDSC 622x138", "gdsc_bnmtf_vb.npz", 10, 10)):
    d = np.load(os.path.join(G, f))
    R, M = d["R"], d["M"]
    for cls in (bnmtf_b200.bnmtf_gibbs_optimised, bnmtf_b200.bnmtf_vb_optimised, bnmtf_b200.nmtf_icm):
        for small in ("1", "0"):
            os.environ["BNMTF_SMALL"] = small
            np.random.seed(0), random.seed(0)
            m = cls(R, M, K, L, pri3, seed=1)
            m.initialise("random", "random")
            m.run(20)
            torch.cuda.synchronize()
            t0 = time.time()
            m.run(300)
            torch.cuda.synchronize()
            wall = time.time() - t0
            print("%-13s %-22s K=%d L=%d small=%s  device %.1f us/sweep   wall %.1f us/sweep   MSE %.6f" % (
                name, cls.__name__, K, L, small, m.all_times[-1] / 300 * 1e6, wall / 300 * 1e6, m.all_performances["MSE"][-1]), flush=True)
