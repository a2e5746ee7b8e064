import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnmtf_b200 as models
from oracle import bnmtf_oracle as orc
def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.max(np.abs(b))))
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-2 * scale)))
for (I, J, K, frac) in [(64, 1030, 31, 0.45)]:
    rng = np.random.RandomState(I + J)
    R = np.abs(rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J)) + 3.0)
    M = (rng.rand(I, J) >= frac).astype(float)
    M[np.arange(I), rng.randint(0, J, I)] = 1.0
    M[rng.randint(0, I, J), np.arange(J)] = 1.0
    pri = {"alpha": 2.0, "beta": 0.5, "lambdaU": 0.3, "lambdaV": 0.7}
    for solver in ("umma", "dmma"):
        os.environ["BNMTF_GRAM"] = solver; os.environ["BNMTF_RX"] = solver
        np.random.seed(4)
        m = models.bnmf_vb_optimised(R, M, K, pri)
        m.initialise("random")
        o = orc.OracleBNMF(R, M, K, pri, mode="vb")
        o.init_vb(m.muU.copy(), m.muV.copy())
        out = []
        for it in range(8):
            mse = o.sweep()["MSE"]
            m.run(1)
            out.append("%.0e/%.0e/%.0e" % (rel(m.all_performances["MSE"][-1], mse), rel(m.expU, o.U), rel(m.varU, o.varU)))
        print(I, J, K, frac, solver or "auto", "VB mse/expU/varU per sweep:", " ".join(out), "| elbo", rel(m.quality("ELBO"), o.elbo()), flush=True)
    np.random.seed(5)
    c = models.nmf_icm(R, M, K, pri)
    c.initialise("random")
    oc = orc.OracleBNMF(R, M, K, pri, mode="icm")
    oc.set_state(c.U.copy(), c.V.copy(), tau=c.tau)
    out = []
    for it in range(8):
        oc.sweep(minimum_TN=0.05)
        c.run(1, minimum_TN=0.05)
        out.append("%.0e/%.0e" % (rel(c.U, oc.U), rel(c.tau, oc.tau)))
    print(I, J, K, frac, "ICM U/tau per sweep:", " ".join(out), flush=True)
