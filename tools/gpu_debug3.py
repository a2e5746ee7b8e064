"""Measure (not assert) VB-NMTF device-vs-oracle deviations: per sweep from the oracle's state, and free running,
for the golden fixtures (K = L) and for rectangular (K, L) from random and K-means starts."""
import os, random, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bnmtf_b200
from oracle import bnmtf_oracle as orc

PRI = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.max(np.abs(b))))
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-11 * scale)))


def set_state(m, o):
    for k in "FSG":
        setattr(m, "exp" + k, getattr(o, k).copy()), setattr(m, "var" + k, getattr(o, "var" + k).copy())
        setattr(m, "mu" + k, getattr(o, "mu" + k).copy()), setattr(m, "tau" + k, getattr(o, "tau" + k).copy())
    m.exptau, m.explogtau, m.alpha_s, m.beta_s = o.exptau, o.explogtau, o.alpha_s_, o.beta_s_


def study(tag, R, M, K, L, init_S, init_FG, its, seed):
    np.random.seed(seed), random.seed(seed)
    m = bnmtf_b200.bnmtf_vb_optimised(R, M, K, L, PRI)
    m.initialise(init_S, init_FG)
    mu0 = (m.muF.copy(), m.muS.copy(), m.muG.copy())
    orders = [orc.OracleBNMTF.shuffled_order(K, L) for _ in range(its)]
    dev_order = lambda od: {"S": [k * L + l for k, l in od["S"]], "F": od["F"], "G": od["G"]}
    # per sweep from the oracle's state
    o = orc.OracleBNMTF(R, M, K, L, PRI, mode="vb")
    o.init_vb(*mu0)
    worst = {}
    for it in range(its):
        set_state(m, o)
        eng = m._push()
        eng.alloc_trace(1)
        perf = o.sweep(order=orders[it])
        eng.sweep(order=dev_order(orders[it]))
        tr = eng.trace.cpu().numpy()[0]
        m._pull(eng)
        for k in "FSG":
            for a, b in (("exp", k), ("var", "var" + k), ("tau", "tau" + k), ("mu", "mu" + k)):
                worst[a + k] = max(worst.get(a + k, 0.0), rel(getattr(m, a + k), getattr(o, b)))
        worst["MSE"] = max(worst.get("MSE", 0.0), rel(tr[1], perf["MSE"]))
    print(tag, "per-sweep worst:", " ".join("%s=%.1e" % kv for kv in sorted(worst.items())))
    # free running
    o = orc.OracleBNMTF(R, M, K, L, PRI, mode="vb")
    o.init_vb(*mu0)
    set_state(m, o)
    eng = m._push()
    eng.alloc_trace(its)
    mse_o = []
    for it in range(its):
        mse_o.append(o.sweep(order=orders[it])["MSE"])
        eng.sweep(order=dev_order(orders[it]))
    tr = eng.trace.cpu().numpy()[:its]
    d = np.abs(tr[:, 1] / np.array(mse_o) - 1)
    m._pull(eng)
    print(tag, "free-running MSE rel by sweep:", " ".join("%.1e" % v for v in d), "| final expF %.1e expS %.1e expG %.1e"
          % (rel(m.expF, o.F), rel(m.expS, o.S), rel(m.expG, o.G)))


if __name__ == "__main__":
    a = np.load("tests/golden/toy_bnmtf_vb.npz")
    b = np.load("tests/golden/gdsc_bnmtf_vb.npz")
    for K, L in ((5, 5), (5, 4), (4, 6), (3, 7)):
        study("toy  K=%d L=%d random/kmeans" % (K, L), a["R"], a["M"], K, L, "random", "kmeans", 8, 2)
        study("toy  K=%d L=%d random/random" % (K, L), a["R"], a["M"], K, L, "random", "random", 8, 2)
    study("gdsc K=5 L=5 random/kmeans", b["R"], b["M"], 5, 5, "random", "kmeans", 8, 2)
    study("gdsc K=6 L=4 random/random", b["R"], b["M"], 6, 4, "random", "random", 8, 2)
