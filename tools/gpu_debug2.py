import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnmtf_b200
from oracle import bnmtf_oracle as orc
g = dict(np.load("tests/golden/gdsc_bnmtf_vb.npz"))
K, L = 5, 5
pri = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}
m = bnmtf_b200.bnmtf_vb_optimised(g["R"], g["M"], K, L, pri)
o = orc.OracleBNMTF(g["R"], g["M"], K, L, pri, mode="vb")
o.init_vb(g["init_muF"], g["init_muS"], g["init_muG"], {"F": g["init_tauF"], "S": g["init_tauS"], "G": g["init_tauG"]})
for k in "FSG":
    setattr(m, "exp" + k, getattr(o, k).copy()); setattr(m, "var" + k, getattr(o, "var" + k).copy())
    setattr(m, "mu" + k, getattr(o, "mu" + k).copy()); setattr(m, "tau" + k, getattr(o, "tau" + k).copy())
m.exptau, m.explogtau, m.alpha_s, m.beta_s = o.exptau, o.explogtau, o.alpha_s_, o.beta_s_
eng = m._push()
oS = [tuple(int(v) for v in x) for x in g["order_S"][0]]; oF = [int(x) for x in g["order_F"][0]]
print("order F", oF)
eng.stats_rows()
eng.phase_S([k * L + l for k, l in oS])
for k, l in oS:
    o.vb_update_S(k, l)
    o.S[k, l], o.varS[k, l] = orc.tn_expectation(o.muS[k, l], o.tauS[k, l]), orc.tn_variance(o.muS[k, l], o.tauS[k, l])
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
print("S: mu rel", rel(m._down(eng.S["mu"]), o.muS).max(), "tau rel", rel(m._down(eng.S["tauf"]), o.tauS).max(),
      "exp rel", rel(m._down(eng.S["fac"]), o.S).max(), "exp abs", np.abs(m._down(eng.S["fac"]) - o.S).max(), "maxS", o.S.max())
print("expS dev\n", m._down(eng.S["fac"]), "\nexpS oracle\n", o.S)
eng.phase_F(oF)
for k in oF:
    o.vb_update_F(k)
    o.F[:, k], o.varF[:, k] = orc.tn_expectation(o.muF[:, k], o.tauF[:, k]), orc.tn_variance(o.muF[:, k], o.tauF[:, k])
print("F mu rel per column", rel(m._down(eng.F.mu), o.muF).max(axis=0))
print("F tau rel per column", rel(m._down(eng.F.tauf), o.tauF).max(axis=0))
print("F exp abs per column", np.abs(m._down(eng.F.fac) - o.F).max(axis=0), "max", o.F.max(axis=0))
