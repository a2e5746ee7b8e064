"""Ad-hoc GPU diagnostics used during bring-up (not a test)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnmtf_b200
from oracle import bnmtf_oracle as orc
g = dict(np.load("tests/golden/toy_bnmf_vb.npz"))
pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
m = bnmtf_b200.bnmf_vb_optimised(g["R"], g["M"], 10, pri)
m.initialise("exp")
m.muU, m.muV = g["init_muU"].copy(), g["init_muV"].copy()
for k in range(10): m.update_exp_U(k)
for k in range(10): m.update_exp_V(k)
print("expU err", np.abs(m.expU - g["init_expU"]).max(), "varV err", np.abs(m.varV - g["init_varV"]).max())
print("esd", m.exp_square_diff())
o = orc.OracleBNMF(g["R"], g["M"], 10, pri, mode="vb"); o.init_vb(g["init_muU"], g["init_muV"])
print("oracle esd", o.exp_square_diff(), "elbo", o.elbo())
m.update_tau(); m.update_exp_tau()
print("exptau", m.exptau, g["init_exptau"], "elbo", m.elbo(), g["init_elbo"])
m.update_U(0); tU, mU = o.column_params(0, "U")
print("update_U(0) tau err", np.abs(m.tauU[:,0]/tU-1).max(), "mu err", np.abs(m.muU[:,0]-mU).max())
m.run(5)
print("MSE", m.all_performances["MSE"], "\nref", g["trace_MSE"][:5])
print("elbo", m.all_elbo, "\nref", g["trace_elbo"][:5])
print("times", m.all_times)
