"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the reference through oracle/ref_shim.py (a temporary py3-rewritten copy under /tmp, never in the
repo), runs its own model classes with fixed numpy / python seeds on the reference's own data
(data_toy/bnmf, data_toy/bnmtf, GDSC IC50) and stores: the inputs, the initial state, scalar traces
per iteration and the state after the last iteration.  The fixtures are what the GPU-side parity tests and
the oracle tests compare against on machines where /root/reference does not exist.
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_shim  # noqa: E402

OUT = HERE


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def seed_all(s=0):
    np.random.seed(s)
    random.seed(s)


def load_inputs(ref):
    root = ref.root
    toy1 = (np.loadtxt(root + "/data_toy/bnmf/R.txt"), np.loadtxt(root + "/data_toy/bnmf/M.txt"))
    toy2 = (np.loadtxt(root + "/data_toy/bnmtf/R.txt"), np.loadtxt(root + "/data_toy/bnmtf/M.txt"))
    import importlib
    gd = importlib.import_module("BNMTF.data_drug_sensitivity.gdsc.load_data")
    (_, X_min, M, _, _, _, _) = quiet(gd.load_gdsc)
    return toy1, toy2, (X_min, M)


def bnmf_priors(I, J, K, lam=0.1):
    return {"alpha": 1.0, "beta": 1.0, "lambdaU": np.ones((I, K)) * lam, "lambdaV": np.ones((J, K)) * lam}


def bnmtf_priors(I, J, K, L, lam=0.1):
    return {"alpha": 1.0, "beta": 1.0, "lambdaF": np.ones((I, K)) * lam,
            "lambdaS": np.ones((K, L)) * lam, "lambdaG": np.ones((J, L)) * lam}


def trace_append(tr, perf, **extra):
    for m in ("MSE", "R^2", "Rp"):
        tr.setdefault(m, []).append(perf[m])
    for k, v in extra.items():
        tr.setdefault(k, []).append(v)


def run_bnmf_vb(ref, R, M, K, its, name, init="random"):
    seed_all(0)
    I, J = R.shape
    m = ref.bnmf_vb_optimised(R, M, K, bnmf_priors(I, J, K))
    quiet(m.initialise, init)
    out = {"R": R, "M": M, "K": K, "lambda": 0.1, "its": its,
           "init_muU": m.muU.copy(), "init_muV": m.muV.copy(), "init_tauU": m.tauU.copy(), "init_tauV": m.tauV.copy(),
           "init_expU": m.expU.copy(), "init_varU": m.varU.copy(), "init_expV": m.expV.copy(), "init_varV": m.varV.copy(),
           "init_exptau": m.exptau, "init_explogtau": m.explogtau, "init_elbo": m.elbo()}
    tr = {}
    for _ in range(its):
        quiet(m.run, 1)
        trace_append(tr, m.predict(M), exptau=m.exptau, explogtau=m.explogtau, elbo=m.elbo(),
                     loglik=m.quality("loglikelihood"), AIC=m.quality("AIC"), BIC=m.quality("BIC"))
    for k in ("expU", "varU", "muU", "tauU", "expV", "varV", "muV", "tauV"):
        out["final_" + k] = getattr(m, k).copy()
    out.update({"trace_" + k: np.array(v) for k, v in tr.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "final MSE", tr["MSE"][-1], "elbo", tr["elbo"][-1])


def run_nmf_icm(ref, R, M, K, its, name, minimum_TN=0.0):
    seed_all(0)
    I, J = R.shape
    m = ref.nmf_icm(R, M, K, bnmf_priors(I, J, K))
    quiet(m.initialise, "random")
    out = {"R": R, "M": M, "K": K, "lambda": 0.1, "its": its, "minimum_TN": minimum_TN,
           "init_U": m.U.copy(), "init_V": m.V.copy(), "init_tau": m.tau}
    quiet(m.run, its, minimum_TN)
    out.update({"final_U": m.U.copy(), "final_V": m.V.copy(), "trace_tau": m.all_tau.copy(),
                "loglik": m.quality("loglikelihood"), "AIC": m.quality("AIC"), "BIC": m.quality("BIC")})
    out.update({"trace_" + k: np.array(v) for k, v in m.all_performances.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "final MSE", m.all_performances["MSE"][-1])


def run_nmf_np(ref, R, M, K, its, name):
    seed_all(0)
    m = ref.NMF(R, M, K)
    m.initialise("exponential", 1.0)
    out = {"R": R, "M": M, "K": K, "its": its, "init_U": m.U.copy(), "init_V": m.V.copy()}
    quiet(m.run, its)
    out.update({"final_U": m.U.copy(), "final_V": m.V.copy(), "final_Idiv": m.compute_I_div()})
    out.update({"trace_" + k: np.array(v) for k, v in m.all_performances.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "final MSE", m.all_performances["MSE"][-1])


def run_bnmf_gibbs_params(ref, R, M, K, name):
    """Deterministic part of Gibbs: (tau,mu) of every conditional for a fixed state, plus one seeded chain
    whose posterior-mean MSE is the distribution-level anchor."""
    seed_all(0)
    I, J = R.shape
    m = ref.bnmf_gibbs_optimised(R, M, K, bnmf_priors(I, J, K))
    quiet(m.initialise, "random")
    out = {"R": R, "M": M, "K": K, "lambda": 0.1, "init_U": m.U.copy(), "init_V": m.V.copy(), "init_tau": m.tau}
    tU = np.array([m.tauU(k) for k in range(K)]).T
    mU = np.array([m.muU(tU[:, k], k) for k in range(K)]).T
    tV = np.array([m.tauV(k) for k in range(K)]).T
    mV = np.array([m.muV(tV[:, k], k) for k in range(K)]).T
    out.update({"tauU": tU, "muU": mU, "tauV": tV, "muV": mV, "alpha_s": m.alpha_s(), "beta_s": m.beta_s()})
    its, burn, thin = 200, 100, 2
    quiet(m.run, its)
    out.update({"chain_its": its, "chain_burn_in": burn, "chain_thinning": thin,
                "chain_trace_MSE": np.array(m.all_performances["MSE"]), "chain_trace_tau": m.all_tau.copy(),
                "chain_quality_MSE": m.quality("MSE", burn, thin), "chain_loglik": m.quality("loglikelihood", burn, thin),
                "chain_exp_tau": m.approx_expectation(burn, thin)[2]})
    # chain-to-chain spread of the reference sampler from the same start (seeds 1..4)
    reps = []
    for s in range(1, 5):
        seed_all(s)
        m2 = ref.bnmf_gibbs_optimised(R, M, K, bnmf_priors(I, J, K))
        m2.U, m2.V = out["init_U"].copy(), out["init_V"].copy()
        m2.tau = m2.alpha_s() / m2.beta_s()
        quiet(m2.run, its)
        reps.append([m2.quality("MSE", burn, thin), m2.approx_expectation(burn, thin)[2]])
    out["chains_MSE_tau"] = np.array(reps)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "chain MSE(posterior mean)", out["chain_quality_MSE"], reps)


def run_bnmtf_vb(ref, R, M, K, L, its, name):
    seed_all(0)
    I, J = R.shape
    m = ref.bnmtf_vb_optimised(R, M, K, L, bnmtf_priors(I, J, K, L))
    quiet(m.initialise, "random", "random")
    out = {"R": R, "M": M, "K": K, "L": L, "lambda": 0.1, "its": its}
    for k in ("muF", "muS", "muG", "tauF", "tauS", "tauG", "expF", "expS", "expG", "varF", "varS", "varG"):
        out["init_" + k] = getattr(m, k).copy()
    out.update({"init_exptau": m.exptau, "init_explogtau": m.explogtau, "init_elbo": m.elbo()})
    tr, orders = {}, {"S": [], "F": [], "G": []}
    for _ in range(its):
        # replay the python-random shuffles of this iteration so the test can feed the same order
        st = random.getstate()
        kl = [(k, l) for k in range(K) for l in range(L)]
        random.shuffle(kl)
        ks = list(range(K))
        random.shuffle(ks)
        ls = list(range(L))
        random.shuffle(ls)
        random.setstate(st)
        orders["S"].append(kl), orders["F"].append(ks), orders["G"].append(ls)
        quiet(m.run, 1)
        trace_append(tr, m.predict(M), exptau=m.exptau, explogtau=m.explogtau, elbo=m.elbo(),
                     loglik=m.quality("loglikelihood"), AIC=m.quality("AIC"), BIC=m.quality("BIC"))
    for k in ("muF", "muS", "muG", "tauF", "tauS", "tauG", "expF", "expS", "expG", "varF", "varS", "varG"):
        out["final_" + k] = getattr(m, k).copy()
    out.update({"trace_" + k: np.array(v) for k, v in tr.items()})
    out.update({"order_" + k: np.array(v) for k, v in orders.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "final MSE", tr["MSE"][-1], "elbo", tr["elbo"][-1])


def run_nmtf_icm(ref, R, M, K, L, its, name, minimum_TN=0.0):
    seed_all(0)
    I, J = R.shape
    m = ref.nmtf_icm(R, M, K, L, bnmtf_priors(I, J, K, L))
    quiet(m.initialise, "random", "random")
    out = {"R": R, "M": M, "K": K, "L": L, "lambda": 0.1, "its": its, "minimum_TN": minimum_TN,
           "init_F": m.F.copy(), "init_S": m.S.copy(), "init_G": m.G.copy(), "init_tau": m.tau}
    quiet(m.run, its, minimum_TN)
    out.update({"final_F": m.F.copy(), "final_S": m.S.copy(), "final_G": m.G.copy(), "trace_tau": m.all_tau.copy(),
                "loglik": m.quality("loglikelihood"), "AIC": m.quality("AIC"), "BIC": m.quality("BIC")})
    out.update({"trace_" + k: np.array(v) for k, v in m.all_performances.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "final MSE", m.all_performances["MSE"][-1])


def run_nmtf_np(ref, R, M, K, L, its, name):
    seed_all(0)
    m = ref.NMTF(R, M, K, L)
    quiet(m.initialise, "exponential", "exponential", 1.0)
    out = {"R": R, "M": M, "K": K, "L": L, "its": its, "init_F": m.F.copy(), "init_S": m.S.copy(), "init_G": m.G.copy()}
    quiet(m.run, its)
    out.update({"final_F": m.F.copy(), "final_S": m.S.copy(), "final_G": m.G.copy(), "final_Idiv": m.compute_I_div()})
    out.update({"trace_" + k: np.array(v) for k, v in m.all_performances.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "final MSE", m.all_performances["MSE"][-1])


def run_bnmtf_gibbs_params(ref, R, M, K, L, name):
    seed_all(0)
    I, J = R.shape
    m = ref.bnmtf_gibbs_optimised(R, M, K, L, bnmtf_priors(I, J, K, L))
    quiet(m.initialise, "random", "random")
    out = {"R": R, "M": M, "K": K, "L": L, "lambda": 0.1,
           "init_F": m.F.copy(), "init_S": m.S.copy(), "init_G": m.G.copy(), "init_tau": m.tau}
    tF = np.array([m.tauF(k) for k in range(K)]).T
    mF = np.array([m.muF(tF[:, k], k) for k in range(K)]).T
    tG = np.array([m.tauG(l) for l in range(L)]).T
    mG = np.array([m.muG(tG[:, l], l) for l in range(L)]).T
    tS = np.array([[m.tauS(k, l) for l in range(L)] for k in range(K)])
    mS = np.array([[m.muS(tS[k, l], k, l) for l in range(L)] for k in range(K)])
    out.update({"tauF": tF, "muF": mF, "tauG": tG, "muG": mG, "tauS": tS, "muS": mS,
                "alpha_s": m.alpha_s(), "beta_s": m.beta_s()})
    its, burn, thin = 200, 100, 2
    quiet(m.run, its)
    out.update({"chain_its": its, "chain_burn_in": burn, "chain_thinning": thin,
                "chain_trace_MSE": np.array(m.all_performances["MSE"]), "chain_trace_tau": m.all_tau.copy(),
                "chain_quality_MSE": m.quality("MSE", burn, thin), "chain_exp_tau": m.approx_expectation(burn, thin)[3]})
    reps = []
    for s in range(1, 5):
        seed_all(s)
        m2 = ref.bnmtf_gibbs_optimised(R, M, K, L, bnmtf_priors(I, J, K, L))
        m2.F, m2.S, m2.G = out["init_F"].copy(), out["init_S"].copy(), out["init_G"].copy()
        m2.tau = m2.alpha_s() / m2.beta_s()
        quiet(m2.run, its)
        reps.append([m2.quality("MSE", burn, thin), m2.approx_expectation(burn, thin)[3]])
    out["chains_MSE_tau"] = np.array(reps)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "chain MSE(posterior mean)", out["chain_quality_MSE"])


def run_distributions(ref, name):
    """Known answers of the scalar/vector distribution helpers + rtnorm samples for two-sample KS tests."""
    seed_all(0)
    mus = np.array([-40.0, -31.0, -29.0, -10.0, -3.0, -1.0, -0.1, 0.0, 0.3, 1.0, 2.5, 8.0, 50.0, 1e-3, -1e3, 4.0])
    taus = np.array([1.0, 1.0, 1.0, 0.5, 2.0, 4.0, 10.0, 1.0, 3.0, 0.25, 1.0, 100.0, 0.01, 1e4, 1e-2, 1e-6])
    out = {"mus": mus, "taus": taus,
           "exp": np.array(ref.tnv.TN_vector_expectation(mus, taus)),
           "var": np.array(ref.tnv.TN_vector_variance(mus, taus)),
           "exp_scalar": np.array([ref.tn.TN_expectation(m, t) for m, t in zip(mus, taus)]),
           "var_scalar": np.array([ref.tn.TN_variance(m, t) for m, t in zip(mus, taus)]),
           "gamma_explog": np.array([ref.gamma.gamma_expectation_log(a, b) for a, b in [(2., 3.), (3601., 5000.), (9., 18.7)]])}
    cases = [(1.0, 1.0 / 9.0), (-2.0, 1.0), (0.5, 4.0), (-8.0, 1.0), (3.0, 1.0), (-1.0, 0.04)]
    n = 4000
    out["draw_cases"] = np.array(cases)
    out["draws"] = np.array([[ref.tn.TN_draw(mu, tau) for _ in range(n)] for mu, tau in cases])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "done")


def main():
    ref = ref_shim.load(rebuild=True)
    (R1, M1), (R2, M2), (Rg, Mg) = load_inputs(ref)
    run_distributions(ref, "distributions")
    run_bnmf_vb(ref, R1, M1, 10, 30, "toy_bnmf_vb")
    run_nmf_icm(ref, R1, M1, 10, 30, "toy_nmf_icm")
    run_nmf_np(ref, R1, M1, 10, 30, "toy_nmf_np")
    run_bnmf_gibbs_params(ref, R1, M1, 10, "toy_bnmf_gibbs")
    run_bnmtf_vb(ref, R2, M2, 5, 5, 30, "toy_bnmtf_vb")
    run_nmtf_icm(ref, R2, M2, 5, 5, 30, "toy_nmtf_icm")
    run_nmtf_np(ref, R2, M2, 5, 5, 30, "toy_nmtf_np")
    run_bnmtf_gibbs_params(ref, R2, M2, 5, 5, "toy_bnmtf_gibbs")
    run_bnmf_vb(ref, Rg, Mg, 10, 10, "gdsc_bnmf_vb")
    run_bnmtf_vb(ref, Rg, Mg, 5, 5, 10, "gdsc_bnmtf_vb")


if __name__ == "__main__":
    main()
