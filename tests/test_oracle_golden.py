"""The CPU oracle against the golden fixtures produced by the reference itself (tests/golden/make_golden.py).

This pins the oracle: every deterministic path (VB, ICM, NP, Gibbs conditional parameters, TN moments,
model-selection metrics) must reproduce the reference's numbers to 1e-9 relative.
"""
import os
import sys

import numpy as np
import pytest

from oracle import bnmtf_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RTOL = 1e-9


def close(a, b, rtol=RTOL, atol=0.0, what=""):
    """rtol elementwise, plus an absolute floor of rtol*1e-3 of the array's largest magnitude (tiny elements of a
    factor matrix cannot be resolved better than the rounding of the big ones they are coupled to)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol + rtol * 1e-3 * scale, err_msg=what)


def priors2(g):
    return {"alpha": 1.0, "beta": 1.0, "lambdaU": float(g["lambda"]), "lambdaV": float(g["lambda"])}


def priors3(g):
    lam = float(g["lambda"])
    return {"alpha": 1.0, "beta": 1.0, "lambdaF": lam, "lambdaS": lam, "lambdaG": lam}


def test_tn_moments_known_answers(golden):
    g = golden("distributions")
    close(orc.tn_expectation(g["mus"], g["taus"]), g["exp"])
    close(orc.tn_variance(g["mus"], g["taus"]), g["var"])
    close(orc.tn_expectation(g["mus"], g["taus"]), g["exp_scalar"])
    close(orc.tn_variance(g["mus"], g["taus"]), g["var_scalar"])
    # reference tests/code/distributions/test_truncated_normal.py: mu<-30 sigma branch -> 1/2000, (1/2000)^2
    assert orc.tn_expectation(-2000.0, 1.0) == pytest.approx(1.0 / 2000.0, rel=1e-15)
    close([orc.gamma_expectation_log(2., 3.), orc.gamma_expectation_log(3601., 5000.), orc.gamma_expectation_log(9., 18.7)],
          g["gamma_explog"])
    assert orc.gamma_expectation_log(2., 3.) == pytest.approx(-0.67582795356964265, abs=1e-14)   # test_gamma.py:18-23


def test_tn_draw_matches_rtnorm_in_distribution(golden):
    from scipy.stats import ks_2samp
    g = golden("distributions")
    rng = np.random.RandomState(1)
    for (mu, tau), ref_draws in zip(g["draw_cases"], g["draws"]):
        ours = orc.tn_draw(np.full(4000, mu), np.full(4000, tau), rng)
        assert (ours >= 0).all()
        assert ks_2samp(ours, ref_draws).pvalue > 1e-3, (mu, tau)
    assert orc.tn_draw(np.array([1.0]), np.array([0.0]), rng)[0] == 0.0   # test_truncated_normal.py:50-54


@pytest.mark.parametrize("name", ["toy_bnmf_vb", "gdsc_bnmf_vb"])
def test_bnmf_vb_trajectory(golden, name):
    g = golden(name)
    o = orc.OracleBNMF(g["R"], g["M"], int(g["K"]), priors2(g), mode="vb")
    o.init_vb(g["init_muU"], g["init_muV"], g["init_tauU"], g["init_tauV"])
    close(o.U, g["init_expU"]), close(o.varV, g["init_varV"])
    close(o.exptau, g["init_exptau"]), close(o.explogtau, g["init_explogtau"]), close(o.elbo(), g["init_elbo"])
    for it in range(int(g["its"])):
        perf = o.sweep()
        for m in ("MSE", "R^2", "Rp"):
            close(perf[m], g["trace_" + m][it], what="%s it %d" % (m, it))
        close(o.exptau, g["trace_exptau"][it]), close(o.elbo(), g["trace_elbo"][it], what="elbo it %d" % it)
        close(o.quality("AIC"), g["trace_AIC"][it]), close(o.quality("BIC"), g["trace_BIC"][it])
    for k in ("U", "V"):
        close(getattr(o, k), g["final_exp" + k]), close(getattr(o, "var" + k), g["final_var" + k])
        close(getattr(o, "mu" + k), g["final_mu" + k]), close(getattr(o, "tau" + k), g["final_tau" + k])


def test_nmf_icm_trajectory(golden):
    g = golden("toy_nmf_icm")
    o = orc.OracleBNMF(g["R"], g["M"], int(g["K"]), priors2(g), mode="icm")
    o.set_state(g["init_U"], g["init_V"])
    close(o.tau, g["init_tau"])
    for it in range(int(g["its"])):
        perf = o.sweep(float(g["minimum_TN"]))
        close(perf["MSE"], g["trace_MSE"][it]), close(o.tau, g["trace_tau"][it])
    close(o.U, g["final_U"]), close(o.V, g["final_V"])
    close(o.quality("loglikelihood"), g["loglik"]), close(o.quality("AIC"), g["AIC"]), close(o.quality("BIC"), g["BIC"])


def test_nmf_np_trajectory(golden):
    g = golden("toy_nmf_np")
    o = orc.OracleBNMF(g["R"], g["M"], int(g["K"]), mode="np")
    o.set_state(g["init_U"], g["init_V"])
    for it in range(int(g["its"])):
        perf = o.sweep()
        close(perf["MSE"], g["trace_MSE"][it]), close(perf["Rp"], g["trace_Rp"][it])
    close(o.U, g["final_U"]), close(o.V, g["final_V"]), close(o.i_divergence(), g["final_Idiv"])


def test_bnmf_gibbs_conditionals(golden):
    g = golden("toy_bnmf_gibbs")
    o = orc.OracleBNMF(g["R"], g["M"], int(g["K"]), priors2(g), mode="gibbs")
    o.set_state(g["init_U"], g["init_V"])
    close(o.tau, g["init_tau"]), close(o.alpha_s(), g["alpha_s"]), close(o.beta_s(), g["beta_s"])
    for k in range(int(g["K"])):
        t, m = o.column_params(k, "U")
        close(t, g["tauU"][:, k]), close(m, g["muU"][:, k])
        t, m = o.column_params(k, "V")
        close(t, g["tauV"][:, k]), close(m, g["muV"][:, k])


def test_bnmf_gibbs_chain_distribution(golden):
    """Matched burn-in / thinning: posterior-mean MSE and E[tau] inside the reference's own chain-to-chain band
    (4 reference chains from the same start; tolerance = max(5 sigma, 5 %))."""
    g = golden("toy_bnmf_gibbs")
    o = orc.OracleBNMF(g["R"], g["M"], int(g["K"]), priors2(g), mode="gibbs", seed=3)
    o.set_state(g["init_U"], g["init_V"])
    its, burn, thin = int(g["chain_its"]), int(g["chain_burn_in"]), int(g["chain_thinning"])
    Us, Vs, taus = [], [], []
    for _ in range(its):
        o.sweep()
        Us.append(o.U.copy()), Vs.append(o.V.copy()), taus.append(o.tau)
    idx = range(burn, its, thin)
    eU, eV = np.mean([Us[i] for i in idx], axis=0), np.mean([Vs[i] for i in idx], axis=0)
    mse = o.predict(g["M"], eU, eV)["MSE"]
    ref_mse, ref_tau = g["chains_MSE_tau"][:, 0], g["chains_MSE_tau"][:, 1]
    assert abs(mse - ref_mse.mean()) <= max(5 * ref_mse.std(), 0.05 * ref_mse.mean())
    assert abs(np.mean([taus[i] for i in idx]) - ref_tau.mean()) <= max(5 * ref_tau.std(), 0.05 * ref_tau.mean())


@pytest.mark.parametrize("name", ["toy_bnmtf_vb", "gdsc_bnmtf_vb"])
def test_bnmtf_vb_trajectory(golden, name):
    g = golden(name)
    o = orc.OracleBNMTF(g["R"], g["M"], int(g["K"]), int(g["L"]), priors3(g), mode="vb")
    o.init_vb(g["init_muF"], g["init_muS"], g["init_muG"],
              {"F": g["init_tauF"], "S": g["init_tauS"], "G": g["init_tauG"]})
    close(o.F, g["init_expF"]), close(o.varS, g["init_varS"]), close(o.elbo(), g["init_elbo"])
    for it in range(int(g["its"])):
        order = {"S": [tuple(x) for x in g["order_S"][it]], "F": list(g["order_F"][it]), "G": list(g["order_G"][it])}
        perf = o.sweep(order=order)
        close(perf["MSE"], g["trace_MSE"][it], what="MSE it %d" % it)
        close(o.exptau, g["trace_exptau"][it]), close(o.elbo(), g["trace_elbo"][it], what="elbo it %d" % it)
    # VB-NMTF amplifies rounding differences: two fp64 CPU evaluations of the same formulas (reference vs oracle,
    # different association order) agree to 2e-12 after one sweep but only to ~2e-8 in the factors after 30
    # sweeps on the toy data, while the scalar traces above stay within 1e-9.  Hence 1e-6 on the final factors.
    for k in "FSG":
        close(getattr(o, k), g["final_exp" + k], rtol=1e-6), close(getattr(o, "var" + k), g["final_var" + k], rtol=1e-6)
        close(getattr(o, "mu" + k), g["final_mu" + k], rtol=1e-6), close(getattr(o, "tau" + k), g["final_tau" + k], rtol=1e-6)


def test_nmtf_icm_trajectory(golden):
    g = golden("toy_nmtf_icm")
    o = orc.OracleBNMTF(g["R"], g["M"], int(g["K"]), int(g["L"]), priors3(g), mode="icm")
    o.set_state(g["init_F"], g["init_S"], g["init_G"])
    close(o.tau, g["init_tau"])
    for it in range(int(g["its"])):
        perf = o.sweep(float(g["minimum_TN"]))
        close(perf["MSE"], g["trace_MSE"][it]), close(o.tau, g["trace_tau"][it])
    close(o.F, g["final_F"]), close(o.S, g["final_S"]), close(o.G, g["final_G"])


def test_nmtf_np_trajectory(golden):
    g = golden("toy_nmtf_np")
    o = orc.OracleBNMTF(g["R"], g["M"], int(g["K"]), int(g["L"]), mode="np")
    o.set_state(g["init_F"], g["init_S"], g["init_G"])
    for it in range(int(g["its"])):
        perf = o.sweep()
        close(perf["MSE"], g["trace_MSE"][it])
    close(o.F, g["final_F"]), close(o.S, g["final_S"]), close(o.G, g["final_G"]), close(o.i_divergence(), g["final_Idiv"])


def test_bnmtf_gibbs_conditionals(golden):
    g = golden("toy_bnmtf_gibbs")
    K, L = int(g["K"]), int(g["L"])
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="gibbs")
    o.set_state(g["init_F"], g["init_S"], g["init_G"])
    close(o.tau, g["init_tau"]), close(o.beta_s(), g["beta_s"])
    for k in range(K):
        t, m = o.params_F(k)
        close(t, g["tauF"][:, k]), close(m, g["muF"][:, k])
        for l in range(L):
            t, m = o.params_S(k, l)
            close(t, g["tauS"][k, l]), close(m, g["muS"][k, l])
    for l in range(L):
        t, m = o.params_G(l)
        close(t, g["tauG"][:, l]), close(m, g["muG"][:, l])


# ---- conditioning of the reference's own formulas (what a 1e-9 comparison can and cannot mean) ----------------
def test_tn_moment_formulas_amplify_the_last_bits_of_x():
    """truncated_normal_vector.py:53-73 evaluates lambda = pdf(x) / (0.5 erfc(x / sqrt 2)) from exp(-x*x/2) and
    exp(-(x/sqrt2)^2): two roundings of the same exponent (~450 at x = 30, ulp 6e-14).  sigma (lambda - x) and
    sigma^2 (1 - lambda (lambda - x)) then cancel, so the reference's mean / variance of a strongly truncated entry move by
    up to ~2.5e-13 x^2 / x^4 when mu changes in its LAST BIT.  The GPU tests compare such entries with the tolerance
    1e-9 + 3e-13 x^power (tests/test_bnmtf_gpu.py::_tn_close, bench.py's variance note); this test pins both facts: the
    noise is real (> 1e-8 on the variance beyond x = 20) and the bound holds."""
    rng = np.random.RandomState(0)
    x = rng.uniform(0.0, 29.99, 400000)
    tau = rng.uniform(0.1, 10.0, x.size)
    mu = -x / np.sqrt(tau)
    mu1 = mu * (1.0 + 1.1e-16 * rng.choice([-1.0, 1.0], x.size))          # rounds to mu itself or to a neighbour
    with np.errstate(all="ignore"):
        v0, v1 = orc.tn_variance(mu, tau), orc.tn_variance(mu1, tau)
        e0, e1 = orc.tn_expectation(mu, tau), orc.tn_expectation(mu1, tau)
    xs = -mu * np.sqrt(tau)
    rv, re = np.abs(v1 / v0 - 1.0), np.abs(e1 / e0 - 1.0)
    assert rv[xs > 20].max() > 1e-8 and re[xs > 20].max() > 3e-11
    assert (rv <= 1e-11 + 3e-13 * xs ** 4).all() and (re <= 1e-12 + 3e-13 * xs ** 2).all()
    assert rv[xs < 10].max() < 1e-9


def test_vb_nmtf_trajectory_sensitivity_fixture():
    """tests/golden/vb_nmtf_sensitivity.json (made by make_sensitivity.py) says how far the CPU trajectory of the VB
    tri-factorisation moves when its start changes in the last bit.  Re-derive it for one other perturbation on the toy
    data: same order of magnitude, and -- the point -- the factors are NOT reproducible to 1e-9 even on the CPU."""
    import json
    sys.path.insert(0, GOLDEN)
    import make_sensitivity as ms
    with open(os.path.join(GOLDEN, "vb_nmtf_sensitivity.json")) as fh:
        stored = json.load(fh)
    g = dict(np.load(os.path.join(GOLDEN, "toy_bnmtf_vb.npz")))
    cur = ms.spread(g, seeds=(11,))
    for k in ms.KEYS:
        assert cur[k] <= 10.0 * stored["toy_bnmtf_vb"][k] + 1e-12, (k, cur[k], stored["toy_bnmtf_vb"][k])
    assert stored["toy_bnmtf_vb"]["expF"] > 1e-9 and cur["expF"] > 1e-10
