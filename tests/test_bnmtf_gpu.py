"""GPU parity of the tri-factorisation models against the golden fixtures (generated from the reference) and the
CPU oracle."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-9, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(1.0, float(np.max(np.abs(b)))) if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * 1e-2 * scale, err_msg=what)


def priors3(g):
    lam = float(g["lambda"])
    return {"alpha": 1.0, "beta": 1.0, "lambdaF": lam, "lambdaS": lam, "lambdaG": lam}


def vb_from_golden(g):
    import bnmtf_b200
    K, L = int(g["K"]), int(g["L"])
    m = bnmtf_b200.bnmtf_vb_optimised(g["R"], g["M"], K, L, priors3(g))
    m.initialise("exp", "exp")
    m.muF, m.muS, m.muG = g["init_muF"].copy(), g["init_muS"].copy(), g["init_muG"].copy()
    m.tauF, m.tauS, m.tauG = g["init_tauF"].copy(), g["init_tauS"].copy(), g["init_tauG"].copy()
    for k in range(K):
        m.update_exp_F(k)
    for k in range(K):
        for l in range(L):
            m.update_exp_S(k, l)
    for l in range(L):
        m.update_exp_G(l)
    m.update_tau()
    m.update_exp_tau()
    return m


@pytest.mark.parametrize("name", ["toy_bnmtf_vb", "gdsc_bnmtf_vb"])
def test_vb_initial_state_and_single_updates(golden, name):
    from oracle import bnmtf_oracle as orc
    g = golden(name)
    K, L = int(g["K"]), int(g["L"])
    m = vb_from_golden(g)
    close(m.expF, g["init_expF"]), close(m.varS, g["init_varS"]), close(m.expG, g["init_expG"])
    close(m.exptau, g["init_exptau"]), close(m.explogtau, g["init_explogtau"])
    if np.isfinite(g["init_elbo"]):
        close(m.elbo(), g["init_elbo"])
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="vb")
    o.init_vb(g["init_muF"], g["init_muS"], g["init_muG"], {"F": g["init_tauF"], "S": g["init_tauS"], "G": g["init_tauG"]})
    close(m.exp_square_diff(), o.exp_square_diff(), rtol=1e-11)
    for k, l in ((0, 0), (K - 1, 1)):
        m.update_S(k, l), o.vb_update_S(k, l)
        close(m.tauS[k, l], o.tauS[k, l], rtol=1e-11), close(m.muS[k, l], o.muS[k, l], rtol=1e-10)
    m.update_F(1), o.vb_update_F(1)
    close(m.tauF[:, 1], o.tauF[:, 1], rtol=1e-11), close(m.muF[:, 1], o.muF[:, 1], rtol=1e-10)
    m.update_G(L - 1), o.vb_update_G(L - 1)
    close(m.tauG[:, L - 1], o.tauG[:, L - 1], rtol=1e-11), close(m.muG[:, L - 1], o.muG[:, L - 1], rtol=1e-10)


def _set_vb_state(m, o):
    for k in "FSG":
        setattr(m, "exp" + k, getattr(o, k).copy()), setattr(m, "var" + k, getattr(o, "var" + k).copy())
        setattr(m, "mu" + k, getattr(o, "mu" + k).copy()), setattr(m, "tau" + k, getattr(o, "tau" + k).copy())
    m.exptau, m.explogtau, m.alpha_s, m.beta_s = o.exptau, o.explogtau, o.alpha_s_, o.beta_s_


def _tn_close(m, o, k, what):
    """mu, tau at 1e-9; exp / var entry by entry with the truncation-dependent tolerance of the REFERENCE's own formula:
    lambda = pdf(x) / (0.5 erfc(x / sqrt 2)) is built from exp(-x*x/2) and exp(-(x/sqrt2)^2), whose argument roundings
    differ, so sigma (lambda - x) and sigma^2 (1 - lambda (lambda - x)) depend on the last bits of x = -mu sqrt(tau) with
    amplitudes 2.5e-13 x^2 and 2.5e-13 x^4 (tests/test_oracle_golden.py::test_tn_moment_formulas_amplify_the_last_bits_of_x
    measures exactly that on the CPU).  Returns the number of entries that needed more than 1e-9."""
    mu_o, tau_o = getattr(o, "mu" + k), getattr(o, "tau" + k)
    close(getattr(m, "mu" + k), mu_o, rtol=1e-9, what="mu%s %s" % (k, what))
    close(getattr(m, "tau" + k), tau_o, rtol=1e-9, what="tau%s %s" % (k, what))
    with np.errstate(all="ignore"):
        x = np.where(mu_o < -30.0 / np.sqrt(tau_o), 0.0, np.maximum(-mu_o * np.sqrt(tau_o), 0.0))   # limit branch: no cancellation
    loose = 0
    for q, power, ref in (("exp", 2, getattr(o, k)), ("var", 4, getattr(o, "var" + k))):
        a = getattr(m, q + k)
        scale = float(np.abs(ref).max())
        tol = 1e-9 + 3e-13 * x ** power
        err = np.abs(a - ref) / (np.abs(ref) + 1e-2 * scale)
        bad = err > tol
        assert not bad.any(), "%s%s %s: %d entries beyond 1e-9 + 3e-13 x^%d, worst %.2e at x = %.1f" % (
            q, k, what, int(bad.sum()), power, float(err.max()), float(x.flat[int(np.argmax(err))]))
        loose += int((err > 1e-9).sum())
        assert not (err > 1e-9)[x <= 10.0].any(), "%s%s %s: an entry with x <= 10 is off by more than 1e-9" % (q, k, what)
    return loose


@pytest.mark.parametrize("name", ["toy_bnmtf_vb", "gdsc_bnmtf_vb"])
def test_vb_every_phase_matches_oracle_from_the_same_state(golden, name):
    """Phase-by-phase parity without error accumulation: before every phase (all S entries / all F columns / all G
    columns, in the reference's shuffled orders) the device state is set to the oracle's, then both run the phase and
    every variational parameter of the updated factor is compared at 1e-9 -- except that exp / var of strongly
    truncated entries get the x-dependent tolerance of _tn_close, because there the REFERENCE's value is itself a
    function of the last bits of its input.  (Within a phase the updated factor's variances are not read, so that
    noise cannot leak into the mu / tau compared here; across phases it does -- varS enters tauF directly -- which is
    why the free-running trajectory test below cannot be a 1e-9 test on every quantity.)"""
    from oracle import bnmtf_oracle as orc
    g = golden(name)
    K, L = int(g["K"]), int(g["L"])
    m = vb_from_golden(g)
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="vb")
    o.init_vb(g["init_muF"], g["init_muS"], g["init_muG"], {"F": g["init_tauF"], "S": g["init_tauS"], "G": g["init_tauG"]})
    loose = 0
    for it in range(min(8, int(g["its"]))):
        oS = [tuple(int(v) for v in x) for x in g["order_S"][it]]
        oF, oG = [int(x) for x in g["order_F"][it]], [int(x) for x in g["order_G"][it]]
        for phase in "SFG":
            _set_vb_state(m, o)
            eng = m._push()
            if phase == "S":
                for k, l in oS:
                    o.vb_update_S(k, l)
                    o.S[k, l], o.varS[k, l] = orc.tn_expectation(o.muS[k, l], o.tauS[k, l]), orc.tn_variance(o.muS[k, l], o.tauS[k, l])
                eng.stats_rows()
                eng.phase_S([k * L + l for k, l in oS])
            elif phase == "F":
                for k in oF:
                    o.vb_update_F(k)
                    o.F[:, k], o.varF[:, k] = orc.tn_expectation(o.muF[:, k], o.tauF[:, k]), orc.tn_variance(o.muF[:, k], o.tauF[:, k])
                eng.stats_rows()
                eng.phase_F(oF)
            else:
                for l in oG:
                    o.vb_update_G(l)
                    o.G[:, l], o.varG[:, l] = orc.tn_expectation(o.muG[:, l], o.tauG[:, l]), orc.tn_variance(o.muG[:, l], o.tauG[:, l])
                eng.stats_cols()
                eng.phase_G(oG)
            m._pull(eng)
            loose += _tn_close(m, o, phase, "it %d" % it)
        o.update_tau_vb()
        # the sweep's scalars from the oracle's end-of-sweep state: exp_square_diff, tau, MSE, ELBO
        _set_vb_state(m, o)
        close(m.exp_square_diff(), o.exp_square_diff(), rtol=1e-10, what="exp_square_diff it %d" % it)
        if np.isfinite(o.elbo()):
            close(m.elbo(), o.elbo(), rtol=1e-10, what="elbo it %d" % it)
    print("%s: %d exp/var entries needed the truncation-dependent tolerance" % (name, loose))


def _trajectory_tolerances(golden):
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "vb_nmtf_sensitivity.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["toy_bnmtf_vb", "gdsc_bnmtf_vb"])
def test_vb_trajectory_matches_reference(golden, name):
    """Free-running trajectory against the reference's golden run, replaying its python-random shuffles.  Tolerance:
    1e-9, widened to 3x the reference's OWN reproducibility where that is worse -- tests/golden/vb_nmtf_sensitivity.json
    holds, per quantity, how far the CPU trajectory moves when its start is perturbed in the last bit (made by
    tests/golden/make_sensitivity.py, re-derived on the CPU by tests/test_oracle_golden.py): on the toy data the traces
    move by ~1e-9 and the factors by ~1e-8 (the TN-moment noise of _tn_close, fed back 30 times); GDSC stays below 1e-9."""
    g = golden(name)
    sens = _trajectory_tolerances(golden)[name]
    tol = lambda key: max(1e-9, 3.0 * sens[key])
    m = vb_from_golden(g)
    its = int(g["its"])
    eng = m._push()
    eng.alloc_trace(its)
    L = int(g["L"])
    for it in range(its):
        order = {"S": [int(k) * L + int(l) for k, l in g["order_S"][it]], "F": [int(x) for x in g["order_F"][it]],
                 "G": [int(x) for x in g["order_G"][it]]}
        eng.sweep(order=order)
    tr = eng.trace.cpu().numpy()[:its]
    close(tr[:, 1], g["trace_MSE"], rtol=tol("MSE"), what="MSE trace")
    close(tr[:, 0], g["trace_exptau"], rtol=tol("exptau"), what="exptau trace")
    ok = np.isfinite(g["trace_elbo"])
    close(tr[ok, 4], g["trace_elbo"][ok], rtol=tol("elbo"), what="ELBO trace")
    m._pull(eng)
    for k in "FSG":
        close(getattr(m, "exp" + k), g["final_exp" + k], rtol=tol("exp" + k), what="exp" + k)


@pytest.mark.parametrize("K,L,init_FG", [(5, 4, "kmeans"), (4, 6, "kmeans"), (3, 7, "random"), (6, 2, "random")])
def test_vb_rectangular_core_matches_oracle(golden, K, L, init_FG):
    """K != L (the golden trajectories are all K = L = 5): host initialisation with the reference's draw order (numpy
    exponentials for S, python-random K-means for F and G), then 8 free-running sweeps with the reference's shuffles,
    against the oracle from the same start.  Measured 4e-10 on the MSE trace, 2e-8 on the factors."""
    from oracle import bnmtf_oracle as orc
    import bnmtf_b200
    g = golden("toy_bnmtf_vb")
    np.random.seed(2), random.seed(2)
    m = bnmtf_b200.bnmtf_vb_optimised(g["R"], g["M"], K, L, priors3(g))
    m.initialise("random", init_FG)
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="vb")
    o.init_vb(m.muF.copy(), m.muS.copy(), m.muG.copy())
    close(m.expF, o.F), close(m.expS, o.S), close(m.expG, o.G), close(m.exptau, o.exptau)
    random.seed(9)
    orders = [orc.OracleBNMTF.shuffled_order(K, L) for _ in range(8)]
    mse = [o.sweep(order=od)["MSE"] for od in orders]
    random.seed(9)
    m.run(8)
    close(m.all_performances["MSE"], mse, rtol=1e-8, what="MSE trace")
    close(m.all_exp_tau[-1], o.exptau, rtol=1e-8)
    for k in "FSG":
        close(getattr(m, "exp" + k), getattr(o, k), rtol=1e-6, what="exp" + k)
    close(m.quality("ELBO"), o.elbo(), rtol=1e-7) if np.isfinite(o.elbo()) else None


def test_large_core_uses_the_global_memory_accumulator(golden):
    """K * L = 13 * 14 = 182: the (KL x KL) normal matrix of the S phase no longer fits shared memory (limit ~158): the
    tiled reduction covers it in several passes over the rows, each CTA writing its blocks to global memory, and the
    coordinate chain reads H from global memory (csrc/nmtf.cu, k_nmtf_sq_tiled with gridDim.y > 1, k_coord_solve<false>) --
    the reference has no size limit and its grid searches go to K, L of 20-30.  Two VB sweeps and two ICM sweeps against
    the oracle."""
    from oracle import bnmtf_oracle as orc
    import bnmtf_b200
    g = golden("toy_bnmtf_vb")
    K, L = 13, 14
    np.random.seed(4), random.seed(4)
    m = bnmtf_b200.bnmtf_vb_optimised(g["R"], g["M"], K, L, priors3(g))
    m.initialise("random", "random")
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="vb")
    o.init_vb(m.muF.copy(), m.muS.copy(), m.muG.copy())
    random.seed(5)
    orders = [orc.OracleBNMTF.shuffled_order(K, L) for _ in range(2)]
    mse = [o.sweep(order=od)["MSE"] for od in orders]
    random.seed(5)
    m.run(2)
    close(m.all_performances["MSE"], mse, rtol=1e-9, what="MSE trace")
    close(m.muS, o.muS, rtol=1e-8), close(m.tauS, o.tauS, rtol=1e-8), close(m.expF, o.F, rtol=1e-8)
    np.random.seed(6)
    c = bnmtf_b200.nmtf_icm(g["R"], g["M"], K, L, priors3(g))
    c.initialise("random", "random")
    oc = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="icm")
    oc.set_state(c.F.copy(), c.S.copy(), c.G.copy(), tau=c.tau)
    c.run(2, minimum_TN=0.1)
    for _ in range(2):
        oc.sweep(minimum_TN=0.1)
    close(c.S, oc.S, rtol=1e-9), close(c.F, oc.F, rtol=1e-9), close(c.tau, oc.tau, rtol=1e-10)


def test_vb_run_uses_python_random_like_reference(golden):
    g = golden("toy_bnmtf_vb")
    a, b = vb_from_golden(g), vb_from_golden(g)
    random.seed(3)
    a.run(3)
    random.seed(3)
    b.run(3)
    close(a.expF, b.expF, rtol=1e-14)
    assert len(a.all_performances["MSE"]) == 3 and len(a.all_exp_tau) == 3


def test_icm_trajectory_matches_reference(golden):
    import bnmtf_b200
    g = golden("toy_nmtf_icm")
    K, L = int(g["K"]), int(g["L"])
    m = bnmtf_b200.nmtf_icm(g["R"], g["M"], K, L, priors3(g))
    m.initialise("exp", "exp")
    m.F, m.S, m.G = g["init_F"].copy(), g["init_S"].copy(), g["init_G"].copy()
    m.tau = (m.alpha_s() - 1.0) / m.beta_s()
    close(m.tau, g["init_tau"])
    m.run(int(g["its"]), minimum_TN=float(g["minimum_TN"]))
    close(m.all_performances["MSE"], g["trace_MSE"]), close(m.all_tau, g["trace_tau"])
    close(m.F, g["final_F"]), close(m.S, g["final_S"]), close(m.G, g["final_G"])
    close(m.quality("loglikelihood"), g["loglik"]), close(m.quality("AIC"), g["AIC"]), close(m.quality("BIC"), g["BIC"])


def test_gibbs_conditionals_match_reference(golden):
    import bnmtf_b200
    g = golden("toy_bnmtf_gibbs")
    K, L = int(g["K"]), int(g["L"])
    m = bnmtf_b200.bnmtf_gibbs_optimised(g["R"], g["M"], K, L, priors3(g))
    m.initialise("exp", "exp")
    m.F, m.S, m.G = g["init_F"].copy(), g["init_S"].copy(), g["init_G"].copy()
    m.tau = m.alpha_s() / m.beta_s()
    close(m.tau, g["init_tau"]), close(m.beta_s(), g["beta_s"])
    for k in range(K):
        t = m.tauF(k)
        close(t, g["tauF"][:, k]), close(m.muF(t, k), g["muF"][:, k])
    for l in range(L):
        t = m.tauG(l)
        close(t, g["tauG"][:, l]), close(m.muG(t, l), g["muG"][:, l])
    for k, l in ((0, 0), (2, 3), (K - 1, L - 1)):
        t = m.tauS(k, l)
        close(t, g["tauS"][k, l]), close(m.muS(t, k, l), g["muS"][k, l])


def test_gibbs_chain_matches_reference_in_distribution(golden):
    import bnmtf_b200
    g = golden("toy_bnmtf_gibbs")
    K, L = int(g["K"]), int(g["L"])
    m = bnmtf_b200.bnmtf_gibbs_optimised(g["R"], g["M"], K, L, priors3(g), seed=5)
    m.initialise("exp", "exp")
    m.F, m.S, m.G = g["init_F"].copy(), g["init_S"].copy(), g["init_G"].copy()
    m.tau = m.alpha_s() / m.beta_s()
    its, burn, thin = int(g["chain_its"]), int(g["chain_burn_in"]), int(g["chain_thinning"])
    all_F, all_S, all_G, all_tau = m.run(its)
    assert all_F.shape == (its, 100, K) and all_S.shape == (its, K, L) and (all_S >= 0).all()
    ref_mse, ref_tau = g["chains_MSE_tau"][:, 0], g["chains_MSE_tau"][:, 1]
    mse = m.quality("MSE", burn, thin)
    assert abs(mse - ref_mse.mean()) <= max(5 * ref_mse.std(), 0.10 * ref_mse.mean())
    exp_tau = m.approx_expectation(burn, thin)[3]
    assert abs(exp_tau - ref_tau.mean()) <= max(5 * ref_tau.std(), 0.10 * ref_tau.mean())


@pytest.mark.parametrize("cls_name,its", [("bnmtf_vb_optimised", 3), ("nmtf_icm", 3), ("bnmtf_gibbs_optimised", 2)])
def test_large_matrices_run_the_tcgen05_statistics(monkeypatch, cls_name, its):
    """From 2^22 entries the tri-factor engine takes its two statistics passes from the two-factor model's tcgen05 kernels
    (fixed-point int8 GEMMs; the statistics of R ~ F S G^T w.r.t. G and w.r.t. F are two-factor statistics:
    bnmtf_vb_optimised.py:245-280, bnmtf_gibbs_optimised.py:195-211).  Same trajectories as with the fp64 kernels."""
    import bnmtf_b200
    rng = np.random.RandomState(3)
    I, J, K, L = 1500, 3000, 6, 5
    R = np.abs(rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (K, L)) @ rng.exponential(1.0, (J, L)).T + rng.normal(size=(I, J))) + 0.5
    M = (rng.rand(I, J) >= 0.2).astype(float)
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}
    cls = getattr(bnmtf_b200, cls_name)
    out = {}
    for impl in ("umma", "dmma"):
        monkeypatch.setenv("BNMTF_NMTF_STATS", impl)
        np.random.seed(11), random.seed(11)
        m = cls(R, M, K, L, pri, seed=5)
        m.initialise("random", "random")
        if cls_name == "nmtf_icm":
            m.run(its, minimum_TN=0.1)
        else:
            m.run(its)
        assert m._engine().stats_impl == impl
        assert m._engine().metrics_mode == ("stats" if impl == "umma" else "direct")
        out[impl] = m
    monkeypatch.delenv("BNMTF_NMTF_STATS")
    a, b = out["umma"], out["dmma"]
    if cls_name == "bnmtf_vb_optimised":
        for name in ("expF", "expS", "expG", "muF", "muS", "muG"):
            close(getattr(a, name), getattr(b, name), what=name)
        close(a.exptau, b.exptau)
    else:
        # (Gibbs: the same Philox streams, so the same draws while the conditionals agree to rounding)
        for name in ("F", "S", "G"):
            close(getattr(a, name), getattr(b, name), rtol=1e-8 if cls_name == "bnmtf_gibbs_optimised" else 1e-9, what=name)
        close(a.tau, b.tau, rtol=1e-8)
    close(a.all_performances["MSE"], b.all_performances["MSE"], rtol=1e-8 if cls_name == "bnmtf_gibbs_optimised" else 1e-9)
    # the automatic choice at this size
    m = cls(R, M, K, L, pri, seed=5)
    np.random.seed(11), random.seed(11)
    m.initialise("random", "random")
    m.run(1) if cls_name != "nmtf_icm" else m.run(1, minimum_TN=0.1)
    assert m._engine().stats_impl == "umma"


@pytest.mark.parametrize("missing", [0.2, 0.8])
@pytest.mark.parametrize("cls_name", ["bnmtf_vb_optimised", "bnmtf_gibbs_optimised"])
def test_metrics_from_the_column_statistics(monkeypatch, cls_name, missing):
    """With the tcgen05 statistics the training metrics of a sweep come from the column statistics of the G phase
    (csrc/nmtf.cu::k_nmtf_mstat) instead of a pass over R (predict_while_running / compute_MSE / compute_R2 / compute_Rp, bnmtf_gibbs_optimised.py:234-258): same
    traces as the direct pass, for both mask polarities, and the direct pass takes over when the guard trips."""
    import bnmtf_b200
    rng = np.random.RandomState(8)
    I, J, K, L = 1200, 3600, 4, 7
    R = np.abs(rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (K, L)) @ rng.exponential(1.0, (J, L)).T + rng.normal(size=(I, J))) + 0.5
    M = (rng.rand(I, J) >= missing).astype(float)
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}
    cls = getattr(bnmtf_b200, cls_name)
    out = {}
    for mode in ("stats", "direct", "guard"):
        monkeypatch.setenv("BNMTF_METRICS", "direct" if mode == "direct" else "stats")
        np.random.seed(2), random.seed(2)
        m = cls(R, M, K, L, pri, seed=9)
        m.initialise("random", "random")
        eng = m._engine()
        assert eng.stats_impl == "umma" and eng.polarity == (0 if missing < 0.5 else 1)
        if mode == "guard":
            eng.guard = 1e9                    # every sweep: the statistics-based sums are rejected on the device
        m.run(3)
        assert int(eng.flag.item()) == (1 if mode == "guard" else 0)
        out[mode] = m
    a, b, c = out["stats"], out["direct"], out["guard"]
    # the sums behind the metrics carry the 1e-11..1e-10 relative error of the fixed-point statistics, and MSE / R^2 / Rp are
    # differences of such sums: 1e-8 on the metrics; the fallback run repeats the direct run up to the rounding of one
    # more column in the Gram kernel's launch plan
    for key in ("MSE", "R^2", "Rp"):
        close(a.all_performances[key], b.all_performances[key], rtol=1e-8, what=key)
        close(c.all_performances[key], b.all_performances[key], rtol=1e-11, what=key + " (fallback)")
    close(a.exptau if cls_name == "bnmtf_vb_optimised" else a.tau, b.exptau if cls_name == "bnmtf_vb_optimised" else b.tau, rtol=1e-8)


@pytest.mark.parametrize("K,L", [(3, 4), (10, 10), (5, 9), (2, 30), (12, 12), (20, 20), (3, 40), (32, 32)])
@pytest.mark.parametrize("vb", [0, 1])
@pytest.mark.parametrize("polarity", [0, 1])
def test_s_phase_reduction_is_the_einsum(K, L, vb, polarity):
    """bnmtf_nmtf_sq_f64 (csrc/nmtf.cu: the register-tiled product k_nmtf_sq_tiled -- one pass over the rows up to K = L = 10,
    several for the larger shapes -- and k_nmtf_sq_partial for shapes it does not take, here VB at K = L = 32: more than 64
    passes) against the defining sums of the S update (bnmtf_vb_optimised.py:256-266,
    bnmtf_gibbs_optimised.py:201-205) written with numpy on the same row statistics."""
    import torch
    from bnmtf_b200 import _lib
    from bnmtf_b200.engine import _ptr, _stream, kp_for, gram_len
    rng = np.random.RandomState(100 * K + L)
    rows, D = 777, K * L
    nt, KP, GL = kp_for(L) // 8, kp_for(L), gram_len(L)
    Gm = rng.rand(rows, 40, L)
    GG = np.einsum("ijl,ijm->ilm", Gm, Gm)                       # per-row Gram matrices (symmetric positive)
    sv, rg = rng.rand(rows, L) * 3.0, rng.randn(rows, L) * 5.0
    F, vF = rng.exponential(1.0, (rows, K)), rng.rand(rows, K) * 0.3
    full_GG, full_sv = GG.max(axis=0) * 1.5 + 1.0, sv.max(axis=0) * 1.5 + 1.0

    def pack(mats):                                               # (n, L, L) -> packed upper-triangular 8x8 tiles
        out = np.zeros((mats.shape[0], GL))
        P = np.zeros((mats.shape[0], KP, KP))
        P[:, :L, :L] = mats
        p = 0
        for ta in range(nt):
            for tb in range(ta, nt):
                out[:, p * 64:(p + 1) * 64] = P[:, 8 * ta:8 * ta + 8, 8 * tb:8 * tb + 8].reshape(-1, 64)
                p += 1
        return out
    # what the statistics kernels hold: sums over the observed set (polarity 1) or over the missing set (polarity 0)
    held_GG = GG if polarity else full_GG[None] - GG
    held_sv = sv if polarity else full_sv[None] - sv
    dev = torch.device("cuda:0")
    up = lambda x: torch.tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)
    svp = np.zeros((rows, KP)); svp[:, :L] = held_sv
    rgp = np.zeros((rows, KP)); rgp[:, :L] = rg
    fullrec = np.zeros(GL + KP); fullrec[:GL] = pack(full_GG[None])[0]; fullrec[GL:GL + L] = full_sv
    Go, SVo, RXo, Gfull, Fd, vFd = up(pack(held_GG)), up(svp), up(rgp), up(fullrec), up(F), up(vF)
    nparts, ln = 37, D * D + 2 * D
    scratch = int(_lib.call("bnmtf_nmtf_sq_scratch_len", K, L, vb))
    assert scratch >= ln and int(_lib.call("bnmtf_nmtf_sq_parts", rows, K, L, vb)) >= 1
    part, out = torch.zeros(nparts * scratch, dtype=torch.float64, device=dev), torch.zeros(ln, dtype=torch.float64, device=dev)
    _lib.call("bnmtf_nmtf_sq_f64", rows, K, L, polarity, vb, _ptr(RXo), _ptr(Go), _ptr(SVo) if vb else 0, _ptr(Gfull), _ptr(Fd),
              _ptr(vFd) if vb else 0, _ptr(part), nparts, _ptr(out), _stream())
    got = out.cpu().numpy()
    A = (F[:, :, None] * F[:, None, :]).reshape(rows, K * K)
    H = (A.T @ GG.reshape(rows, L * L)).reshape(K, K, L, L).transpose(0, 2, 1, 3).copy()      # [k, l, k', l']
    prec = np.einsum("ik,il->kl", F * F, np.einsum("ill->il", GG))
    if vb:
        covG = np.einsum("ik,im,il->kml", F, F, sv)
        covF = np.einsum("ik,iln->kln", vF, GG)
        for k in range(K):
            for l in range(L):
                for k2 in range(K):
                    if k2 != k:
                        H[k, l, k2, l] += covG[k, k2, l]
                for l2 in range(L):
                    if l2 != l:
                        H[k, l, k, l2] += covF[k, l, l2]
        prec = np.einsum("ik,il->kl", vF + F * F, np.einsum("ill->il", GG) + sv)
    rhs = np.einsum("ik,il->kl", F, rg)
    Hg = got[:D * D].reshape(D, D)
    off = ~np.eye(D, dtype=bool)                                  # (the diagonal of H is never read by the coordinate updates)
    close(Hg[off], H.reshape(D, D)[off], rtol=1e-11, what="H")
    close(got[D * D:D * D + D], prec.reshape(-1), rtol=1e-11, what="precision")
    close(got[D * D + D:], rhs.reshape(-1), rtol=1e-11, what="right-hand side")
