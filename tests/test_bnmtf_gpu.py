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


@pytest.mark.parametrize("name", ["toy_bnmtf_vb", "gdsc_bnmtf_vb"])
def test_vb_every_sweep_matches_oracle_from_the_same_state(golden, name):
    """Per-sweep parity without error accumulation: before every sweep the device state is set to the oracle's
    state (which itself follows the reference's golden trajectory to ~1e-9), then both do one sweep with the
    reference's shuffled orders and all variational parameters are compared.

    Tolerance 1e-7, not 1e-9: the reference's TN variance sigma^2 (1 - lambda (lambda - x)) cancels catastrophically
    for strongly truncated entries (x = -mu sqrt(tau) between ~10 and the 30-sigma switch; most of S in these runs):
    a 1-ulp difference between CUDA's and SciPy's erfc/exp becomes up to ~x^4 * 2e-13 relative in varS, and varS
    enters tauF/tauG directly (measured: 7e-9 in tauF after the first S phase on GDSC, tools/gpu_debug2.py).  The
    same spread exists between two SciPy builds, so it is the formula's conditioning, not the kernels'.  The
    single-update tests above (fixed inputs, no variance recomputation in between) hold 1e-10."""
    tol = 2e-7
    from oracle import bnmtf_oracle as orc
    g = golden(name)
    K, L = int(g["K"]), int(g["L"])
    m = vb_from_golden(g)
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, priors3(g), mode="vb")
    o.init_vb(g["init_muF"], g["init_muS"], g["init_muG"], {"F": g["init_tauF"], "S": g["init_tauS"], "G": g["init_tauG"]})
    for it in range(min(12, int(g["its"]))):
        _set_vb_state(m, o)
        eng = m._push()
        eng.alloc_trace(1)
        oS = [tuple(int(v) for v in x) for x in g["order_S"][it]]
        order = {"S": oS, "F": [int(x) for x in g["order_F"][it]], "G": [int(x) for x in g["order_G"][it]]}
        perf = o.sweep(order=order)
        eng.sweep(order={"S": [k * L + l for k, l in oS], "F": order["F"], "G": order["G"]})
        tr = eng.trace.cpu().numpy()[0]
        m._pull(eng)
        for k in "FSG":
            close(getattr(m, "exp" + k), getattr(o, k), rtol=tol, what="exp%s it %d" % (k, it))
            # the variance itself: up to ~1.5e-7 for the entries just inside the 30-sigma switch (measured), so 1e-6
            close(getattr(m, "var" + k), getattr(o, "var" + k), rtol=1e-6, what="var%s it %d" % (k, it))
            close(getattr(m, "tau" + k), getattr(o, "tau" + k), rtol=tol, what="tau%s it %d" % (k, it))
            close(getattr(m, "mu" + k), getattr(o, "mu" + k), rtol=tol, what="mu%s it %d" % (k, it))
        close(tr[1], perf["MSE"], rtol=tol), close(tr[0], o.exptau, rtol=tol), close(tr[6], o.exp_square_diff(), rtol=tol)
        if np.isfinite(o.elbo()):
            close(tr[4], o.elbo(), rtol=tol, what="elbo it %d" % it)


@pytest.mark.parametrize("name,rtol", [("toy_bnmtf_vb", 1e-6), ("gdsc_bnmtf_vb", 1e-9)])
def test_vb_trajectory_matches_reference(golden, name, rtol):
    """Free-running trajectory against the reference's golden run, replaying its python-random shuffles.
    VB-NMTF on the toy data amplifies rounding differences (two fp64 CPU evaluations -- reference vs oracle -- drift
    from 2e-12 after one sweep to 1e-9 in the traces / 2e-8 in the factors after 30 sweeps, tests/test_oracle_golden.py);
    the device path sums in yet another order, hence 1e-6 there.  GDSC stays within 1e-9."""
    g = golden(name)
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
    close(tr[:, 1], g["trace_MSE"], rtol=rtol, what="MSE trace")
    close(tr[:, 0], g["trace_exptau"], rtol=rtol, what="exptau trace")
    ok = np.isfinite(g["trace_elbo"])
    close(tr[ok, 4], g["trace_elbo"][ok], rtol=rtol, what="ELBO trace")
    m._pull(eng)
    for k in "FSG":
        close(getattr(m, "exp" + k), g["final_exp" + k], rtol=max(rtol, 1e-8) * 100, what="exp" + k)


@pytest.mark.parametrize("K,L,init_FG", [(5, 4, "kmeans"), (4, 6, "kmeans"), (3, 7, "random"), (6, 2, "random")])
def test_vb_rectangular_core_matches_oracle(golden, K, L, init_FG):
    """K != L (the golden trajectories are all K = L = 5): host initialisation with the reference's draw order (numpy
    exponentials for S, python-random K-means for F and G), then 8 free-running sweeps with the reference's shuffles,
    against the oracle from the same start.  Measured 4e-10 on the MSE trace, 2e-8 on the factors (tools/gpu_debug3.py)."""
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
