"""Single-kernel sweeps for small two-factor problems (csrc/small.cu, bnmtf_small_sweeps_f64) against the multi-kernel path
of the same engine (itself pinned against the reference's goldens): same run() of bnmf_gibbs_optimised.py:121-157,
bnmf_vb_optimised.py:121-153, nmf_icm.py:114-149, same Philox streams -- so the same traces, factors and draws up to the
rounding of the row statistics (plain fp64 sums here, fixed-point tensor-core sums there)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PRI = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}


def problem(I, J, K, seed):
    rng = np.random.RandomState(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J))
    M = (rng.rand(I, J) >= 0.25).astype(float)
    M[np.arange(I), rng.randint(0, J, I)] = 1.0
    M[rng.randint(0, I, J), np.arange(J)] = 1.0
    return R, M


def run(cls_name, R, M, K, its, small, monkeypatch, **kw):
    import bnmtf_b200
    monkeypatch.setenv("BNMTF_SMALL", "1" if small else "0")
    np.random.seed(4)
    m = getattr(bnmtf_b200, cls_name)(R, M, K, PRI, seed=9)
    m.initialise("random")
    out = m.run(its, **kw) if cls_name != "nmf_icm" else m.run(its, minimum_TN=0.1)
    assert bool(m._engine().small_cluster()) == small
    return m, out


# K = 10: three column groups per warp; 7: four groups and idle lanes; 16: two groups; 1; ragged row / column counts
@pytest.mark.parametrize("I,J,K", [(100, 80, 10), (130, 45, 7), (257, 300, 16), (40, 33, 1), (622, 138, 10), (33, 700, 5)])
@pytest.mark.parametrize("cls_name", ["bnmf_vb_optimised", "nmf_icm", "bnmf_gibbs_optimised"])
def test_single_kernel_sweeps_match_the_multi_kernel_path(monkeypatch, cls_name, I, J, K):
    R, M = problem(I, J, K, I + J + K)
    its = 12
    a, _ = run(cls_name, R, M, K, its, True, monkeypatch)
    b, _ = run(cls_name, R, M, K, its, False, monkeypatch)
    tol = dict(rtol=1e-8, atol=1e-10)      # (12 Gauss-Seidel sweeps amplify the 1e-15 differences of the statistics)
    for metric in ("MSE", "R^2", "Rp"):
        np.testing.assert_allclose(a.all_performances[metric], b.all_performances[metric], **tol)
    if cls_name == "bnmf_vb_optimised":
        for name in ("expU", "expV", "varU", "varV", "muU", "tauU"):
            np.testing.assert_allclose(getattr(a, name), getattr(b, name), rtol=1e-7, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(a.all_exp_tau, b.all_exp_tau, **tol)
        np.testing.assert_allclose(a.quality("ELBO"), b.quality("ELBO"), rtol=1e-9)
    else:
        np.testing.assert_allclose(a.U, b.U, rtol=1e-6, atol=1e-9)       # (single entries of an over-parameterised ICM fit move by 1e-7)
        np.testing.assert_allclose(a.V, b.V, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(a.all_tau, b.all_tau, **tol)
    assert len(a.all_times) == its and all(t1 > t0 for t0, t1 in zip(a.all_times, a.all_times[1:])) and a.all_times[0] > 0


def test_gibbs_draws_and_running_sums_come_from_the_kernel(monkeypatch):
    R, M = problem(120, 90, 6, 3)
    a, out_a = run("bnmf_gibbs_optimised", R, M, 6, 25, True, monkeypatch)
    b, out_b = run("bnmf_gibbs_optimised", R, M, 6, 25, False, monkeypatch)
    all_U, all_V, all_tau = out_a
    ref_U, ref_V, ref_tau = out_b
    assert np.asarray(all_U).shape == (25, 120, 6) and np.asarray(all_V).shape == (25, 90, 6)
    np.testing.assert_allclose(np.asarray(all_U), np.asarray(ref_U), rtol=1e-6, atol=1e-9)      # the same chain, draw for draw
    np.testing.assert_allclose(np.asarray(all_tau), np.asarray(ref_tau), rtol=1e-8)
    np.testing.assert_array_equal(np.asarray(all_U)[-1], a.U)
    eU, eV, etau = a.approx_expectation(5, 2)
    np.testing.assert_allclose(eU, np.asarray(all_U)[5::2].mean(axis=0), rtol=1e-13)
    s, _ = run("bnmf_gibbs_optimised", R, M, 6, 25, True, monkeypatch, summary=(5, 2))
    sU, sV, stau = s.approx_expectation(5, 2)
    np.testing.assert_allclose(sU, eU, rtol=1e-14), np.testing.assert_allclose(sV, eV, rtol=1e-14)


def test_large_or_wide_problems_keep_the_multi_kernel_path():
    from bnmtf_b200 import _lib
    assert _lib.call("bnmtf_small_cluster_size", 100, 80, 10, 1) == 4
    assert _lib.call("bnmtf_small_cluster_size", 622, 138, 10, 1) == 16
    assert _lib.call("bnmtf_small_cluster_size", 622, 138, 17, 0) == 0          # K > 16
    assert _lib.call("bnmtf_small_cluster_size", 5000, 300, 10, 0) == 0         # more than 256 rows per CTA of the largest cluster
    assert _lib.call("bnmtf_small_cluster_size", 2000, 2000, 10, 1) == 0        # does not fit shared memory


# ---- tri-factorisation ------------------------------------------------------------------------------------------
PRI3 = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}


def run3(cls_name, R, M, K, L, its, small, monkeypatch):
    import random
    import bnmtf_b200
    monkeypatch.setenv("BNMTF_SMALL", "1" if small else "0")
    np.random.seed(4), random.seed(4)
    m = getattr(bnmtf_b200, cls_name)(R, M, K, L, PRI3, seed=9)
    m.initialise("random", "random")
    out = m.run(its) if cls_name != "nmtf_icm" else m.run(its, minimum_TN=0.1)
    from bnmtf_b200 import _lib
    fits = _lib.call("bnmtf_small_tri_cluster_size", R.shape[0], R.shape[1], K, L, int(cls_name == "bnmtf_vb_optimised")) > 0
    assert bool(m._engine().small_cluster()) == (small and fits)
    return m, out


@pytest.mark.parametrize("I,J,K,L", [(100, 80, 5, 5), (130, 45, 7, 3), (257, 300, 4, 9), (622, 138, 7, 7), (40, 33, 2, 3)])
@pytest.mark.parametrize("cls_name", ["bnmtf_vb_optimised", "nmtf_icm", "bnmtf_gibbs_optimised"])
def test_tri_factor_single_kernel_sweeps_match_the_multi_kernel_path(monkeypatch, cls_name, I, J, K, L):
    """csrc/small.cu::k_small_tri against bnmtf.py::BNMTFEngine's per-phase kernels: bnmtf_vb_optimised.run (shuffled S, F, G
    orders, covariance terms, ELBO), nmtf_icm.run, bnmtf_gibbs_optimised.run (same Philox streams: the same chain)."""
    R, M = problem(I, J, max(K, L), I + J + K + L)
    R = np.abs(R) + 0.5
    its = 8
    a, out_a = run3(cls_name, R, M, K, L, its, True, monkeypatch)
    b, out_b = run3(cls_name, R, M, K, L, its, False, monkeypatch)
    # (VB-NMTF feeds the last bits of its truncated-normal moments back every sweep: tests/golden/vb_nmtf_sensitivity.json;
    # R^2 near zero is 1 - SSE/SST: an absolute floor)
    tol = dict(rtol=1e-6, atol=1e-8)
    for metric in ("MSE", "R^2", "Rp"):
        x, y = np.array(a.all_performances[metric]), np.array(b.all_performances[metric])
        if metric == "Rp":
            # (the first ICM sweep clamps every entry to minimum_TN: a constant prediction, Rp = 0/0 up to rounding)
            ok = np.isfinite(x) & np.isfinite(y) & (np.abs(y) > 1e-6)
            x, y = x[ok], y[ok]
        np.testing.assert_allclose(x, y, **tol)
    if cls_name == "bnmtf_vb_optimised":
        for name in ("expF", "expS", "expG", "varF", "varS", "varG", "muS", "tauS"):
            np.testing.assert_allclose(getattr(a, name), getattr(b, name), rtol=1e-6, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(a.all_exp_tau, b.all_exp_tau, **tol)
        np.testing.assert_allclose(a.all_elbo, b.all_elbo, rtol=1e-8)
    else:
        for name in ("F", "S", "G"):
            np.testing.assert_allclose(getattr(a, name), getattr(b, name), rtol=1e-6, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(a.all_tau, b.all_tau, **tol)
    if cls_name == "bnmtf_gibbs_optimised":
        for x, y in zip(out_a[:3], out_b[:3]):
            assert np.asarray(x).shape == np.asarray(y).shape
            np.testing.assert_allclose(np.asarray(x), np.asarray(y), rtol=1e-5, atol=1e-9)
    assert len(a.all_times) == its and all(t1 > t0 for t0, t1 in zip(a.all_times, a.all_times[1:]))


def test_tri_factor_eligibility():
    from bnmtf_b200 import _lib
    assert _lib.call("bnmtf_small_tri_cluster_size", 100, 80, 5, 5, 1) == 4
    assert _lib.call("bnmtf_small_tri_cluster_size", 622, 138, 5, 5, 1) == 16
    assert _lib.call("bnmtf_small_tri_cluster_size", 622, 138, 10, 10, 0) == 0      # K*L > 64: the per-phase kernels are as fast
    assert _lib.call("bnmtf_small_tri_cluster_size", 5000, 300, 5, 5, 0) == 0
