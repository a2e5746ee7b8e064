"""Dynamic range of the fixed-point statistics kernels (csrc/rx_umma.cu, csrc/gram_umma.cu).

Both kernels use ONE power-of-two scale per row of R and per factor / product column, 48 bits below it.  All other
test data (Exp(1) x Exp(1) + N(0,1), GDSC with R in [1, 34]) is benign for that.  Here: rows of R with outliers 1e6 - 1e9
times their typical entry, and factor columns spanning 1e-8 ... 1e3 -- either the result still holds 1e-9 against
numpy longdouble, or a guard has switched to the fp64 kernels (and then it holds too):
  * static guard (dataset pack): Dataset.wide -> the engine keeps the fp64 R.X kernel, with a warning;
  * dynamic guard (every phase): bnmtf_range_guard_f64 raises a device flag, the gated fp64 kernels recompute the
    statistics, BNMFEngine.range_trips counts the phases in which that happened.
Reference formulas: bnmf_vb_optimised.py:189-195 (per row: tauU, muU from the masked sums).
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LD = np.longdouble
PRI = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}


def base(I, J, K, seed):
    rng = np.random.RandomState(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J))
    M = (rng.rand(I, J) >= 0.2).astype(float)
    return rng, R, M


def vb_row_params(R, M, expU, expV, varV, lam, tau, k):
    """tauU[:, k], muU[:, k] of bnmf_vb_optimised.update_U(k) in longdouble."""
    Rl, Ml, U, V = R.astype(LD), M.astype(LD), expU.astype(LD), expV.astype(LD)
    t = LD(tau) * (Ml @ (V[:, k] ** 2 + varV[:, k].astype(LD)))
    resid = Rl - U @ V.T + np.outer(U[:, k], V[:, k])
    mu = (-LD(lam) + LD(tau) * ((Ml * resid) @ V[:, k])) / t
    return t.astype(float), mu.astype(float)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-2 * float(np.abs(b).max()))))


def model_with_state(R, M, K, expU, expV, seed=0):
    import bnmtf_b200
    m = bnmtf_b200.bnmf_vb_optimised(R, M, K, PRI, seed=seed)
    m.initialise("exp")
    rng = np.random.RandomState(99)
    m.expU, m.expV = expU.copy(), expV.copy()
    m.varU, m.varV = 0.01 * expU ** 2 * rng.rand(*expU.shape), 0.01 * expV ** 2 * rng.rand(*expV.shape)
    m.exptau = 0.7
    return m


@pytest.mark.parametrize("factor", [1e4, 1e9])
def test_row_outliers(factor):
    """Isolated outliers in three rows.  1e4 x the typical entry: no flag, tensor-core kernels, 1e-9 holds.  1e9 x: the
    static guard flags the dataset (typical entries would keep 18 of 48 bits), fp64 kernels, 1e-9 holds."""
    I, J, K = 300, 2000, 8
    rng, R, M = base(I, J, K, 1)
    for i in (3, 150, 299):
        j = int(rng.randint(J))
        R[i, j], M[i, j] = factor * 5.0, 1.0
    expU, expV = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = model_with_state(R, M, K, expU, expV)          # the engine is built by initialise()
        eng = m._engine()
    if factor > 1e6:
        assert eng.wide_dataset and eng.rx == "dmma" and eng.gram == "dmma" and any("outliers" in str(x.message) for x in w)
    else:
        assert not eng.wide_dataset and eng.rx == "umma" and eng.gram == "umma"
    for k in (0, K - 1):
        m.update_U(k)
        t, mu = vb_row_params(R, M, expU, expV, m.varV, 0.1, 0.7, k)
        assert rel(m.tauU[:, k], t) < 1e-11 and rel(m.muU[:, k], mu) < 1e-9, (k, rel(m.muU[:, k], mu))
        for i in (3, 150, 299):                                                    # the outlier rows on their own scale
            assert abs(m.muU[i, k] / mu[i] - 1.0) < 1e-9 and abs(m.tauU[i, k] / t[i] - 1.0) < 1e-11
    if factor > 1e6:
        # the same update with the guard off: the fixed-point kernels are scale-covariant (the outlier dominates the sums
        # by the factor by which it coarsens the quantum), so even here they hold -- the guard is a backstop
        import os
        os.environ["BNMTF_RANGE_GUARD"] = "0"
        try:
            m2 = model_with_state(R, M, K, expU, expV)
            assert m2._engine().rx == "umma"
            m2.update_U(0)
            t, mu = vb_row_params(R, M, expU, expV, m2.varV, 0.1, 0.7, 0)
            print("1e9 outliers on the tensor-core kernels (guard off): mu off by %.1e" % rel(m2.muU[:, 0], mu))
            assert rel(m2.muU[:, 0], mu) < 1e-9
        finally:
            del os.environ["BNMTF_RANGE_GUARD"]


def test_a_heavy_column():
    """One column of R on a 1e8 x larger scale (another unit, say) and a state that FITS it: V_j is 1e8 x larger too and
    its products set the fixed-point scale of every product column.  Every row has that outlier, so the static guard
    flags the dataset and the fp64 kernels are used: 1e-9 holds."""
    I, J, K = 300, 2000, 8
    rng = np.random.RandomState(4)
    U0, V0 = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    V0[77] *= 1e8
    R = U0 @ V0.T + rng.normal(size=(I, J))
    M = (rng.rand(I, J) >= 0.2).astype(float)
    M[:, 77] = 1.0
    expU, expV = U0 * (1.0 + 1e-3 * rng.rand(I, K)), V0.copy()
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        m = model_with_state(R, M, K, expU, expV)
    assert m._engine().wide_dataset
    m.update_U(0)
    t, mu = vb_row_params(R, M, expU, expV, m.varV, 0.1, 0.7, 0)
    assert rel(m.tauU[:, 0], t) < 1e-11 and rel(m.muU[:, 0], mu) < 1e-9, rel(m.muU[:, 0], mu)


def test_moderate_row_outliers_stay_on_the_tensor_cores():
    """An outlier 1000 x the typical entry costs 10 of 48 bits: no flag, and 1e-9 still holds with room to spare."""
    I, J, K = 300, 2000, 8
    rng, R, M = base(I, J, K, 2)
    R[7, 11], M[7, 11] = 5000.0, 1.0
    expU, expV = rng.exponential(1.0, (I, K)), rng.exponential(1.0, (J, K))
    m = model_with_state(R, M, K, expU, expV)
    eng = m._engine()
    assert not eng.wide_dataset and eng.rx == "umma"
    m.update_U(2)
    t, mu = vb_row_params(R, M, expU, expV, m.varV, 0.1, 0.7, 2)
    assert rel(m.tauU[:, 2], t) < 1e-11 and rel(m.muU[:, 2], mu) < 1e-10
    assert int(eng.range_trips.item()) == 0


def test_wide_factor_columns():
    """Factor columns spanning 1e-8 ... 1e3 (log-uniform).  (a) every row observes entries of every magnitude: the sums
    are dominated by the large entries, which keep all their bits -- no flag, 1e-9 holds.  (b) some rows observe ONLY
    columns whose factor entries are tiny: their statistics would have few significant bits -- the guard trips, the fp64
    kernels recompute the phase, 1e-9 holds."""
    I, J, K = 256, 4096, 6
    rng, R, M = base(I, J, K, 3)
    expU = rng.exponential(1.0, (I, K))
    expV = 10.0 ** rng.uniform(-8.0, 3.0, (J, K))
    R = expU @ expV.T + rng.normal(size=(I, J))
    m = model_with_state(R, M, K, expU, expV)
    eng = m._engine()
    m.update_U(1)
    t, mu = vb_row_params(R, M, expU, expV, m.varV, 0.1, 0.7, 1)
    assert int(eng.range_trips.item()) == 0
    assert rel(m.tauU[:, 1], t) < 1e-11 and rel(m.muU[:, 1], mu) < 1e-9, rel(m.muU[:, 1], mu)
    # (b): rows 0..9 observe only the columns where column 1 of the factor is below 1e-5
    tiny = expV[:, 1] < 1e-5
    M2 = M.copy()
    M2[:10] = 0.0
    M2[:10, tiny] = 1.0
    m2 = model_with_state(R, M2, K, expU, expV)
    eng2 = m2._engine()
    m2.update_U(1)
    t, mu = vb_row_params(R, M2, expU, expV, m2.varV, 0.1, 0.7, 1)
    assert int(eng2.range_trips.item()) >= 1, "rows that only see tiny factor entries must trip the guard"
    assert rel(m2.tauU[:, 1], t) < 1e-11 and rel(m2.muU[:, 1], mu) < 1e-9, rel(m2.muU[:, 1], mu)
    assert rel(m2.tauU[:10, 1], t[:10]) < 1e-11 and rel(m2.muU[:10, 1], mu[:10]) < 1e-9      # the rows in question, on their own scale
    assert rel(m2.tauU[10:, 1], t[10:]) < 1e-11 and rel(m2.muU[10:, 1], mu[10:]) < 1e-9
    # and a whole sweep on such data runs through the guard as well (no NaN, MSE finite)
    m2.run(2)
    assert np.isfinite(m2.all_performances["MSE"]).all()
