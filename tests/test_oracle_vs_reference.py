"""The CPU oracle against the LIVE reference (build container only: /root/reference imported through
oracle/ref_shim.py) on fresh seeded inputs the golden fixtures do not contain -- rectangular K != L, other sizes and
missing fractions.  Complements tests/test_oracle_golden.py (committed fixtures, runs anywhere)."""
import contextlib
import io
import os
import random
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not os.path.isdir("/root/reference/code/models"), reason="reference tree not present")]

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import bnmtf_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_shim
    return ref_shim.load()


def make(I, J, K, frac, seed, L=None):
    rng = np.random.RandomState(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J)) + 3.0
    M = (rng.rand(I, J) >= frac).astype(float)
    M[np.arange(I), rng.randint(0, J, I)] = 1.0
    M[rng.randint(0, I, J), np.arange(J)] = 1.0
    return np.abs(R), M


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-9 * max(1.0, float(np.abs(b).max())))))


@pytest.mark.parametrize("I,J,K,frac", [(23, 17, 4, 0.3), (40, 55, 7, 0.6)])
def test_bnmf_vb_and_icm_sweeps(ref, I, J, K, frac):
    R, M = make(I, J, K, frac, 1)
    pri = {"alpha": 2.0, "beta": 0.5, "lambdaU": 0.3, "lambdaV": 0.7}
    np.random.seed(4)
    m = ref.bnmf_vb_optimised(R, M, K, pri)
    with quiet():
        m.initialise("random")
    o = orc.OracleBNMF(R, M, K, pri, mode="vb")
    o.init_vb(m.muU.copy(), m.muV.copy())
    for _ in range(6):
        with quiet():
            m.run(1)
        perf = o.sweep()
        assert rel(o.U, m.expU) < 1e-10 and rel(o.varV, m.varV) < 1e-9 and rel(o.exptau, m.exptau) < 1e-11
        assert rel(perf["MSE"], m.all_performances["MSE"][-1]) < 1e-11
        assert rel(o.elbo(), m.elbo()) < 1e-10
    for metric in ("loglikelihood", "AIC", "BIC", "MSE"):
        assert rel(o.quality(metric), m.quality(metric)) < 1e-10
    np.random.seed(5)
    c = ref.nmf_icm(R, M, K, pri)
    with quiet():
        c.initialise("random")
    oc = orc.OracleBNMF(R, M, K, pri, mode="icm")
    oc.set_state(c.U.copy(), c.V.copy(), tau=c.tau)
    for _ in range(6):
        with quiet():
            c.run(1, minimum_TN=0.05)
        oc.sweep(minimum_TN=0.05)
        assert rel(oc.U, c.U) < 1e-10 and rel(oc.V, c.V) < 1e-10 and rel(oc.tau, c.tau) < 1e-11


@pytest.mark.parametrize("K,L", [(3, 5), (6, 2)])
def test_bnmtf_vb_with_the_reference_shuffles(ref, K, L):
    R, M = make(31, 26, max(K, L), 0.35, 2)
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.2, "lambdaS": 0.5, "lambdaG": 0.4}
    np.random.seed(6), random.seed(6)
    m = ref.bnmtf_vb_optimised(R, M, K, L, pri)
    with quiet():
        m.initialise("random", "random")
    o = orc.OracleBNMTF(R, M, K, L, pri, mode="vb")
    o.init_vb(m.muF.copy(), m.muS.copy(), m.muG.copy())
    for it in range(5):
        random.seed(100 + it)
        with quiet():
            m.run(1)
        random.seed(100 + it)
        o.sweep(order=orc.OracleBNMTF.shuffled_order(K, L))
        # VB-NMTF means cancel O(scale) terms: 1e-8 between two evaluation orders
        assert rel(o.F, m.expF) < 1e-8 and rel(o.S, m.expS) < 1e-8 and rel(o.G, m.expG) < 1e-8
        assert rel(o.exptau, m.exptau) < 1e-9
    assert rel(o.quality("MSE"), m.quality("MSE")) < 1e-9


def test_np_models_and_gibbs_conditionals(ref):
    R, M = make(19, 14, 3, 0.25, 3)
    np.random.seed(7)
    n = ref.NMF(R, M, 3)
    with quiet():
        n.initialise("random", expo_prior=0.5)
    on = orc.OracleBNMF(R, M, 3, mode="np")
    on.set_state(n.U.copy(), n.V.copy())
    for _ in range(5):
        with quiet():
            n.run(1)
        on.sweep()
        assert rel(on.U, n.U) < 1e-11 and rel(on.V, n.V) < 1e-11
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
    np.random.seed(8)
    g = ref.bnmf_gibbs_optimised(R, M, 3, pri)
    with quiet():
        g.initialise("random")
    og = orc.OracleBNMF(R, M, 3, pri, mode="gibbs")
    og.set_state(g.U.copy(), g.V.copy(), tau=g.tau)
    for k in range(3):
        tU, mU = og.column_params(k, "U")
        assert rel(tU, g.tauU(k)) < 1e-12 and rel(mU, g.muU(g.tauU(k), k)) < 1e-10
        tV, mV = og.column_params(k, "V")
        assert rel(tV, g.tauV(k)) < 1e-12 and rel(mV, g.muV(g.tauV(k), k)) < 1e-10
    assert rel(og.beta_s(), g.beta_s()) < 1e-12
