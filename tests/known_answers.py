"""Known-answer checks carried over from the reference's own unit tests, written once and applied
to any implementation of the model classes: the reference's classes (in the build container, which proves the numbers
below were transcribed correctly) and this package's GPU classes.

Sources (relative to /root/reference/tests/code): test_bnmf_gibbs_optimised.py:143-203 (alpha*, beta*, tauU, muU, tauV,
muV on the 5 x 3 mask with three holes), :239-360 (approx_expectation, predict, compute_statistics, quality);
test_bnmf_vb_optimised.py:160-216 (ELBO), :230-309 (exp_square_diff, update_tau, update_U/V, update_exp_*);
test_nmf_icm.py (same conditionals, tau = Gamma mode); test_nmf_np.py:89-167 (multiplicative updates).

The reference asserts most of these with `==` on doubles; a GPU reduction sums in another order, so every comparison
here is `rel <= 1e-12` (or the reference's own looser bound where it has one: 1e-5 for the TN moments quoted to six
digits).  That is the only relaxation."""
import math

import numpy as np

RTOL = 1e-12


def eq(got, want, tol=RTOL, what=""):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.all(np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want))), (what, got, want)


def five_by_three():
    I, J, K = 5, 3, 2
    R, M = np.ones((I, J)), np.ones((I, J))
    M[0, 0], M[2, 2], M[3, 1] = 0, 0, 0                      # |Omega| = 12
    lambdaU, lambdaV = 2 * np.ones((I, K)), 3 * np.ones((J, K))
    pri = {'alpha': 3, 'beta': 1, 'lambdaU': lambdaU, 'lambdaV': lambdaV}
    return I, J, K, R, M, lambdaU, lambdaV, pri


def check_conditionals(cls, icm=False):
    """Gibbs / ICM conditionals from the 'exp' start U = 1/2, V = 1/3 with tau = 3."""
    I, J, K, R, M, lambdaU, lambdaV, pri = five_by_three()
    m = cls(R, M, K, pri)
    m.initialise('exp')
    eq(m.U, np.full((I, K), 0.5)), eq(m.V, np.full((J, K), 1. / 3.))
    eq(m.alpha_s(), 3 + 6.)
    eq(m.beta_s(), 1 + .5 * (12 * (2. / 3.) ** 2))
    m.tau = 3.
    tauU = 3. * np.array([[2. / 9.] * 2, [1. / 3.] * 2, [2. / 9.] * 2, [2. / 9.] * 2, [1. / 3.] * 2])
    muU = 1. / tauU * (3. * np.array([[10. / 18.] * 2, [15. / 18.] * 2, [10. / 18.] * 2, [10. / 18.] * 2, [15. / 18.] * 2]) - lambdaU)
    tauV = 3. * np.ones((J, K))
    muV = 1. / tauV * (3. * np.full((J, K), 4. * (5. / 6.) * (1. / 2.)) - lambdaV)
    for k in range(K):
        eq(m.tauU(k), tauU[:, k], what="tauU"), eq(m.muU(tauU[:, k], k), muU[:, k], what="muU")
        eq(m.tauV(k), tauV[:, k], what="tauV"), eq(m.muV(tauV[:, k], k), muV[:, k], what="muV")


def check_gibbs_summaries(cls):
    I, J, K = 5, 3, 2
    pri = {'alpha': 3, 'beta': 1, 'lambdaU': 2 * np.ones((I, K)), 'lambdaV': 3 * np.ones((J, K))}
    Us = [np.ones((I, K)) * 3 * n ** 2 for n in range(1, 11)]
    Vs = [np.ones((J, K)) * 2 * n ** 2 for n in range(1, 11)]
    taus = [n ** 2 for n in range(1, 11)]
    m = cls(np.ones((I, J)), np.ones((I, J)), K, pri)
    m.all_U, m.all_V, m.all_tau = Us, Vs, taus
    eU, eV, etau = m.approx_expectation(2, 3)                # samples 2, 5, 8 -> n = 3, 6, 9
    eq(etau, (9. + 36. + 81.) / 3.), eq(eU, np.full((I, K), 9. + 36. + 81.)), eq(eV, np.full((J, K), (9. + 36. + 81.) * 2. / 3.))
    # predict on a test mask
    R = np.arange(1., 16.).reshape(5, 3)
    Us = [u.copy() for u in Us]
    Us[2][0, 0] = 24
    m = cls(R, np.ones((I, J)), K, pri)
    m.all_U, m.all_V, m.all_tau = Us, Vs, taus
    M_test = np.array([[0, 0, 1], [0, 1, 0], [0, 0, 0], [1, 1, 0], [0, 0, 0]])
    perf = m.predict(M_test, 2, 3)
    sq = 444408561. + 447872569. + 447660964. + 447618649
    eq(perf['MSE'], sq / 4.)
    eq(perf['R^2'], 1. - sq / (4.25 ** 2 + 2.25 ** 2 + 2.75 ** 2 + 3.75 ** 2))
    eq(perf['Rp'], 357. / (math.sqrt(44.75) * math.sqrt(5292.)), tol=1e-10)
    # compute_MSE / R2 / Rp and the quality measures
    R2x2, M2x2 = np.array([[1, 2], [3, 4]], dtype=float), np.array([[1, 1], [0, 1]])
    pri3 = {'alpha': 3, 'beta': 1, 'lambdaU': 2 * np.ones((2, 3)), 'lambdaV': 3 * np.ones((2, 3))}
    m = cls(R2x2, M2x2, 3, pri3)
    R_pred, M_pred = np.array([[500, 550], [1220, 1342]], dtype=float), np.array([[0, 0], [1, 1]])
    eq(m.compute_MSE(M_pred, R2x2, R_pred), (1217 ** 2 + 1338 ** 2) / 2.0)
    eq(m.compute_R2(M_pred, R2x2, R_pred), 1. - (1217 ** 2 + 1338 ** 2) / (0.5 ** 2 + 0.5 ** 2))
    eq(m.compute_Rp(M_pred, R2x2, R_pred), 61. / (math.sqrt(.5) * math.sqrt(7442.)), tol=1e-10)
    m.all_U = [np.ones((2, 3)) for _ in range(10)]
    m.all_V = [2 * np.ones((2, 3)) for _ in range(10)]
    m.all_tau = [3. for _ in range(10)]
    ll = 3. / 2. * (math.log(3.) - math.log(2 * math.pi)) - 3. / 2. * (5 ** 2 + 4 ** 2 + 2 ** 2)
    eq(m.quality('loglikelihood', 4, 2), ll)
    eq(m.quality('AIC', 4, 2), -2 * ll + 2 * 12), eq(m.quality('BIC', 4, 2), -2 * ll + 12 * math.log(3))
    eq(m.quality('MSE', 4, 2), (5 ** 2 + 4 ** 2 + 2 ** 2) / 3.)
    try:
        m.quality('FAIL', 4, 2)
    except AssertionError as e:
        assert str(e) == "Unrecognised metric for model quality: FAIL."
    else:
        raise AssertionError("quality('FAIL') did not raise")


def check_vb_elbo(cls):
    """ELBO from a hand-set state (including arbitrary alpha_s, beta_s attributes)."""
    I, J, K, R, M, lambdaU, lambdaV, pri = five_by_three()
    m = cls(R, M, K, pri)
    m.expU, m.expV = 5 * np.ones((I, K)), 6 * np.ones((J, K))
    m.varU, m.varV = 11 * np.ones((I, K)), 12 * np.ones((J, K))
    m.exptau, m.explogtau = 8., 9.
    m.muU, m.muV = 14 * np.ones((I, K)), 15 * np.ones((J, K))
    m.tauU, m.tauV = np.ones((I, K)) / 100., np.ones((J, K)) / 101.
    m.alpha_s, m.beta_s = 20., 21.
    elbo = 12. / 2. * (9. - math.log(2 * math.pi)) - 8. / 2. * (41772 + 19872) \
        + 5 * 2 * (math.log(2.) - 2. * 5.) + 3 * 2 * (math.log(3.) - 3. * 6.) \
        + 3. * np.log(1.) - np.log(math.gamma(3.)) + 2. * 9. - 1. * 8. \
        - 20. * np.log(21.) + np.log(math.gamma(20.)) - 19. * 9. + 21. * 8. \
        - 0.5 * 5 * 2 * math.log(1. / 100.) + 0.5 * 5 * 2 * math.log(2 * math.pi) + 5 * 2 * math.log(1. - 0.080756659233771066) \
        + 0.5 * 5 * 2 * 1. / 100. * (11. + 81.) \
        - 0.5 * 3 * 2 * math.log(1. / 101.) + 0.5 * 3 * 2 * math.log(2 * math.pi) + 3 * 2 * math.log(1. - 0.067776752211548219) \
        + 0.5 * 3 * 2 * 1. / 101. * (12. + 81.)
    eq(m.elbo(), elbo, what="elbo")


def check_vb_updates(cls):
    I, J, K, R, M, lambdaU, lambdaV, pri = five_by_three()

    def fresh():
        m = cls(R, M, K, pri)
        m.expU, m.expV = 1. / lambdaU, 1. / lambdaV
        m.varU, m.varV = np.ones((I, K)) * 2, np.ones((J, K)) * 3
        return m
    m = fresh()
    eq(m.exp_square_diff(), 172.66666666666666)
    m.update_tau()
    eq(m.alpha_s, 3 + 12. / 2.), eq(m.beta_s, 1 + 172.66666666666666 / 2.)
    for k in range(K):
        m = fresh()
        m.muU, m.tauU, m.exptau = np.zeros((I, K)), np.zeros((I, K)), 3.
        m.update_U(k)
        for i in range(I):
            t = 3. * (M[i] * (m.expV[:, k] ** 2 + m.varV[:, k])).sum()
            eq(m.tauU[i, k], t)
            eq(m.muU[i, k], (1. / t) * (-2. + 3. * (M[i] * ((R[i] - np.dot(m.expU[i], m.expV.T) + m.expU[i, k] * m.expV[:, k]) * m.expV[:, k])).sum()))
        m = fresh()
        m.muV, m.tauV, m.exptau = np.zeros((J, K)), np.zeros((J, K)), 3.
        m.update_V(k)
        for j in range(J):
            t = 3. * (M[:, j] * (m.expU[:, k] ** 2 + m.varU[:, k])).sum()
            eq(m.tauV[j, k], t)
            eq(m.muV[j, k], (1. / t) * (-3. + 3. * (M[:, j] * ((R[:, j] - np.dot(m.expU, m.expV[j]) + m.expU[:, k] * m.expV[j, k]) * m.expU[:, k])).sum()))


def check_vb_moments(cls):
    I, J, K, R, M, lambdaU, lambdaV, pri = five_by_three()
    for k in range(K):
        m = cls(R, M, K, pri)
        m.initialise()
        m.tauU = 4 * np.ones((I, K))
        m.update_exp_U(k)
        eq(m.expU[:, k], np.full(I, 0.5 + 1. / 2. * 0.2876155949126352), tol=1e-5)
        eq(m.varU[:, k], np.full(I, 1. / 4. * (1. - 0.37033832534958433)), tol=1e-5)
        m = cls(R, M, K, pri)
        m.initialise()
        m.tauV = 4 * np.ones((J, K))
        m.update_exp_V(k)
        eq(m.expV[:, k], np.full(J, 1. / 3. + 1. / 2. * 0.4273551839464883), tol=1e-5)
        eq(m.varV[:, k], np.full(J, 1. / 4. * (1. - 0.4675359092102624)), tol=1e-5)
    m = cls(R, M, K, pri)
    m.initialise()
    eq(m.exptau, (3 + 12. / 2.) / (1 + 35.4113198623 / 2.), tol=1e-11)
    eq(m.explogtau, 2.1406414779556 - math.log(1 + 35.4113198623 / 2.), tol=1e-11)


def check_constructor_messages(cls, three_factor=False):
    args = (2, 3) if three_factor else (2,)
    pri = ({'alpha': 1, 'beta': 1, 'lambdaF': 1, 'lambdaS': 1, 'lambdaG': 1} if three_factor else
           {'alpha': 1, 'beta': 1, 'lambdaU': 1, 'lambdaV': 1})
    cases = [((np.ones(3), np.ones((2, 3))), "Input matrix R is not a two-dimensional array, but instead 1-dimensional."),
             ((np.ones((4, 3, 2)), np.ones((2, 3))), "Input matrix R is not a two-dimensional array, but instead 3-dimensional."),
             ((np.ones((3, 2)), np.ones((2, 3))), "Input matrix R is not of the same size as the indicator matrix M: (3, 2) and (2, 3) respectively.")]
    for (R, M), msg in cases:
        try:
            cls(R, M, *args, pri)
        except AssertionError as e:
            assert str(e) == msg, (str(e), msg)
        else:
            raise AssertionError("no AssertionError for " + msg)
    R, M = np.ones((2, 3)), np.ones((2, 3))
    M[0] = 0
    try:
        cls(R, M, *args, pri)
    except AssertionError as e:
        assert str(e) == "Fully unobserved row in R, row 0."
    else:
        raise AssertionError("empty row accepted")
    M = np.ones((2, 3))
    M[:, 2] = 0
    try:
        cls(R, M, *args, pri)
    except AssertionError as e:
        assert str(e) == "Fully unobserved column in R, column 2."
    else:
        raise AssertionError("empty column accepted")
