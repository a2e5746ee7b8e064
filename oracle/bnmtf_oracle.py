"""CPU oracle for the BNMF / BNMTF update sweep -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's per-iteration update sweep (Gibbs, VB, ICM, NP for NMF and NMTF),
written from the reference's formulas; each function cites the reference file:line it follows (paths are
relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker or the timed CPU baseline.

Parity pinning: tests/test_oracle_vs_reference.py runs this module against the reference's own code
(imported through oracle/ref_shim.py in the build container) on fresh seeded inputs, and
tests/test_oracle_golden.py against the committed golden fixtures in tests/golden/ (generated from the
reference by tests/golden/make_golden.py; runs anywhere).  Deterministic paths
(VB, ICM, NP, and the mu/tau parameters of Gibbs) agree to ~1e-12; the random draws agree in distribution
only (the reference uses the GPL table sampler rtnorm.py, which is deliberately NOT reproduced here --
the oracle draws by inverse CDF, a documented deviation that makes the oracle *faster* than the reference).

The algorithmic cost structure follows the reference on purpose (the full prediction U.V^T is recomputed
for every column update), so that timing this module is a fair stand-in for timing the reference.
"""
import math
import random as _pyrandom

import numpy as np
from scipy.special import erfc, gammaln, ndtri, psi

SQRT2 = math.sqrt(2.0)
LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------------
# distributions  (code/models/distributions/*.py)
# --------------------------------------------------------------------------------------------------
def _clean(v):
    """Reference clamp: negative / non-finite results become 0 (truncated_normal_vector.py:61,73)."""
    v = np.asarray(v, dtype=float)
    return np.where(np.isfinite(v) & (v >= 0.0), v, 0.0)


def _std_pdf(x):
    return np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def tn_expectation(mu, tau):
    """Mean of N(mu,1/tau) truncated to [0,inf) -- truncated_normal_vector.py:53-61, truncated_normal.py:47-55."""
    mu, tau = np.asarray(mu, dtype=float), np.asarray(tau, dtype=float)
    with np.errstate(all="ignore"):
        sigma = 1.0 / np.sqrt(tau)
        x = -mu / sigma
        lam = _std_pdf(x) / (0.5 * erfc(x / SQRT2))
        e = mu + sigma * lam
        e = np.where(mu < -30.0 * sigma, 1.0 / (np.abs(mu) * tau), e)
    return _clean(e)


def tn_variance(mu, tau):
    """Variance of the same -- truncated_normal_vector.py:64-73, truncated_normal.py:58-67."""
    mu, tau = np.asarray(mu, dtype=float), np.asarray(tau, dtype=float)
    with np.errstate(all="ignore"):
        sigma = 1.0 / np.sqrt(tau)
        x = -mu / sigma
        lam = _std_pdf(x) / (0.5 * erfc(x / SQRT2))
        v = sigma ** 2 * (1.0 - lam * (lam - x))
        v = np.where(mu < -30.0 * sigma, (1.0 / (np.abs(mu) * tau)) ** 2, v)
    return _clean(v)


def tn_mode(mu):
    """truncated_normal_vector.py:76-78."""
    return np.maximum(0.0, np.asarray(mu, dtype=float))


def tn_draw(mu, tau, rng):
    """One draw per element from N(mu,1/tau) truncated to [0,inf); tau==0 -> 0 (truncated_normal_vector.py:37-50).

    Same distribution as rtnorm.rtnorm(a=0,b=inf,mu,sigma) (rtnorm.py:16-87) but by inverse CDF on the
    upper tail, with the shifted-exponential limit far in the tail.  `rng` is a numpy RandomState.
    """
    mu, tau = np.atleast_1d(np.asarray(mu, dtype=float)), np.atleast_1d(np.asarray(tau, dtype=float))
    u = rng.uniform(size=mu.shape)
    u = np.clip(u, 1e-300, 1.0)
    with np.errstate(all="ignore"):
        sigma = 1.0 / np.sqrt(tau)
        a = -mu / sigma                               # lower bound in standard units
        tail = 0.5 * erfc(a / SQRT2)                  # P(Z > a)
        z = -ndtri(u * tail)
        far = a > 30.0                                # P(Z>a) < 1e-197: TN ~ a + Exp(rate a)
        z = np.where(far, a - np.log(u) / np.where(far, a, 1.0), z)
        z = np.maximum(z, a)
        d = mu + sigma * z
    d = _clean(d)
    return np.where(tau == 0.0, 0.0, d)


def gamma_expectation(alpha, beta):      # gamma.py:17-19
    return float(alpha) / float(beta)


def gamma_expectation_log(alpha, beta):  # gamma.py:22-24
    return float(psi(float(alpha)) - math.log(float(beta)))


def gamma_mode(alpha, beta):             # gamma.py:27-29
    return (float(alpha) - 1.0) / float(beta)


def gamma_draw(alpha, beta, rng):        # gamma.py:11-14
    return rng.gamma(shape=float(alpha), scale=1.0 / float(beta))


# --------------------------------------------------------------------------------------------------
# metrics  (bnmf_gibbs_optimised.py:208-223, identical in all eight model files)
# --------------------------------------------------------------------------------------------------
def masked_metrics(M, R, P):
    n = float(M.sum())
    mse = (M * (R - P) ** 2).sum() / n
    mean_r = (M * R).sum() / n
    ss_tot = float((M * (R - mean_r) ** 2).sum())
    ss_res = float((M * (R - P) ** 2).sum())
    r2 = 1.0 - ss_res / ss_tot if ss_tot != 0.0 else np.inf
    mean_p = (M * P).sum() / n
    cov = (M * (R - mean_r) * (P - mean_p)).sum()
    var_r = (M * (R - mean_r) ** 2).sum()
    var_p = (M * (P - mean_p) ** 2).sum()
    with np.errstate(all="ignore"):
        rp = cov / float(math.sqrt(var_r) * math.sqrt(var_p))
    return {"MSE": mse, "R^2": r2, "Rp": rp}


def _expand(lam, shape):
    lam = np.array(lam, dtype=float)
    return lam * np.ones(shape) if lam.shape == () else lam.copy()


# --------------------------------------------------------------------------------------------------
# BNMF: Gibbs / VB / ICM / NP
# --------------------------------------------------------------------------------------------------
class OracleBNMF:
    """State + one-sweep update for R ~ U V^T.  mode in {'gibbs','vb','icm','np'}.

    gibbs: bnmf_gibbs_optimised.py:121-177   vb: bnmf_vb_optimised.py:121-215
    icm:   nmf_icm.py:114-168                np: nmf_np.py:86-118
    """

    def __init__(self, R, M, K, priors=None, mode="vb", seed=0):
        self.R, self.M = np.array(R, dtype=float), np.array(M, dtype=float)
        self.I, self.J = self.R.shape
        self.K, self.mode = int(K), mode
        self.size_Omega = self.M.sum()
        self.rng = np.random.RandomState(seed)
        if mode != "np":
            self.alpha, self.beta = float(priors["alpha"]), float(priors["beta"])
            self.lambdaU = _expand(priors["lambdaU"], (self.I, self.K))
            self.lambdaV = _expand(priors["lambdaV"], (self.J, self.K))

    # ---- state -----------------------------------------------------------------------------------
    def set_state(self, U, V, tau=None, varU=None, varV=None, muU=None, muV=None, tauU=None, tauV=None):
        """Gibbs/ICM/NP: U,V (and tau).  VB: U,V are expU,expV and the variational parameters are given."""
        self.U, self.V = np.array(U, dtype=float), np.array(V, dtype=float)
        if self.mode == "vb":
            self.varU, self.varV = np.array(varU, dtype=float), np.array(varV, dtype=float)
            self.muU, self.muV = np.array(muU, dtype=float), np.array(muV, dtype=float)
            self.tauU, self.tauV = np.array(tauU, dtype=float), np.array(tauV, dtype=float)
            self.update_tau_vb()
        elif self.mode in ("gibbs", "icm"):
            self.tau = float(tau) if tau is not None else self._tau_from_state()

    def init_vb(self, muU, muV, tauU=None, tauV=None):
        """bnmf_vb_optimised.initialise (:93-117) after mu has been chosen."""
        tauU = np.ones((self.I, self.K)) if tauU is None else tauU
        tauV = np.ones((self.J, self.K)) if tauV is None else tauV
        self.set_state(tn_expectation(muU, tauU), tn_expectation(muV, tauV),
                       varU=tn_variance(muU, tauU), varV=tn_variance(muV, tauV),
                       muU=muU, muV=muV, tauU=tauU, tauV=tauV)

    def _tau_from_state(self):
        a, b = self.alpha_s(), self.beta_s()
        return a / b if self.mode == "gibbs" else gamma_mode(a, b)   # bnmf_gibbs:117, nmf_icm.py:110

    # ---- shared pieces ---------------------------------------------------------------------------
    def alpha_s(self):          # bnmf_gibbs_optimised.py:161-162
        return self.alpha + self.size_Omega / 2.0

    def sum_sq_residual(self):
        return (self.M * (self.R - self.U @ self.V.T) ** 2).sum()

    def beta_s(self):           # :164-165
        return self.beta + 0.5 * self.sum_sq_residual()

    def exp_square_diff(self):  # bnmf_vb_optimised.py:185-187
        U2, V2 = self.varU + self.U ** 2, self.varV + self.V ** 2
        return (self.M * ((self.R - self.U @ self.V.T) ** 2 + (U2 @ V2.T - (self.U ** 2) @ (self.V ** 2).T))).sum()

    def update_tau_vb(self):    # :181-183, :213-215
        self.alpha_s_, self.beta_s_ = self.alpha + self.size_Omega / 2.0, self.beta + 0.5 * self.exp_square_diff()
        self.exptau = gamma_expectation(self.alpha_s_, self.beta_s_)
        self.explogtau = gamma_expectation_log(self.alpha_s_, self.beta_s_)

    def column_params(self, k, side):
        """(tau_vector, mu_vector) of the conditional / variational factor for column k.

        side 'U': bnmf_gibbs_optimised.py:167-171, bnmf_vb_optimised.py:189-191
        side 'V': :173-177, :193-195
        """
        R, M = (self.R, self.M) if side == "U" else (self.R.T, self.M.T)
        A, B = (self.U, self.V) if side == "U" else (self.V, self.U)
        lam = self.lambdaU if side == "U" else self.lambdaV
        if self.mode == "vb":
            varB = self.varV if side == "U" else self.varU
            prec, weight = self.exptau, varB[:, k] + B[:, k] ** 2
        else:
            prec, weight = self.tau, B[:, k] ** 2
        tau_k = prec * (M * weight).sum(axis=1)
        partial = R - A @ B.T + np.outer(A[:, k], B[:, k])
        with np.errstate(all="ignore"):
            mu_k = 1.0 / tau_k * (-lam[:, k] + prec * (M * (partial * B[:, k])).sum(axis=1))
        return tau_k, mu_k

    def _update_column(self, k, side, minimum_TN=0.0):
        if self.mode == "np":
            return self._np_column(k, side)
        tau_k, mu_k = self.column_params(k, side)
        A = self.U if side == "U" else self.V
        if self.mode == "gibbs":
            A[:, k] = tn_draw(mu_k, tau_k, self.rng)
        elif self.mode == "icm":                                       # nmf_icm.py:129-130
            A[:, k] = np.maximum(tn_mode(mu_k), minimum_TN)
        else:
            mu, tau, var = ((self.muU, self.tauU, self.varU) if side == "U" else (self.muV, self.tauV, self.varV))
            tau[:, k], mu[:, k] = tau_k, mu_k
            A[:, k], var[:, k] = tn_expectation(mu_k, tau_k), tn_variance(mu_k, tau_k)   # :199-211

    def _np_column(self, k, side):                                      # nmf_np.py:114-118
        ratio = self.R / (self.U @ self.V.T)
        if side == "U":
            self.U[:, k] = self.U[:, k] * (self.M * (self.V[:, k] * ratio)).sum(axis=1) / (self.M * self.V[:, k]).sum(axis=1)
        else:
            self.V[:, k] = self.V[:, k] * ((self.U[:, k] * ratio.T).T * self.M).sum(axis=0) / (self.U[:, k] * self.M.T).T.sum(axis=0)

    # ---- one iteration of run() --------------------------------------------------------------------
    def sweep(self, minimum_TN=0.0):
        for k in range(self.K):
            self._update_column(k, "U", minimum_TN)
        for k in range(self.K):
            self._update_column(k, "V", minimum_TN)
        if self.mode == "gibbs":
            self.tau = gamma_draw(self.alpha_s(), self.beta_s(), self.rng)          # :144
        elif self.mode == "icm":
            self.tau = gamma_mode(self.alpha_s(), self.beta_s())                    # nmf_icm.py:137
        elif self.mode == "vb":
            self.update_tau_vb()
        return masked_metrics(self.M, self.R, self.U @ self.V.T)

    def predict(self, M_pred, U=None, V=None):
        U, V = (self.U if U is None else U), (self.V if V is None else V)
        return masked_metrics(np.asarray(M_pred, dtype=float), self.R, U @ V.T)

    def i_divergence(self):     # nmf_np.py:146-148
        Rx = np.where(self.M > 0, self.R, 1.0)
        P = self.U @ self.V.T
        return (self.M * (Rx * np.log(Rx / P) - Rx + P)).sum()

    # ---- model-selection metrics -------------------------------------------------------------------
    def log_likelihood(self, U=None, V=None, tau=None, logtau=None):     # bnmf_vb:258-262, bnmf_gibbs:247-251
        U, V = (self.U if U is None else U), (self.V if V is None else V)
        if self.mode == "vb" and tau is None:
            tau, logtau = self.exptau, self.explogtau
        elif tau is None:
            tau, logtau = self.tau, math.log(self.tau)
        elif logtau is None:
            logtau = math.log(tau)
        return self.size_Omega / 2.0 * (logtau - LOG2PI) - tau / 2.0 * (self.M * (self.R - U @ V.T) ** 2).sum()

    def n_params(self):
        return self.I * self.K + self.J * self.K

    def quality(self, metric, **kw):                                     # bnmf_vb_optimised.py:247-262
        ll = self.log_likelihood(**kw)
        if metric == "loglikelihood":
            return ll
        if metric == "BIC":
            return -2.0 * ll + self.n_params() * math.log(self.size_Omega)
        if metric == "AIC":
            return -2.0 * ll + 2.0 * self.n_params()
        if metric == "MSE":
            U, V = kw.get("U", self.U), kw.get("V", self.V)
            return masked_metrics(self.M, self.R, U @ V.T)["MSE"]
        if metric in ("ELBO", "elbo"):
            return self.elbo() if self.mode == "vb" else 0.0
        raise AssertionError("Unrecognised metric for model quality: %s." % metric)

    def elbo(self):                                                      # bnmf_vb_optimised.py:163-177
        def tn_terms(n, mu, tau, exp, var):
            return (-0.5 * np.log(tau).sum() + n / 2.0 * LOG2PI
                    + np.log(0.5 * erfc(-mu * np.sqrt(tau) / SQRT2)).sum()
                    + (tau / 2.0 * (var + (exp - mu) ** 2)).sum())
        a, b, a_s, b_s = self.alpha, self.beta, self.alpha_s_, self.beta_s_
        return (self.size_Omega / 2.0 * (self.explogtau - LOG2PI) - self.exptau / 2.0 * self.exp_square_diff()
                + np.log(self.lambdaU).sum() - (self.lambdaU * self.U).sum()
                + np.log(self.lambdaV).sum() - (self.lambdaV * self.V).sum()
                + a * math.log(b) - gammaln(a) + (a - 1.0) * self.explogtau - b * self.exptau
                - a_s * math.log(b_s) + gammaln(a_s) - (a_s - 1.0) * self.explogtau + b_s * self.exptau
                + tn_terms(self.I * self.K, self.muU, self.tauU, self.U, self.varU)
                + tn_terms(self.J * self.K, self.muV, self.tauV, self.V, self.varV))


# --------------------------------------------------------------------------------------------------
# BNMTF: Gibbs / VB / ICM / NP
# --------------------------------------------------------------------------------------------------
class OracleBNMTF:
    """State + one-sweep update for R ~ F S G^T.

    gibbs: bnmtf_gibbs_optimised.py:138-211   vb: bnmtf_vb_optimised.py:160-293
    icm:   nmtf_icm.py:132-204                np: nmtf_np.py:120-174
    """

    def __init__(self, R, M, K, L, priors=None, mode="vb", seed=0):
        self.R, self.M = np.array(R, dtype=float), np.array(M, dtype=float)
        self.I, self.J = self.R.shape
        self.K, self.L, self.mode = int(K), int(L), mode
        self.size_Omega = self.M.sum()
        self.rng = np.random.RandomState(seed)
        if mode != "np":
            self.alpha, self.beta = float(priors["alpha"]), float(priors["beta"])
            self.lambdaF = _expand(priors["lambdaF"], (self.I, self.K))
            self.lambdaS = _expand(priors["lambdaS"], (self.K, self.L))
            self.lambdaG = _expand(priors["lambdaG"], (self.J, self.L))

    def set_state(self, F, S, G, tau=None, var=None, mu=None, taus=None):
        """var/mu/taus: dicts with keys 'F','S','G' (VB only)."""
        self.F, self.S, self.G = (np.array(x, dtype=float) for x in (F, S, G))
        if self.mode == "vb":
            self.varF, self.varS, self.varG = (np.array(var[k], dtype=float) for k in "FSG")
            self.muF, self.muS, self.muG = (np.array(mu[k], dtype=float) for k in "FSG")
            self.tauF, self.tauS, self.tauG = (np.array(taus[k], dtype=float) for k in "FSG")
            self.update_tau_vb()
        elif self.mode in ("gibbs", "icm"):
            if tau is not None:
                self.tau = float(tau)
            else:
                a, b = self.alpha_s(), self.beta_s()
                self.tau = a / b if self.mode == "gibbs" else gamma_mode(a, b)

    def init_vb(self, muF, muS, muG, taus=None):
        """bnmtf_vb_optimised.initialise (:104-156) after the mu's have been chosen."""
        taus = taus or {}
        tF = taus.get("F", np.ones((self.I, self.K)))
        tS = taus.get("S", np.ones((self.K, self.L)))
        tG = taus.get("G", np.ones((self.J, self.L)))
        self.set_state(tn_expectation(muF, tF), tn_expectation(muS, tS), tn_expectation(muG, tG),
                       var={"F": tn_variance(muF, tF), "S": tn_variance(muS, tS), "G": tn_variance(muG, tG)},
                       mu={"F": muF, "S": muS, "G": muG}, taus={"F": tF, "S": tS, "G": tG})

    def pred(self):
        return self.F @ (self.S @ self.G.T)

    def alpha_s(self):
        return self.alpha + self.size_Omega / 2.0

    def beta_s(self):           # bnmtf_gibbs_optimised.py:192-193
        return self.beta + 0.5 * (self.M * (self.R - self.pred()) ** 2).sum()

    def exp_square_diff(self):  # bnmtf_vb_optimised.py:239-243
        F, S, G, vF, vS, vG = self.F, self.S, self.G, self.varF, self.varS, self.varG
        M = self.M
        t1 = (M * (self.R - self.pred()) ** 2).sum()
        t2 = (M * ((vF + F ** 2) @ (vS + S ** 2) @ (vG + G ** 2).T - (F ** 2) @ (S ** 2) @ (G ** 2).T)).sum()
        t3 = (M * (vF @ ((S @ G.T) ** 2 - (S ** 2) @ (G.T ** 2)))).sum()
        t4 = (M * (((F @ S) ** 2 - (F ** 2) @ (S ** 2)) @ vG.T)).sum()
        return t1 + t2 + t3 + t4

    def update_tau_vb(self):    # :235-237, :291-293
        self.alpha_s_, self.beta_s_ = self.alpha + self.size_Omega / 2.0, self.beta + 0.5 * self.exp_square_diff()
        self.exptau = gamma_expectation(self.alpha_s_, self.beta_s_)
        self.explogtau = gamma_expectation_log(self.alpha_s_, self.beta_s_)

    # ---- Gibbs / ICM conditionals (bnmtf_gibbs_optimised.py:195-211) -----------------------------
    def params_F(self, k):
        x = self.S[k] @ self.G.T
        tau_k = self.tau * (self.M * x ** 2).sum(axis=1)
        part = self.R - self.pred() + np.outer(self.F[:, k], x)
        with np.errstate(all="ignore"):
            mu_k = 1.0 / tau_k * (-self.lambdaF[:, k] + self.tau * (self.M * (part * x)).sum(axis=1))
        return tau_k, mu_k

    def params_G(self, l):
        x = self.F @ self.S[:, l]
        tau_l = self.tau * (self.M.T * x ** 2).sum(axis=1)
        part = self.R - self.pred() + np.outer(x, self.G[:, l])
        with np.errstate(all="ignore"):
            mu_l = 1.0 / tau_l * (-self.lambdaG[:, l] + self.tau * (self.M * (part.T * x).T).sum(axis=0))
        return tau_l, mu_l

    def params_S(self, k, l):
        fg = np.outer(self.F[:, k], self.G[:, l])
        tau_kl = self.tau * (self.M * fg ** 2).sum()
        part = self.R - self.pred() + self.S[k, l] * fg
        with np.errstate(all="ignore"):
            mu_kl = 1.0 / tau_kl * (-self.lambdaS[k, l] + self.tau * (self.M * (part * fg)).sum())
        return tau_kl, mu_kl

    # ---- VB updates (bnmtf_vb_optimised.py:245-280) ---------------------------------------------------
    def vb_update_F(self, k):
        F, S, G, vS, vG, M = self.F, self.S, self.G, self.varS, self.varG, self.M
        x = S[k] @ G.T
        extra = (vS[k] + S[k] ** 2) @ (vG + G ** 2).T - (S[k] ** 2) @ (G ** 2).T
        self.tauF[:, k] = self.exptau * (M * (extra + x ** 2)).sum(axis=1)
        diff = (M * ((self.R - self.pred() + np.outer(F[:, k], x)) * x)).sum(axis=1)
        cov = (M * ((S[k] * (F @ S)) @ vG.T - np.outer(F[:, k], (S[k] ** 2) @ vG.T))).sum(axis=1)
        self.muF[:, k] = 1.0 / self.tauF[:, k] * (-self.lambdaF[:, k] + self.exptau * diff - self.exptau * cov)

    def vb_update_G(self, l):
        F, S, G, vF, vS, M = self.F, self.S, self.G, self.varF, self.varS, self.M
        x = F @ S[:, l]
        extra = (vF + F ** 2) @ (vS[:, l] + S[:, l] ** 2) - (F ** 2) @ (S[:, l] ** 2)
        self.tauG[:, l] = self.exptau * (M.T * (extra + x ** 2)).sum(axis=1)
        diff = (M * ((self.R - self.pred() + np.outer(x, G[:, l])).T * x).T).sum(axis=0)
        cov = (M * (vF @ (S[:, l] * (S @ G.T).T).T - np.outer(vF @ (S[:, l] ** 2), G[:, l]))).sum(axis=0)
        self.muG[:, l] = 1.0 / self.tauG[:, l] * (-self.lambdaG[:, l] + self.exptau * diff - self.exptau * cov)

    def vb_update_S(self, k, l):
        F, S, G, vF, vG, M = self.F, self.S, self.G, self.varF, self.varG, self.M
        fg = np.outer(F[:, k], G[:, l])
        self.tauS[k, l] = self.exptau * (M * np.outer(vF[:, k] + F[:, k] ** 2, vG[:, l] + G[:, l] ** 2)).sum()
        diff = (M * ((self.R - self.pred() + S[k, l] * fg) * fg)).sum()
        cov_G = (M * np.outer(F[:, k] * (F @ S[:, l] - F[:, k] * S[k, l]), vG[:, l])).sum()
        cov_F = (M * np.outer(vF[:, k], G[:, l] * (S[k] @ G.T - S[k, l] * G[:, l]))).sum()
        self.muS[k, l] = 1.0 / self.tauS[k, l] * (-self.lambdaS[k, l] + self.exptau * (diff - cov_G - cov_F))

    # ---- NP multiplicative updates (nmtf_np.py:155-174) ------------------------------------------------
    def np_update_F(self, k):
        x = self.S[k] @ self.G.T
        self.F[:, k] *= (self.M * self.R / self.pred() * x).sum(axis=1) / (self.M * x).sum(axis=1)

    def np_update_G(self, l):
        x = self.F @ self.S[:, l]
        self.G[:, l] *= ((self.M * self.R / self.pred()).T * x).T.sum(axis=0) / (self.M.T * x).T.sum(axis=0)

    def np_update_S(self, k, l):
        fg = self.M * np.outer(self.F[:, k], self.G[:, l])
        self.S[k, l] *= (self.R * fg / self.pred()).sum() / fg.sum()

    # ---- one iteration ---------------------------------------------------------------------------------
    def _apply(self, target, idx, tau_p, mu_p, minimum_TN):
        if self.mode == "gibbs":
            target[idx] = tn_draw(mu_p, tau_p, self.rng).reshape(np.shape(target[idx]))
        else:
            target[idx] = np.maximum(tn_mode(mu_p), minimum_TN)

    def sweep(self, minimum_TN=0.0, order=None):
        """order (VB only): dict with 'S' (list of (k,l)), 'F' (list of k), 'G' (list of l); the reference
        shuffles these with python `random.shuffle` every iteration (bnmtf_vb_optimised.py:171-190)."""
        K, L = self.K, self.L
        kl = [(k, l) for k in range(K) for l in range(L)]
        if self.mode == "vb":
            order = order or {"S": kl, "F": list(range(K)), "G": list(range(L))}
            for k, l in order["S"]:
                self.vb_update_S(k, l)
                self.S[k, l], self.varS[k, l] = tn_expectation(self.muS[k, l], self.tauS[k, l]), tn_variance(self.muS[k, l], self.tauS[k, l])
            for k in order["F"]:
                self.vb_update_F(k)
                self.F[:, k], self.varF[:, k] = tn_expectation(self.muF[:, k], self.tauF[:, k]), tn_variance(self.muF[:, k], self.tauF[:, k])
            for l in order["G"]:
                self.vb_update_G(l)
                self.G[:, l], self.varG[:, l] = tn_expectation(self.muG[:, l], self.tauG[:, l]), tn_variance(self.muG[:, l], self.tauG[:, l])
            self.update_tau_vb()
        elif self.mode == "np":                                          # nmtf_np.py:129-136: S, F, G
            for k, l in kl:
                self.np_update_S(k, l)
            for k in range(K):
                self.np_update_F(k)
            for l in range(L):
                self.np_update_G(l)
        else:                                                            # bnmtf_gibbs:152-166: F, S, G
            for k in range(K):
                self._apply(self.F, (slice(None), k), *self.params_F(k), minimum_TN)
            for k, l in kl:
                self._apply(self.S, (k, l), *self.params_S(k, l), minimum_TN)
            for l in range(L):
                self._apply(self.G, (slice(None), l), *self.params_G(l), minimum_TN)
            a, b = self.alpha_s(), self.beta_s()
            self.tau = gamma_draw(a, b, self.rng) if self.mode == "gibbs" else gamma_mode(a, b)
        return masked_metrics(self.M, self.R, self.pred())

    @staticmethod
    def shuffled_order(K, L):
        """The three python-`random` shuffles of one bnmtf_vb iteration, in the reference's call order."""
        kl = [(k, l) for k in range(K) for l in range(L)]
        _pyrandom.shuffle(kl)
        ks = list(range(K))
        _pyrandom.shuffle(ks)
        ls = list(range(L))
        _pyrandom.shuffle(ls)
        return {"S": kl, "F": ks, "G": ls}

    def predict(self, M_pred, F=None, S=None, G=None):
        F, S, G = (self.F if F is None else F), (self.S if S is None else S), (self.G if G is None else G)
        return masked_metrics(np.asarray(M_pred, dtype=float), self.R, F @ (S @ G.T))

    def i_divergence(self):     # nmtf_np.py:202-204
        Rx = np.where(self.M > 0, self.R, 1.0)
        P = self.pred()
        return (self.M * (Rx * np.log(Rx / P) - Rx + P)).sum()

    def n_params(self):
        return self.I * self.K + self.K * self.L + self.J * self.L

    def log_likelihood(self, F=None, S=None, G=None, tau=None, logtau=None):
        F, S, G = (self.F if F is None else F), (self.S if S is None else S), (self.G if G is None else G)
        if self.mode == "vb" and tau is None:
            tau, logtau = self.exptau, self.explogtau
        elif tau is None:
            tau, logtau = self.tau, math.log(self.tau)
        elif logtau is None:
            logtau = math.log(tau)
        return self.size_Omega / 2.0 * (logtau - LOG2PI) - tau / 2.0 * (self.M * (self.R - F @ (S @ G.T)) ** 2).sum()

    def quality(self, metric, **kw):
        ll = self.log_likelihood(**kw)
        if metric == "loglikelihood":
            return ll
        if metric == "BIC":
            return -2.0 * ll + self.n_params() * math.log(self.size_Omega)
        if metric == "AIC":
            return -2.0 * ll + 2.0 * self.n_params()
        if metric == "MSE":
            return self.predict(self.M, kw.get("F"), kw.get("S"), kw.get("G"))["MSE"]
        if metric in ("ELBO", "elbo"):
            return self.elbo() if self.mode == "vb" else 0.0
        raise AssertionError("Unrecognised metric for model quality: %s." % metric)

    def elbo(self):             # bnmtf_vb_optimised.py:208-226
        def tn_terms(n, mu, tau, exp, var):
            return (-0.5 * np.log(tau).sum() + n / 2.0 * LOG2PI
                    + np.log(0.5 * erfc(-mu * np.sqrt(tau) / SQRT2)).sum()
                    + (tau / 2.0 * (var + (exp - mu) ** 2)).sum())
        a, b, a_s, b_s = self.alpha, self.beta, self.alpha_s_, self.beta_s_
        return (self.size_Omega / 2.0 * (self.explogtau - LOG2PI) - self.exptau / 2.0 * self.exp_square_diff()
                + np.log(self.lambdaF).sum() - (self.lambdaF * self.F).sum()
                + np.log(self.lambdaS).sum() - (self.lambdaS * self.S).sum()
                + np.log(self.lambdaG).sum() - (self.lambdaG * self.G).sum()
                + a * math.log(b) - gammaln(a) + (a - 1.0) * self.explogtau - b * self.exptau
                - a_s * math.log(b_s) + gammaln(a_s) - (a_s - 1.0) * self.explogtau + b_s * self.exptau
                + tn_terms(self.I * self.K, self.muF, self.tauF, self.F, self.varF)
                + tn_terms(self.K * self.L, self.muS, self.tauS, self.S, self.varS)
                + tn_terms(self.J * self.L, self.muG, self.tauG, self.G, self.varG))
