"""Device-backed counterparts of the reference's distribution helpers (code/models/distributions/*.py).

TN_* evaluate / sample N(mu, 1/tau) truncated to [0, inf) in CUDA kernels (bnmtf_tn_moments_f64,
bnmtf_tn_draw_f64: erfc-based moments with the reference's mu < -30 sigma exponential switch; Philox-driven
inverse-CDF / exponential-rejection draws).  The scalar Gamma moments are plain arithmetic and stay on the host;
exponential_draw keeps numpy's generator because the reference's seeded initialisations depend on its stream.
"""
import math

import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream, require_cuda

_counter = [0]


def _next_stream_id():
    _counter[0] += 1
    return _counter[0]


def _seed():
    return int(np.random.randint(0, 2 ** 31 - 1))


def _dev(x, device):
    return torch.from_numpy(np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))).to(device)


# ---- exponential.py / gamma.py / normal.py ---------------------------------------------------------------
def exponential_draw(lambdax):
    return np.random.exponential(scale=1.0 / lambdax, size=None)


def gamma_draw(alpha, beta):
    """One Gamma(shape alpha, rate beta) draw on the device (Marsaglia-Tsang + Philox)."""
    dev = require_cuda()
    out = torch.empty(1, dtype=torch.float64, device=dev)
    _lib.call("bnmtf_gamma_draw_f64", float(alpha), float(beta), 1, _seed(), _next_stream_id(), _ptr(out), _stream())
    return float(out.cpu()[0])


def gamma_draws(alpha, beta, n, seed=None):
    dev = require_cuda()
    out = torch.empty(int(n), dtype=torch.float64, device=dev)
    _lib.call("bnmtf_gamma_draw_f64", float(alpha), float(beta), int(n), _seed() if seed is None else int(seed),
              _next_stream_id(), _ptr(out), _stream())
    return out.cpu().numpy()


def gamma_expectation(alpha, beta):
    return float(alpha) / float(beta)


def gamma_expectation_log(alpha, beta):
    from scipy.special import psi
    return float(psi(float(alpha))) - math.log(float(beta))


def gamma_mode(alpha, beta):
    return (float(alpha) - 1) / float(beta)


def normal_draw(mu, tau):
    return np.random.normal(loc=mu, scale=1.0 / math.sqrt(tau), size=None)


# ---- truncated_normal_vector.py ----------------------------------------------------------------------------
def _moments(mus, taus):
    dev = require_cuda()
    mu, tau = _dev(mus, dev), _dev(taus, dev)
    ex, var = torch.empty_like(mu), torch.empty_like(mu)
    _lib.call("bnmtf_tn_moments_f64", _ptr(mu), _ptr(tau), mu.numel(), _ptr(ex), _ptr(var), _stream())
    return ex.cpu().numpy(), var.cpu().numpy()


def TN_matrix_moments(mus, taus):
    """Expectation and variance of every entry of a matrix of truncated normals in one device call (what the K column-wise
    update_exp_* calls of the reference's initialise() compute: truncated_normal_vector.py:53-73, entry by entry)."""
    shape = np.shape(mus)
    ex, var = _moments(np.ravel(np.asarray(mus, dtype=np.float64)), np.ravel(np.asarray(taus, dtype=np.float64)))
    return ex.reshape(shape), var.reshape(shape)


def TN_vector_expectation(mus, taus):
    return list(_moments(mus, taus)[0])


def TN_vector_variance(mus, taus):
    return list(_moments(mus, taus)[1])


def TN_vector_mode(mus):
    return np.maximum(np.zeros(len(mus)), mus)


def TN_vector_draw(mus, taus, seed=None):
    dev = require_cuda()
    mu, tau = _dev(mus, dev), _dev(taus, dev)
    out = torch.empty_like(mu)
    _lib.call("bnmtf_tn_draw_f64", _ptr(mu), _ptr(tau), mu.numel(), _seed() if seed is None else int(seed),
              _next_stream_id(), _ptr(out), _stream())
    return list(out.cpu().numpy())


# ---- truncated_normal.py (scalar) ---------------------------------------------------------------------------
def TN_draw(mu, tau):
    if tau == 0.:
        return 0.
    return float(TN_vector_draw([mu], [tau])[0])


def TN_expectation(mu, tau):
    return float(_moments([mu], [tau])[0][0])


def TN_variance(mu, tau):
    return float(_moments([mu], [tau])[1][0])


def TN_mode(mu):
    return max(0.0, mu)
