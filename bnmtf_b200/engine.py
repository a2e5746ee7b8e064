"""Device-side state and sweep driver for the two-factor model (R ~ U V^T).

PyTorch is used for exactly three things here: allocating fp64/int32 device buffers, host<->device copies, and
the current CUDA stream.  Every arithmetic step is a call into libbnmtf_b200.so through ctypes (bnmtf_b200/_lib.py).
"""
import math

import numpy as np
import torch

from . import _lib

MODE = {"gibbs": 0, "vb": 1, "icm": 2}
TRACE_COLS = ("tau", "MSE", "R^2", "Rp", "ELBO", "sum_e2", "exp_square_diff", "explogtau")
S_TAU, S_LOGTAU, S_ALPHA_S, S_BETA_S, S_SUM_E2, S_ESD, S_MSE, S_R2, S_RP, S_ELBO = range(10)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.BnmtfError("no CUDA device visible: bnmtf_b200 has no CPU path")
    _lib.load()
    return torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def ld_for(n):
    return (int(n) + 63) // 64 * 64


def kp_for(K):
    return 8 * ((K + 1 + 7) // 8)


def gram_len(K):
    nt = kp_for(K) // 8
    return nt * (nt + 1) // 2 * 64


class Dataset:
    """R and its observation mask on the device, in both orientations (row phase: I x ld(J); column phase:
    J x ld(I)), masks as bit words.  Rows [row_lo,row_hi) / columns [col_lo,col_hi) select the shard this rank
    owns when the matrix is split across GPUs (parallel.py); by default everything."""

    def __init__(self, I, J, device):
        self.I, self.J, self.device = int(I), int(J), device
        self.ldJ, self.ldI = ld_for(J), ld_for(I)
        self.R = self.bits = self.RT = self.bitsT = None
        self.n_obs = None

    @classmethod
    def from_host(cls, R, M, device=None):
        device = require_cuda(device)
        R = np.ascontiguousarray(R, dtype=np.float64)
        M = np.ascontiguousarray(M, dtype=np.float64)
        I, J = R.shape
        ds = cls(I, J, device)
        Rd = torch.from_numpy(R).to(device)
        Md = torch.from_numpy(M).to(device)
        ds.R = torch.empty((I, ds.ldJ), dtype=torch.float64, device=device)
        ds.bits = torch.empty((I, ds.ldJ // 32), dtype=torch.int32, device=device)
        _lib.call("bnmtf_pack_dataset_f64", _ptr(Rd), _ptr(Md), I, J, ds.ldJ, _ptr(ds.R), _ptr(ds.bits), _stream())
        ds._make_transpose()
        ds.n_obs = float(M.sum())
        return ds

    @classmethod
    def from_device(cls, R_padded, bits, I, J, RT_padded=None, bitsT=None, n_obs=None):
        """Adopt already-resident buffers in the library's layout (synthetic benchmark data is generated
        directly on the GPU: a 65536 x 32768 fp64 matrix never exists on the host)."""
        ds = cls(I, J, R_padded.device)
        assert R_padded.shape == (I, ds.ldJ) and bits.shape == (I, ds.ldJ // 32)
        ds.R, ds.bits = R_padded, bits
        if RT_padded is None:
            ds._make_transpose()
        else:
            ds.RT, ds.bitsT = RT_padded, bitsT
        ds.n_obs = float(n_obs) if n_obs is not None else None
        return ds

    def _make_transpose(self):
        I, J = self.I, self.J
        self.RT = torch.empty((J, self.ldI), dtype=torch.float64, device=self.device)
        self.bitsT = torch.empty((J, self.ldI // 32), dtype=torch.int32, device=self.device)
        _lib.call("bnmtf_transpose_dataset_f64", _ptr(self.R), _ptr(self.bits), I, J, self.ldJ,
                  _ptr(self.RT), _ptr(self.bitsT), self.ldI, _stream())

    def pack_mask(self, M):
        """Bit-pack another I x J 0/1 mask (predict(M_pred))."""
        Md = torch.from_numpy(np.ascontiguousarray(M, dtype=np.float64)).to(self.device)
        bits = torch.empty((self.I, self.ldJ // 32), dtype=torch.int32, device=self.device)
        _lib.call("bnmtf_pack_mask_f64", _ptr(Md), self.I, self.J, self.ldJ, _ptr(bits), _stream())
        return bits


class Factor:
    """One factor matrix (n x K) with its variational / conditional parameters and its padded image."""

    def __init__(self, n, K, device, vb):
        self.n, self.K = n, K
        z = lambda: torch.zeros((n, K), dtype=torch.float64, device=device)
        self.fac, self.lam = z(), z()
        self.mu, self.tauf = z(), z()
        self.var = z() if vb else None
        self.n_alloc = ld_for(n) + 8
        KP = kp_for(K)
        self.Xp = torch.zeros((self.n_alloc, KP), dtype=torch.float64, device=device)
        self.Vp = torch.zeros((self.n_alloc, KP), dtype=torch.float64, device=device) if vb else None

    def pad(self):
        _lib.call("bnmtf_pad_factor_f64", _ptr(self.fac), _ptr(self.var), self.n, self.K, self.n_alloc,
                  _ptr(self.Xp), _ptr(self.Vp), _stream())


class BNMFEngine:
    """All device work of bnmf_gibbs_optimised / bnmf_vb_optimised / nmf_icm for one (R, M, K)."""

    def __init__(self, dataset, K, mode, alpha, beta, seed=0, trace_cap=0, shard=None):
        self.ds, self.K, self.mode = dataset, int(K), mode
        self.m = MODE[mode]
        self.vb = mode == "vb"
        self.alpha, self.beta, self.seed = float(alpha), float(beta), int(seed) & (2 ** 64 - 1)
        dev = dataset.device
        I, J = dataset.I, dataset.J
        self.U = Factor(I, K, dev, self.vb)
        self.V = Factor(J, K, dev, self.vb)
        self.scalars = torch.zeros(16, dtype=torch.float64, device=dev)
        self.iter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.iter_scratch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.trace = None
        self.trace_cap = 0
        self.trace_base = 0         # value of the (never reset) sweep counter when the current trace was allocated
        self.sweeps_done = 0        # host mirror of self.iter; also the Philox stream id, so it only ever grows
        self.polarity = 0 if dataset.n_obs is None or dataset.n_obs >= 0.5 * I * J else 1
        KP, GL = kp_for(K), gram_len(K)
        nmax = max(I, J)
        # segment counts: enough CTAs to fill 148 SMs when the matrix has few rows
        self.nseg = {}
        for side, (rows, ld) in enumerate(((I, dataset.ldJ), (J, dataset.ldI))):
            # CTAs per launch: aim for >= 6 waves of resident CTAs (148 SMs x 3 resp. 2 CTAs) so that the tail
            # wave costs little; segments are column ranges whose partial results the solver adds up in order
            rb = (rows + 127) // 128
            nrx = max(1, min(ld // 128, -(-2664 // rb)))
            gb = (rows + 7) // 8
            ng = max(1, min(-(-(ld // 32) // 32), -(-592 // gb)))
            nm = max(1, min(ld // 128, -(-1776 // rb)))
            self.nseg[side] = (nrx, ng, nm)
        mrx = max(self.nseg[0][0] * I, self.nseg[1][0] * J)
        mg = max(self.nseg[0][1] * I, self.nseg[1][1] * J)
        f64 = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
        self.RXpart = f64(mrx, KP)
        self.Gpart = f64(mg, GL)
        self.SVpart = f64(mg, KP) if self.vb else None
        self.Gfull = f64(GL + KP)
        self.gscratch = f64(64 * (GL + KP))
        self.extra = f64(nmax) if self.vb else None
        self.ex1 = f64(1)
        self.el8 = f64(8)
        self.m8 = f64(8)
        self.mpart = f64(((I + 127) // 128) * self.nseg[0][2] * 8)
        self.m8 = f64(8)
        self.nb_terms = 64
        self.elpart = f64(2 * self.nb_terms * 8) if self.vb else None
        self.sterm = None
        self.order_dev = None
        self.alpha_s = None
        self.shard = shard          # set by parallel.ShardedBNMF: (rank, world, row ranges, comm hooks)
        # static sums of the training mask {sum r, sum r^2, |Omega|}: one full-mode metrics pass with zero factors
        self.statics = f64(3)
        _lib.call("bnmtf_masked_metrics_f64", _ptr(dataset.R), _ptr(dataset.bits), I, dataset.ldJ, _ptr(self.U.Xp),
                  _ptr(self.V.Xp), self.K, self.nseg[0][2], 0, _ptr(self.mpart), _ptr(self.m8), _stream())
        self.statics.copy_(self.m8[4:7])
        if dataset.n_obs is None:
            dataset.n_obs = float(self.statics[2].item())
            self.polarity = 0 if dataset.n_obs >= 0.5 * I * J else 1
        self._set_omega(dataset.n_obs)

    # ---- constants depending on |Omega| ---------------------------------------------------------------
    def _set_omega(self, n_obs):
        from scipy.special import gammaln, psi   # host-side scalar constants only
        self.size_Omega = float(n_obs)
        self.alpha_s = self.alpha + self.size_Omega / 2.0
        self.digamma_alpha_s = float(psi(self.alpha_s))
        self.lgamma_alpha = float(gammaln(self.alpha))
        self.lgamma_alpha_s = float(gammaln(self.alpha_s))

    # ---- pieces ---------------------------------------------------------------------------------------
    def _sides(self, side):
        ds = self.ds
        if side == 0:
            return self.U, self.V, ds.R, ds.bits, ds.I, ds.ldJ
        return self.V, self.U, ds.RT, ds.bitsT, ds.J, ds.ldI

    def stats(self, side, need_rx=True):
        """Layer-1 passes for one phase: statistics of the rows of R (side 0) / R^T (side 1) w.r.t. the other factor."""
        me, other, R, bits, rows, ld = self._sides(side)
        nrx, ng, _ = self.nseg[side]
        other.pad()
        if self.polarity == 0:
            _lib.call("bnmtf_gram_full_f64", _ptr(other.Xp), _ptr(other.Vp), other.n, self.K, ld,
                      _ptr(self.Gfull), _ptr(self.gscratch), _stream())
        if need_rx:
            _lib.call("bnmtf_stats_rx_f64", _ptr(R), _ptr(bits), rows, ld, _ptr(other.Xp), self.K, nrx,
                      _ptr(self.RXpart), _stream())
        _lib.call("bnmtf_stats_gram_f64", _ptr(bits), rows, ld, _ptr(other.Xp), _ptr(other.Vp), self.K,
                  self.polarity, ng, _ptr(self.Gpart), _ptr(self.SVpart), _stream())

    def solve(self, side, order=None, n_order=None, apply=True, minimum_TN=0.0, want_sterm=False, want_extra=False,
              use_iter=True):
        me, other, R, bits, rows, ld = self._sides(side)
        nrx, ng, _ = self.nseg[side]
        order_ptr = 0
        if order is not None:
            self.order_dev = torch.tensor(list(order), dtype=torch.int32, device=self.ds.device)
            order_ptr, n_order = _ptr(self.order_dev), len(order)
        elif n_order is None:
            n_order = self.K
        if want_sterm and (self.sterm is None or self.sterm.shape[0] < rows):
            self.sterm = torch.zeros((max(self.ds.I, self.ds.J), self.K), dtype=torch.float64, device=self.ds.device)
        _lib.call("bnmf_row_solve_f64", self.m, rows, self.K, nrx, ng, self.polarity,
                  _ptr(self.RXpart), _ptr(self.Gpart), _ptr(self.SVpart), _ptr(self.Gfull),
                  _ptr(me.fac), _ptr(me.var), _ptr(me.mu), _ptr(me.tauf), _ptr(me.lam),
                  _ptr(self.scalars), order_ptr, n_order, 1 if apply else 0, float(minimum_TN),
                  self.seed, _ptr(self.iter if use_iter else self.iter_scratch), side,
                  _ptr(self.sterm) if want_sterm else 0, _ptr(self.extra) if want_extra else 0, _stream())

    def metrics(self, bits=None):
        """Masked sums over `bits` (default: the training mask) with the current factors -> self.m8."""
        ds = self.ds
        self.U.pad()
        self.V.pad()
        self._metrics_padded(ds.bits if bits is None else bits)

    def _metrics_padded(self, bits):
        ds = self.ds
        statics = _ptr(self.statics) if bits is ds.bits else 0
        _lib.call("bnmtf_masked_metrics_f64", _ptr(ds.R), _ptr(bits), ds.I, ds.ldJ, _ptr(self.U.Xp), _ptr(self.V.Xp),
                  self.K, self.nseg[0][2], statics, _ptr(self.mpart), _ptr(self.m8), _stream())

    def _vb_terms(self):
        nb = self.nb_terms
        for i, f in enumerate((self.U, self.V)):
            _lib.call("bnmtf_vb_factor_terms_f64", _ptr(f.fac), _ptr(f.var), _ptr(f.mu), _ptr(f.tauf), _ptr(f.lam),
                      f.n * f.K, self.elpart[i * nb * 8:].data_ptr(), nb, _stream())
        _lib.call("bnmtf_reduce8_f64", _ptr(self.elpart), 2 * nb, _ptr(self.el8), _stream())

    def finish(self, update_tau=True, record=True):
        # the kernel writes trace row number *iter; rebase the pointer so that row trace_base is row 0
        trace_ptr = _ptr(self.trace) - self.trace_base * 64 if (record and self.trace is not None) else 0
        _lib.call("bnmf_finish_sweep_f64", self.m, self.alpha, self.beta, self.digamma_alpha_s, self.lgamma_alpha,
                  self.lgamma_alpha_s, (self.ds.I + self.ds.J) * self.K, _ptr(self.m8), _ptr(self.ex1), _ptr(self.el8),
                  _ptr(self.scalars), trace_ptr, _ptr(self.iter if record else self.iter_scratch),
                  self.trace_base + self.trace_cap if record else 0, self.seed, 1 if update_tau else 0, _stream())
        if record:
            self.sweeps_done += 1

    def refresh_scalars(self, update_tau=True):
        """Recompute metrics / exp_square_diff / ELBO (and optionally tau) for the CURRENT state without advancing
        the sweep counter: initialise(), exp_square_diff(), elbo(), quality() of the white-box API."""
        if self.vb:
            self.stats(1, need_rx=False)
            self.solve(1, n_order=0, apply=False, want_extra=True, use_iter=False)
            _lib.call("bnmtf_reduce1_f64", _ptr(self.extra), self.ds.J, _ptr(self.ex1), _stream())
            self._vb_terms()
        self.metrics()
        self.finish(update_tau=update_tau, record=False)

    # ---- the sweep ------------------------------------------------------------------------------------
    def sweep(self, minimum_TN=0.0):
        """One iteration of run(): all U columns, all V columns, tau, metrics (reference run() bodies)."""
        self.stats(0)
        self.solve(0, minimum_TN=minimum_TN)
        self.stats(1)
        self.solve(1, minimum_TN=minimum_TN, want_extra=self.vb)
        self.V.pad()                      # U's padded image is current (made for the column phase)
        self._metrics_padded(self.ds.bits)
        if self.vb:
            _lib.call("bnmtf_reduce1_f64", _ptr(self.extra), self.ds.J, _ptr(self.ex1), _stream())
            self._vb_terms()
        self.finish(update_tau=True, record=True)

    def profile_sweep(self, reps=3):
        """Per-kernel CUDA-event timings (ms, mean over reps) of the two streaming passes and the solver, for the
        roofline block of bench.py.  Runs real sweeps (the state advances)."""
        names = ("stats_rx", "stats_gram", "row_solve", "masked_metrics")
        acc = {n: 0.0 for n in names}
        count = {n: 0 for n in names}

        def timed(name, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            acc[name] += e0.elapsed_time(e1)
            count[name] += 1
        for _ in range(reps):
            for side in (0, 1):
                me, other, R, bits, rows, ld = self._sides(side)
                nrx, ng, _ = self.nseg[side]
                other.pad()
                if self.polarity == 0:
                    _lib.call("bnmtf_gram_full_f64", _ptr(other.Xp), _ptr(other.Vp), other.n, self.K, ld,
                              _ptr(self.Gfull), _ptr(self.gscratch), _stream())
                timed("stats_rx", lambda: _lib.call("bnmtf_stats_rx_f64", _ptr(R), _ptr(bits), rows, ld, _ptr(other.Xp),
                                                    self.K, nrx, _ptr(self.RXpart), _stream()))
                timed("stats_gram", lambda: _lib.call("bnmtf_stats_gram_f64", _ptr(bits), rows, ld, _ptr(other.Xp),
                                                      _ptr(other.Vp), self.K, self.polarity, ng, _ptr(self.Gpart),
                                                      _ptr(self.SVpart), _stream()))
                timed("row_solve", lambda: self.solve(side, want_extra=self.vb and side == 1))
            self.V.pad()
            timed("masked_metrics", lambda: self._metrics_padded(self.ds.bits))
            if self.vb:
                _lib.call("bnmtf_reduce1_f64", _ptr(self.extra), self.ds.J, _ptr(self.ex1), _stream())
                self._vb_terms()
            self.finish(update_tau=True, record=False)
        return {n: acc[n] / max(1, count[n]) for n in names}

    def alloc_trace(self, iterations):
        self.trace_cap = int(iterations)
        self.trace = torch.zeros((max(1, self.trace_cap), 8), dtype=torch.float64, device=self.ds.device)
        self.trace_base = self.sweeps_done
