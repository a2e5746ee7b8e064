"""Drop-in classes for the reference's non-probabilistic models:

    NMF   (code/models/nmf_np.py:32)   -- Lee & Seung multiplicative updates (I-divergence) with a mask
    NMTF  (code/models/nmtf_np.py:39)  -- Yoo & Choi multiplicative tri-factorisation with a mask

Same constructor / initialise / run / train / update_* / predict / compute_I_div surface as the reference.  The
prediction matrix lives in a device scratch buffer and every update is a CUDA kernel (csrc/np.cu).
"""
import itertools
import time

import numpy as np
import torch

from . import _lib
from .bnmf import _metrics_from_sums
from .engine import Dataset, _ptr, _stream, require_cuda


class _NPEngine:
    """Device buffers shared by NMF and NMTF: the dataset in both orientations, the prediction P = A B^T in both
    orientations (rebuilt at the start of each phase), reduction scratch."""

    def __init__(self, R, M, device):
        self.ds = Dataset.from_host(R, M, device)
        ds = self.ds
        f64 = lambda *s: torch.zeros(s, dtype=torch.float64, device=ds.device)
        self.P = f64(ds.I, ds.ldJ)
        self.PT = f64(ds.J, ds.ldI)
        self.nparts = 128
        self.part = f64(self.nparts * 8 + 8)
        self.out8 = f64(8)

    def dev(self, x):
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.ds.device)

    def build_pred(self, A, B, transposed=False):
        ds = self.ds
        if transposed:   # PT = B A^T : rows J
            _lib.call("bnmtf_np_build_pred_f64", _ptr(B), _ptr(A), ds.J, ds.I, ds.ldI, A.shape[1], _ptr(self.PT), _stream())
        else:
            _lib.call("bnmtf_np_build_pred_f64", _ptr(A), _ptr(B), ds.I, ds.J, ds.ldJ, A.shape[1], _ptr(self.P), _stream())

    def row_update(self, A, B, transposed=False):
        """All columns of A (rows of R, or of R^T when transposed) against B; P / PT must be current."""
        ds = self.ds
        if transposed:
            _lib.call("bnmtf_np_row_update_f64", _ptr(ds.RT), _ptr(ds.bitsT), _ptr(self.PT), ds.J, ds.I, ds.ldI,
                      _ptr(A), _ptr(B), A.shape[1], _stream())
        else:
            _lib.call("bnmtf_np_row_update_f64", _ptr(ds.R), _ptr(ds.bits), _ptr(self.P), ds.I, ds.J, ds.ldJ,
                      _ptr(A), _ptr(B), A.shape[1], _stream())

    def sums(self, bits=None):
        ds = self.ds
        _lib.call("bnmtf_np_metrics_f64", _ptr(ds.R), _ptr(ds.bits if bits is None else bits), _ptr(self.P), ds.I, ds.J,
                  ds.ldJ, _ptr(self.part), self.nparts, _ptr(self.out8), _stream())
        return self.out8.cpu().numpy()

    def sums_into(self, trace, row):
        """The eight masked sums of the current prediction -> trace[row] on the device (no host round trip)."""
        ds = self.ds
        _lib.call("bnmtf_np_metrics_f64", _ptr(ds.R), _ptr(ds.bits), _ptr(self.P), ds.I, ds.J, ds.ldJ, _ptr(self.part),
                  self.nparts, _ptr(trace, row), _stream())

    def run_loop(self, iterations, body):
        """iterations x body() with the per-iteration sums kept in a device trace and CUDA events for all_times; the launch
        sequence of an iteration has fixed arguments except the trace row, which advances through a device-side
        pointer table -- so it is captured once and replayed as a CUDA graph (these problems are launch-latency-bound).
        Returns (trace as numpy, cumulative seconds per iteration)."""
        import os
        from .engine import capture_graph, thread_flags
        dev = self.ds.device
        trace = torch.zeros((max(1, iterations), 8), dtype=torch.float64, device=dev)
        start = torch.cuda.Event(enable_timing=True)
        marks = []
        use_graph = int(os.environ.get("BNMTF_GRAPH", "1")) >= 1 and not getattr(thread_flags, "no_graph", False) and iterations >= 8
        graph = None
        start.record()
        for it in range(iterations):
            if use_graph and it >= 1:
                # iterations 1.. write their sums to the scratch row out8; a tiny copy moves them to trace[it] afterwards
                if graph is None:
                    count0 = _lib.launch_count[0]

                    def captured():
                        body()
                        self.sums_into(self.out8.view(1, 8), 0)
                    graph = capture_graph(captured)
                    self._graph_kernels = _lib.launch_count[0] - count0
                    _lib.launch_count[0] = count0
                graph.replay()
                _lib.launch_count[0] += self._graph_kernels
                trace[it].copy_(self.out8)
            else:
                body()
                self.sums_into(trace, it)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
        torch.cuda.synchronize()
        return trace.cpu().numpy()[:iterations], [start.elapsed_time(ev) / 1e3 for ev in marks]


class _NPBase(object):
    def _common_init(self, R, M, device):
        self.R = np.array(R, dtype=float)
        self.M = np.array(M, dtype=float)
        self.metrics = ['MSE', 'R^2', 'Rp']
        assert len(self.R.shape) == 2, "Input matrix R is not a two-dimensional array, " \
            "but instead %s-dimensional." % len(self.R.shape)
        assert self.R.shape == self.M.shape, "Input matrix R is not of the same size as " \
            "the indicator matrix M: %s and %s respectively." % (self.R.shape, self.M.shape)
        (self.I, self.J) = self.R.shape
        self.check_empty_rows_columns()
        # For computing the I-div it is better if unknown values are 1's, not 0's (reference nmf_np.py:48-51)
        self.R_excl_unknown = np.where(self.M != 0, self.R, 1.)
        self._device_arg, self._eng = device, None
        self.verbose = False

    def check_empty_rows_columns(self):
        sums_columns = self.M.sum(axis=0)
        sums_rows = self.M.sum(axis=1)
        for i, c in enumerate(sums_rows):
            assert c != 0, "Fully unobserved row in R, row %s." % i
        for j, c in enumerate(sums_columns):
            assert c != 0, "Fully unobserved column in R, column %s." % j

    def _engine(self):
        if self._eng is None:
            self._eng = _NPEngine(self.R, self.M, require_cuda(self._device_arg))
        return self._eng

    def _dense_sums(self, M, R, R_pred):
        dev = require_cuda(self._device_arg)
        t = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(x, np.shape(R)), dtype=np.float64)).to(dev) for x in (R, R_pred, M)]
        part = torch.zeros(64 * 8, dtype=torch.float64, device=dev)
        out = torch.zeros(8, dtype=torch.float64, device=dev)
        _lib.call("bnmtf_dense_metrics_f64", _ptr(t[0]), _ptr(t[1]), _ptr(t[2]), t[0].numel(), _ptr(part), 64, _ptr(out), _stream())
        return out.cpu().numpy()

    def compute_MSE(self, M, R, R_pred):
        return _metrics_from_sums(self._dense_sums(M, R, R_pred))['MSE']

    def compute_R2(self, M, R, R_pred):
        return _metrics_from_sums(self._dense_sums(M, R, R_pred))['R^2']

    def compute_Rp(self, M, R, R_pred):
        return _metrics_from_sums(self._dense_sums(M, R, R_pred))['Rp']

    def _start_run(self):
        self.all_times = []
        self.all_performances = {}
        for metric in self.metrics:
            self.all_performances[metric] = []
        self.all_i_div = []

    def _record(self, sums, iteration, seconds):
        perf = _metrics_from_sums(sums)
        for metric in self.metrics:
            self.all_performances[metric].append(perf[metric])
        self.all_i_div.append(float(sums[7]))
        if self.verbose:
            print("Iteration %s. I-divergence: %s. MSE: %s. R^2: %s. Rp: %s." % (iteration, sums[7], perf['MSE'], perf['R^2'], perf['Rp']))
        self.all_times.append(seconds)


class NMF(_NPBase):
    def __init__(self, R, M, K, device=None):
        self.K = K
        self._common_init(R, M, device)

    def initialise(self, init_UV='random', expo_prior=1.):
        assert init_UV in ['ones', 'random', 'exponential'], "Unrecognised init option for U,V: %s." % init_UV
        if init_UV == 'ones':
            self.U = np.ones((self.I, self.K))
            self.V = np.ones((self.J, self.K))
        elif init_UV == 'random':
            self.U = np.random.rand(self.I, self.K)
            self.V = np.random.rand(self.J, self.K)
        elif init_UV == 'exponential':
            self.U = np.random.exponential(scale=1.0 / expo_prior, size=(self.I, self.K))
            self.V = np.random.exponential(scale=1.0 / expo_prior, size=(self.J, self.K))

    def run(self, iterations):
        assert hasattr(self, 'U') and hasattr(self, 'V'), "U and V have not been initialised - please run NMF.initialise() first."
        eng = self._engine()
        U, V = eng.dev(self.U), eng.dev(self.V)
        self._start_run()

        def body():
            eng.build_pred(U, V)
            eng.row_update(U, V)
            eng.build_pred(U, V, transposed=True)
            eng.row_update(V, U, transposed=True)
            eng.build_pred(U, V)
        trace, secs = eng.run_loop(iterations, body)
        for it in range(iterations):
            self._record(trace[it], it + 1, secs[it])
        self.U, self.V = U.cpu().numpy(), V.cpu().numpy()

    def train(self, iterations, init_UV='random', expo_prior=1.):
        self.initialise(init_UV=init_UV, expo_prior=expo_prior)
        self.run(iterations=iterations)

    def update_U(self, k):
        self._column(k, 'U')

    def update_V(self, k):
        self._column(k, 'V')

    def _column(self, k, side):
        """White-box single-column update: run the row kernel on a one-column view (A[:,k], B[:,k]) with P current."""
        eng = self._engine()
        U, V = eng.dev(self.U), eng.dev(self.V)
        if side == 'U':
            eng.build_pred(U, V)
            a, b = U[:, k:k + 1].contiguous(), V[:, k:k + 1].contiguous()
            eng.row_update(a, b)
            self.U[:, k] = a.cpu().numpy()[:, 0]
        else:
            eng.build_pred(U, V, transposed=True)
            a, b = V[:, k:k + 1].contiguous(), U[:, k:k + 1].contiguous()
            eng.row_update(a, b, transposed=True)
            self.V[:, k] = a.cpu().numpy()[:, 0]

    def _sums(self, M_pred=None):
        eng = self._engine()
        eng.build_pred(eng.dev(self.U), eng.dev(self.V))
        return eng.sums(None if M_pred is None else eng.ds.pack_mask(M_pred))

    def predict(self, M_pred):
        return _metrics_from_sums(self._sums(M_pred))

    def compute_I_div(self):
        return float(self._sums()[7])

    def give_update(self, iteration):
        sums = self._sums()
        if not hasattr(self, 'all_performances'):
            self._start_run()
        self._record(sums, iteration, time.time())


class NMTF(_NPBase):
    def __init__(self, R, M, K, L, device=None):
        self.K = K
        self.L = L
        self._common_init(R, M, device)

    def initialise(self, init_S='random', init_FG='random', expo_prior=1.):
        assert init_S in ['ones', 'random', 'exponential'], "Unrecognised init option for S: %s." % init_S
        assert init_FG in ['ones', 'random', 'exponential', 'kmeans'], "Unrecognised init option for F,G: %s." % init_FG
        if init_S == 'ones':
            self.S = np.ones((self.K, self.L))
        elif init_S == 'random':
            self.S = np.random.rand(self.K, self.L)
        elif init_S == 'exponential':
            self.S = np.random.exponential(scale=1.0 / expo_prior, size=(self.K, self.L))
        if init_FG == 'ones':
            self.F = np.ones((self.I, self.K))
            self.G = np.ones((self.J, self.L))
        elif init_FG == 'random':
            self.F = np.random.rand(self.I, self.K)
            self.G = np.random.rand(self.J, self.L)
        elif init_FG == 'exponential':
            self.F = np.random.exponential(scale=1.0 / expo_prior, size=(self.I, self.K))
            self.G = np.random.exponential(scale=1.0 / expo_prior, size=(self.J, self.L))
        elif init_FG == 'kmeans':
            from .kmeans import KMeans
            kmeans_F = KMeans(self.R, self.M, self.K)
            kmeans_F.initialise()
            kmeans_F.cluster()
            self.F = kmeans_F.clustering_results + 0.2
            kmeans_G = KMeans(self.R.T, self.M.T, self.L)
            kmeans_G.initialise()
            kmeans_G.cluster()
            self.G = kmeans_G.clustering_results + 0.2

    # ---- device pieces ---------------------------------------------------------------------------------------
    def _mm(self, A, B, transB):
        n, p = A.shape
        q = B.shape[0] if transB else B.shape[1]
        C = torch.empty((n, q), dtype=torch.float64, device=A.device)
        _lib.call("bnmtf_small_matmul_f64", _ptr(A), _ptr(B), n, p, q, 1 if transB else 0, _ptr(C), _stream())
        return C

    def _phase_S(self, eng, F, S, G, pairs):
        ds = eng.ds
        eng.build_pred(self._mm(F, S, False), G)
        for k, l in pairs:
            _lib.call("bnmtf_np_s_update_f64", _ptr(ds.R), _ptr(ds.bits), _ptr(eng.P), ds.I, ds.J, ds.ldJ, _ptr(F), self.K, k,
                      _ptr(G), self.L, l, _ptr(S), _ptr(eng.part), eng.nparts, _stream())

    def _phase_F(self, eng, F, S, G, single=None):
        X = self._mm(G, S, True)              # J x K : (S G^T)^T
        eng.build_pred(F, X)
        if single is None:
            eng.row_update(F, X)
        else:
            a, b = F[:, single:single + 1].contiguous(), X[:, single:single + 1].contiguous()
            eng.row_update(a, b)
            F[:, single] = a[:, 0]

    def _phase_G(self, eng, F, S, G, single=None):
        X = self._mm(F, S, False)             # I x L : F S
        eng.build_pred(X, G, transposed=True)
        if single is None:
            eng.row_update(G, X, transposed=True)
        else:
            a, b = G[:, single:single + 1].contiguous(), X[:, single:single + 1].contiguous()
            eng.row_update(a, b, transposed=True)
            G[:, single] = a[:, 0]

    def _state(self):
        eng = self._engine()
        return eng, eng.dev(self.F), eng.dev(self.S), eng.dev(self.G)

    def run(self, iterations):
        assert hasattr(self, 'F') and hasattr(self, 'S') and hasattr(self, 'G'), \
            "F, S and G have not been initialised - please run NMTF.initialise() first."
        eng, F, S, G = self._state()
        self._start_run()
        pairs = list(itertools.product(range(0, self.K), range(0, self.L)))

        def body():
            self._phase_S(eng, F, S, G, pairs)       # reference order: S, F, G (nmtf_np.py:129-136)
            self._phase_F(eng, F, S, G)
            self._phase_G(eng, F, S, G)
            eng.build_pred(self._mm(F, S, False), G)
        trace, secs = eng.run_loop(iterations, body)
        for it in range(iterations):
            self._record(trace[it], it + 1, secs[it])
        self.F, self.S, self.G = F.cpu().numpy(), S.cpu().numpy(), G.cpu().numpy()

    def train(self, iterations, init_S='random', init_FG='random', expo_prior=1.):
        self.initialise(init_S=init_S, init_FG=init_FG, expo_prior=expo_prior)
        self.run(iterations=iterations)

    def update_F(self, k):
        eng, F, S, G = self._state()
        self._phase_F(eng, F, S, G, single=k)
        self.F = F.cpu().numpy()

    def update_G(self, l):
        eng, F, S, G = self._state()
        self._phase_G(eng, F, S, G, single=l)
        self.G = G.cpu().numpy()

    def update_S(self, k, l):
        eng, F, S, G = self._state()
        self._phase_S(eng, F, S, G, [(k, l)])
        self.S = S.cpu().numpy()

    def triple_dot(self, M1, M2, M3):
        eng = self._engine()
        return self._mm(self._mm(eng.dev(M1), eng.dev(M2), False), eng.dev(M3), False).cpu().numpy()

    def _sums(self, M_pred=None):
        eng, F, S, G = self._state()
        eng.build_pred(self._mm(F, S, False), G)
        return eng.sums(None if M_pred is None else eng.ds.pack_mask(M_pred))

    def predict(self, M_pred):
        return _metrics_from_sums(self._sums(M_pred))

    def compute_I_div(self):
        return float(self._sums()[7])

    def give_update(self, iteration):
        sums = self._sums()
        if not hasattr(self, 'all_performances'):
            self._start_run()
        self._record(sums, iteration, time.time())
