"""World-size-2 tests of the sharding logic on CPU (gloo): partition arithmetic, the two exchanges of a sharded
sweep (all-gather of factor rows, all-reduce of partial sums), and -- with the CPU oracle standing in for the
kernels -- the fact the design rests on: updating each rank's rows independently inside a phase and gathering them
reproduces the unsharded sweep exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bnmtf_b200.engine import Comm, Partition
from bnmtf_b200.parallel import shard_ranges


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run_world(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def test_partition_covers_rows_exactly():
    for n, w in [(100, 2), (80, 3), (7, 8), (65536, 8), (1, 2), (622, 4)]:
        parts = [Partition(n, w, r) for r in range(w)]
        seen = []
        for p in parts:
            assert p.n_pad == p.S * w >= n
            seen += list(range(p.lo(), p.lo() + p.cnt()))
        assert seen == list(range(n))
        assert shard_ranges(n, w) == [(p.lo(), p.cnt()) for p in parts]


def _gather_and_reduce(rank, world):
    part = Partition(7, world, rank)
    full = torch.zeros((part.n_pad, 3), dtype=torch.float64)
    full[rank * part.S:rank * part.S + part.S] = rank + 1.0
    comm = Comm(world, rank)
    comm.gather_rows(full, part)
    red = torch.full((24,), float(rank + 1), dtype=torch.float64)
    comm.allreduce(red)
    return full.numpy().copy(), red.numpy().copy()


def test_comm_gather_rows_and_allreduce_gloo():
    res = run_world(_gather_and_reduce, 2)
    for full, red in res:
        assert (full[:4] == 1.0).all() and (full[4:8] == 2.0).all()
        assert (red == 3.0).all()


def _sharded_oracle_sweep(rank, world):
    """One VB sweep where every rank updates only its own rows of U (then V) with the oracle's column update and the
    rows are exchanged through Comm -- the host-side skeleton of engine.BNMFEngine.sweep()."""
    from oracle import bnmtf_oracle as orc
    rng = np.random.RandomState(0)
    I, J, K = 23, 17, 3
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J))
    M = (rng.rand(I, J) >= 0.2).astype(float)
    M[0, :] = 1
    M[:, 0] = 1
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
    o = orc.OracleBNMF(R, M, K, pri, mode="vb")
    o.init_vb(1.0 / o.lambdaU, 1.0 / o.lambdaV)
    comm = Comm(world, rank)
    for side, n in (("U", I), ("V", J)):
        part = Partition(n, world, rank)
        lo, cnt = part.lo(), part.cnt()
        exp, var = (o.U, o.varU) if side == "U" else (o.V, o.varV)
        for k in range(K):
            tau_k, mu_k = o.column_params(k, side)
            exp[lo:lo + cnt, k] = orc.tn_expectation(mu_k, tau_k)[lo:lo + cnt]   # only my rows
            var[lo:lo + cnt, k] = orc.tn_variance(mu_k, tau_k)[lo:lo + cnt]
        for arr in (exp, var):
            full = torch.zeros((part.n_pad, K), dtype=torch.float64)
            full[:n] = torch.from_numpy(arr)
            comm.gather_rows(full, part)
            arr[:] = full[:n].numpy()
    # partial sum of squared residuals over my rows of R, then all-reduce (the metric exchange)
    part = Partition(I, world, rank)
    lo, cnt = part.lo(), part.cnt()
    e2 = torch.tensor([(M[lo:lo + cnt] * (R[lo:lo + cnt] - o.U[lo:lo + cnt] @ o.V.T) ** 2).sum()], dtype=torch.float64)
    comm.allreduce(e2)
    return o.U.copy(), o.V.copy(), float(e2[0])


def test_sharded_sweep_equals_unsharded_gloo():
    from oracle import bnmtf_oracle as orc
    res = run_world(_sharded_oracle_sweep, 2)
    rng = np.random.RandomState(0)
    I, J, K = 23, 17, 3
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.normal(size=(I, J))
    M = (rng.rand(I, J) >= 0.2).astype(float)
    M[0, :] = 1
    M[:, 0] = 1
    o = orc.OracleBNMF(R, M, K, {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}, mode="vb")
    o.init_vb(1.0 / o.lambdaU, 1.0 / o.lambdaV)
    for k in range(K):
        o._update_column(k, "U")
    for k in range(K):
        o._update_column(k, "V")
    for U, V, e2 in res:
        np.testing.assert_allclose(U, o.U, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(V, o.V, rtol=1e-12, atol=1e-13)
        assert e2 == pytest.approx(o.sum_sq_residual(), rel=1e-12)


def _shared_host(rank, world):
    from bnmtf_b200.bnmf import _shared_pinned
    t, arr, keep = _shared_pinned((6, 3), register=False)     # page-locking needs a GPU; the sharing itself does not
    part = Partition(6, world, rank)
    arr[part.lo():part.lo() + part.cnt()] = rank + 1.0         # each rank fills its own rows, as its device->host copy would
    dist.barrier()
    seen = arr.copy()
    dist.barrier()
    return seen, os.path.exists("/dev/shm/" + keep.name.lstrip("/"))


def test_shared_host_buffers_give_every_rank_the_whole_array():
    """Sharded runs download only a rank's own factor rows; the host arrays live in shared memory so that every rank
    still sees the complete state (bnmtf_b200/bnmf.py::_shared_pinned).  The segment is unlinked as soon as all ranks
    have attached: nothing stays behind in /dev/shm."""
    res = run_world(_shared_host)
    want = np.repeat(np.array([1.0, 1.0, 1.0, 2.0, 2.0, 2.0])[:, None], 3, axis=1)
    for seen, still_there in res:
        np.testing.assert_array_equal(seen, want)
        assert not still_there
