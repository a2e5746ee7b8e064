"""2-GPU run of the row-sharded engine against the single-GPU engine (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
from bnmtf_b200 import parallel, bnmf
rank, world = parallel.init_process_group("nccl")
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
g = dict(np.load(os.path.join(%r, "tests", "golden", "gdsc_bnmf_vb.npz")))
pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
m = bnmf.bnmf_vb_optimised(g["R"], g["M"], int(g["K"]), pri, seed=1, distributed=True)
m.initialise("exp")
m.muU, m.muV = g["init_muU"].copy(), g["init_muV"].copy()
for k in range(int(g["K"])): m.update_exp_U(k)
for k in range(int(g["K"])): m.update_exp_V(k)
m.update_tau(); m.update_exp_tau()
m.run(int(g["its"]))
vb_xfer = m._xfer_bytes()
vb_graph = m._engine()._graph is not None
gb = bnmf.bnmf_gibbs_optimised(g["R"], g["M"], 5, pri, seed=7, distributed=True)
gb.initialise("exp")
ret = gb.run(6)
gb_xfer = gb._xfer_bytes()
eU, eV, etau = gb.approx_expectation(2, 2)
allU = ret[0]                      # first access: own rows of every draw gathered from both ranks, then downloaded
gs = bnmf.bnmf_gibbs_optimised(g["R"], g["M"], 5, pri, seed=7, distributed=True)
gs.initialise("exp")
gs.run(6, summary=(2, 2))
sU, sV, stau = gs.approx_expectation(2, 2)
if rank == 0:
    np.savez(sys.argv[1], expU=m.expU, expV=m.expV, muU=m.muU, tauV=m.tauV, mse=np.array(m.all_performances["MSE"]),
             elbo=np.array(m.all_elbo), gU=gb.U, gtau=np.array(gb.all_tau), allU=allU, eU=eU, eV=eV, sU=sU, sV=sV,
             vb_xfer=np.array(vb_xfer), gb_xfer=np.array(gb_xfer), vb_graph=np.array(vb_graph),
             shared=np.array(not gb.__dict__.get("_no_shared_host", False)))
torch.distributed.barrier()
torch.distributed.destroy_process_group()
'''


def test_two_gpu_sharded_matches_reference_and_single_gpu(tmp_path, golden):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    out = tmp_path / "out.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script), str(out)]
    subprocess.run(cmd, check=True, timeout=600)
    r = dict(np.load(out))
    g = golden("gdsc_bnmf_vb")
    scale = np.abs(g["final_expU"]).max()
    np.testing.assert_allclose(r["expU"], g["final_expU"], rtol=1e-9, atol=1e-11 * scale)
    np.testing.assert_allclose(r["expV"], g["final_expV"], rtol=1e-9, atol=1e-11 * scale)
    np.testing.assert_allclose(r["mse"], g["trace_MSE"], rtol=1e-9)
    ok = np.isfinite(g["trace_elbo"])
    np.testing.assert_allclose(r["elbo"][ok], g["trace_elbo"][ok], rtol=1e-9)
    # Gibbs: the Philox stream is keyed by global row index, so the sharded chain equals the single-GPU chain
    from bnmtf_b200 import bnmf
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
    gb = bnmf.bnmf_gibbs_optimised(g["R"], g["M"], 5, pri, seed=7)
    gb.initialise("exp")
    gb.run(6)
    np.testing.assert_allclose(r["gU"], gb.U, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(r["gtau"], gb.all_tau, rtol=1e-10)
    # every draw, gathered lazily from the two ranks' own rows; posterior means from the stored draws and from running sums
    np.testing.assert_allclose(r["allU"], gb.all_U, rtol=1e-10, atol=1e-12)
    eU, eV, _ = gb.approx_expectation(2, 2)
    for a, b in ((r["eU"], eU), (r["eV"], eV), (r["sU"], eU), (r["sV"], eV)):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    # the other rank's rows of mu / tau arrive through the shared host buffers
    np.testing.assert_allclose(r["muU"], g["final_muU"], rtol=1e-9, atol=1e-11 * np.abs(g["final_muU"]).max())
    np.testing.assert_allclose(r["tauV"], g["final_tauV"], rtol=1e-9)
    assert bool(r["vb_graph"]), "the sharded sweep must have been replayed as a CUDA graph"
    if bool(r["shared"]):
        I, J, K = g["R"].shape[0], g["R"].shape[1], int(g["K"])
        full_vb = 4 * (I + J) * K * 8
        assert r["vb_xfer"][1] < 0.6 * full_vb + 4096, "a rank must only download its own rows: %r of %d" % (r["vb_xfer"], full_vb)


def test_fused_peer_exchange_equals_the_nccl_all_gather():
    """The solver kernels store every finished factor row straight into the other ranks' copies (symmetric memory,
    NVLink P2P) instead of an NCCL all-gather after the kernel: same kernels, same values, only the transport differs,
    so sharded Gibbs and VB runs with BNMTF_PEER=1 and =0 must agree EXACTLY -- for both row solvers (20000 rows per
    rank: thread per row; the small shapes: warp per row) and with an empty-tail shard (tools/peer_check.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29618", os.path.join(ROOT, "tools", "peer_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PEER CHECK OK" in res.stdout
