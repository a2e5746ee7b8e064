"""Row/column-sharded runs across the GPUs of one node (one process per GPU, torch.distributed over NCCL).

The sharding itself lives in engine.py (Partition / Comm / Dataset shards); this module holds the launch-side
helpers: process-group setup, synthetic shard generation for the benchmark, and the sharded benchmark leg.
"""
import json
import os

import numpy as np


def init_process_group(backend=None):
    """Join the torchrun rendezvous (MASTER_ADDR/PORT, RANK, WORLD_SIZE from the environment)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def shard_ranges(n, world):
    """[(lo, cnt)] of the contiguous equal-size shards used everywhere (same rule as engine.Partition)."""
    S = -(-n // world)
    return [(min(n, r * S), max(0, min(S, n - min(n, r * S)))) for r in range(world)]


def make_synthetic_shards(I, J, K, device, rank, world, seed=0, tile=4096):
    """This rank's rows of R and of R^T for the synthetic workload of bench.make_synthetic (identical data for any
    world size: every row tile is generated from a tile-keyed generator; a rank regenerates all tiles once to cut
    its column block for the R^T shard)."""
    import torch
    from . import _lib
    from .engine import _ptr, _stream, ld_for
    rng = np.random.RandomState(seed)
    U0 = rng.exponential(1.0, size=(I, K))
    V0 = rng.exponential(1.0, size=(J, K))
    (rlo, rcnt), (clo, ccnt) = shard_ranges(I, world)[rank], shard_ranges(J, world)[rank]
    ldJ, ldI = ld_for(J), ld_for(I)
    f64, i32 = torch.float64, torch.int32
    R = torch.zeros((max(rcnt, 1), ldJ), dtype=f64, device=device)
    bits = torch.zeros((max(rcnt, 1), ldJ // 32), dtype=i32, device=device)
    RT = torch.zeros((max(ccnt, 1), ldI), dtype=f64, device=device)
    bitsT = torch.zeros((max(ccnt, 1), ldI // 32), dtype=i32, device=device)
    V0d = torch.from_numpy(V0).to(device)
    n_obs = 0.0
    assert tile % 64 == 0
    for t0 in range(0, I, tile):
        t1 = min(I, t0 + tile)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + t0)
        U0d = torch.from_numpy(U0[t0:t1]).to(device)
        Rt = U0d @ V0d.T + torch.randn((t1 - t0, J), dtype=f64, device=device, generator=g)
        Mt = (torch.rand((t1 - t0, J), dtype=f64, device=device, generator=g) >= 0.2).to(f64)
        n_obs += float(Mt.sum())
        # rows of this tile that belong to the rank's R shard
        a, b = max(t0, rlo), min(t1, rlo + rcnt)
        if a < b:
            _lib.call("bnmtf_pack_dataset_f64", _ptr(Rt, a - t0), _ptr(Mt, a - t0), b - a, J, ldJ,
                      _ptr(R, a - rlo), _ptr(bits, a - rlo), _stream())
        # the rank's column block of this tile, transposed, goes into columns [t0,t1) of the R^T shard
        if ccnt > 0:
            w = t1 - t0
            wl = ld_for(w)
            tmpR = torch.zeros((ccnt, wl), dtype=f64, device=device)
            tmpB = torch.zeros((ccnt, wl // 32), dtype=i32, device=device)
            RtT = Rt[:, clo:clo + ccnt].T.contiguous()
            MtT = Mt[:, clo:clo + ccnt].T.contiguous()
            _lib.call("bnmtf_pack_dataset_f64", _ptr(RtT), _ptr(MtT), ccnt, w, wl, _ptr(tmpR), _ptr(tmpB), _stream())
            RT[:, t0:t1] = tmpR[:, :w]
            nw = (w + 31) // 32
            bitsT[:, t0 // 32:t0 // 32 + nw] = tmpB[:, :nw]
            del tmpR, tmpB, RtT, MtT
        torch.cuda.synchronize()
        del Rt, Mt
    return R, bits, RT, bitsT, n_obs


def bench_sharded(args, rank, world, device, clock_sampler_cls, measured_peaks, build_roofline):
    """bench.py leg for N > 1: the same 65536 x 32768 problem, rows of R / R^T split over N ranks (strong scaling).
    Timed with CUDA events on every rank, max over ranks, barrier + synchronize on both sides."""
    import torch
    import torch.distributed as dist
    from . import _lib, bnmf, engine
    init_process_group("nccl")
    I, J, K = args.rows, args.cols, args.K
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
    R, bits, RT, bitsT, n_obs = make_synthetic_shards(I, J, K, device, rank, world)
    ds = engine.Dataset.from_device(R, bits, I, J, RT, bitsT, n_obs=n_obs, world=world, rank=rank)
    models = {}
    for mode, cls in (("gibbs", bnmf.bnmf_gibbs_optimised), ("vb", bnmf.bnmf_vb_optimised)):
        m = cls.from_dataset(ds, K, pri, seed=1)
        m.initialise("exp")
        m._push()
        models[mode] = m
    engs = {k: m._engine() for k, m in models.items()}
    for e in engs.values():
        e.alloc_trace(args.warmup + args.steps + 64)
    for _ in range(args.warmup):
        for e in engs.values():
            e.sweep()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = clock_sampler_cls(device.index)
    sampler.start()
    launches0 = _lib.launch_count[0]
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in engs}
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k, e in engs.items():
        ev[k][0].record()
        for _ in range(args.steps):
            e.sweep()
        ev[k][1].record()
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = _lib.launch_count[0] - launches0
    sampler.stop_flag = True
    times = torch.tensor([t0.elapsed_time(t1), ev["gibbs"][0].elapsed_time(ev["gibbs"][1]),
                          ev["vb"][0].elapsed_time(ev["vb"][1])], dtype=torch.float64, device=device)
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, g_ms, v_ms = (float(x) for x in times.cpu())
    prof = {k: e.profile_sweep(reps=1) for k, e in engs.items()}
    mse = {k: float(e.scalars.cpu()[engine.S_MSE]) for k, e in engs.items()}
    sampler.join(timeout=2)
    # end to end through the class API on every rank: replicated factor state uploaded from host numpy, one sharded
    # sweep, state and trace read back; wall clock between barriers, max over ranks
    e2e = None
    if not getattr(args, "no_e2e", False):
        import time
        for m in models.values():
            m.run(1)
        torch.cuda.synchronize()
        dist.barrier()
        w0 = time.time()
        for m in models.values():
            for _ in range(args.steps):
                m.run(1)
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([time.time() - w0], dtype=torch.float64, device=device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        fU, fV = I * K * 8, J * K * 8
        e2e = {"value": 2.0 * args.steps / float(dt.item()), "unit": "sweeps/s",
               "h2d_bytes_per_step": int(world * (2 * (fU + fV) + 5 * (fU + fV)) / 2.0),
               "d2h_bytes_per_step": int(world * (2 * (fU + fV) + 4 * (fU + fV)) / 2.0),
               "note": "model.run(1) per step on every rank (factor state is replicated: each rank uploads and reads back "
                       "all of it); the R / R^T shards stay resident"}
    if rank == 0:
        N = float(I) * J
        value = 2.0 * args.steps / (total_ms / 1e3)
        sweep_s = total_ms / 1e3 / (2.0 * args.steps)
        line = {"metric": "BNMF Gibbs+VB sweeps/sec at %dx%d K=%d" % (I, J, K), "value": value, "unit": "sweeps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / (2.0 * args.steps),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BNMF Gibbs+VB sweep, %dx%d fp64, 20%% missing, K=%d" % (I, J, K),
                           "parallelism": "rows of R and of R^T sharded over %d ranks; all-gather of each updated factor, "
                                          "one 24-double all-reduce per sweep (NCCL)" % world,
                           "l2": "inputs far larger than L2", "gibbs_sweeps_per_s": args.steps / (g_ms / 1e3),
                           "vb_sweeps_per_s": args.steps / (v_ms / 1e3), "train_mse_after": mse},
                "roofline": build_roofline(engs, prof, I, J, K, n_obs, sweep_s, world=world),
                "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary()}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
