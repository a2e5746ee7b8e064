"""The REPLICA decomposition (BASELINE.json configs 1-3 and 5): independent fits, one per GPU.

    python bench.py --workload cv    [--gpus N --steps K --warmup W]     # config 5
    python bench.py --workload small [--gpus N --steps K --warmup W]     # configs 1-3
    (N > 1: torchrun, one rank per GPU, exactly as for the default workload)

cv     VB tri-factorisation of the GDSC IC50 matrix (622 x 138, 81 % observed) under 10-fold cross-validation with the
       (K, L) grid [5..10]^2: 360 (fold, grid point) fits, each = K-means start for F and G (init_FG='kmeans', as in
       the reference's experiments), ITS sweeps, test-set prediction -- the loops of
       code/cross_validation/parallel_matrix_cross_validation.py:52-65 and greedy_search_cross_validation.py:58-101.
       A STEP is one fit; warmup + steps fits are dealt round-robin to the ranks (no collective on the data path);
       value = sweeps/s over all ranks (fits x ITS / time), fits/s beside it.  "scaling": "strong" -- the job list is
       fixed, more GPUs take fewer fits each.
small  the reference's own timing experiments (plots/time_toy, plots/time_Sanger): the eight
       model classes on the toy matrices (100 x 80) and BNMF / BNMTF VB + Gibbs on GDSC, ITS iterations each, one after
       the other; a STEP is one pass over that list; value = problem-iterations/s summed over ranks (every rank runs
       its own chains: "weak"), iterations/s per problem beside it.

nmtf   the tri-factorisation at the headline shape (SURVEY.md section 8d "NMTF at C4-like sizes"; north-star (2)):
       BNMTF Gibbs + VB sweeps on the synthetic 65536 x 32768, 20 %-missing matrix with K = L = 10.  Its statistics
       passes are the two-factor model's tcgen05 kernels (the statistics of R ~ F S G^T w.r.t. G and w.r.t. F are
       two-factor statistics); the K*L sequential S updates run on the (KL x KL) normal equations reduced over rows.
       A STEP is one Gibbs + one VB sweep; value = sweeps/s (every rank its own chains at N > 1: "weak").  The line
       carries a parity block: GPU classes with the tcgen05 statistics forced vs the reference's classes at 2048 x 1024.

All legs time the reference's own classes (baseline/_ref) on the host cores for a bounded sample of the same work
(`cpu_baseline`, rank 0 only); `--impl reference` prints that as the reference arm.
"""
import contextlib
import io
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CV_ITS = 200            # sweeps per fit in the cv leg (the reference's experiment runs 1000; throughput per sweep is the same)
SMALL_ITS = 200
GRID = [(K, L) for K in range(5, 11) for L in range(5, 11)]
FOLDS = 10


def _data(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return g["R"], g["M"]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def priors2():
    return {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}


def priors3():
    return {"alpha": 1.0, "beta": 1.0, "lambdaF": 0.1, "lambdaS": 0.1, "lambdaG": 0.1}


# ------------------------------------------------------------------------------------------------------
def cv_jobs():
    """[(fold, K, L)]: grid points interleaved over the folds so that any prefix of the list mixes sizes."""
    from bnmtf_b200 import mask
    R, M = _data("gdsc_bnmtf_vb")
    random.seed(0)
    folds_test = mask.compute_folds_attempts(R.shape[0], R.shape[1], FOLDS, 100, M)
    folds_train = mask.compute_Ms(folds_test)
    jobs = [(f, K, L) for i, (K, L) in enumerate(GRID) for f in [(i + d) % FOLDS for d in range(FOLDS)]]
    jobs = [jobs[(i * 37) % len(jobs)] for i in range(len(jobs))]          # 37 is coprime to 360: a fixed shuffle
    return R, folds_train, folds_test, jobs


def cv_fit(cls, R, train, test, K, L, its, seed):
    np.random.seed(seed), random.seed(seed)
    m = cls(R, train, K, L, priors3())
    with quiet():
        m.initialise("random", "kmeans")
        m.run(its)
    return m.predict(test)


def small_problems(models):
    toy2, toy3, gd = _data("toy_bnmf_vb"), _data("toy_bnmtf_vb"), _data("gdsc_bnmf_vb")
    P = []
    P.append(("toy BNMF Gibbs K=10", lambda: models.bnmf_gibbs_optimised(toy2[0], toy2[1], 10, priors2()), ("random",), {}))
    P.append(("toy BNMF VB K=10", lambda: models.bnmf_vb_optimised(toy2[0], toy2[1], 10, priors2()), ("random",), {}))
    P.append(("toy NMF ICM K=10", lambda: models.nmf_icm(toy2[0], toy2[1], 10, priors2()), ("random",), {"minimum_TN": 0.1}))
    P.append(("toy NMF NP K=10", lambda: models.NMF(toy2[0], toy2[1], 10), ("random", 0.1), {}))
    P.append(("toy BNMTF Gibbs K=L=5", lambda: models.bnmtf_gibbs_optimised(toy3[0], toy3[1], 5, 5, priors3()), ("random", "random"), {}))
    P.append(("toy BNMTF VB K=L=5", lambda: models.bnmtf_vb_optimised(toy3[0], toy3[1], 5, 5, priors3()), ("random", "random"), {}))
    P.append(("toy NMTF ICM K=L=5", lambda: models.nmtf_icm(toy3[0], toy3[1], 5, 5, priors3()), ("random", "random"), {"minimum_TN": 0.1}))
    P.append(("toy NMTF NP K=L=5", lambda: models.NMTF(toy3[0], toy3[1], 5, 5), ("random", "random", 0.1), {}))
    P.append(("GDSC BNMF VB K=10", lambda: models.bnmf_vb_optimised(gd[0], gd[1], 10, priors2()), ("random",), {}))
    P.append(("GDSC BNMF Gibbs K=10", lambda: models.bnmf_gibbs_optimised(gd[0], gd[1], 10, priors2()), ("random",), {}))
    P.append(("GDSC BNMTF VB K=L=5", lambda: models.bnmtf_vb_optimised(gd[0], gd[1], 5, 5, priors3()), ("random", "random"), {}))
    P.append(("GDSC BNMTF Gibbs K=L=5", lambda: models.bnmtf_gibbs_optimised(gd[0], gd[1], 5, 5, priors3()), ("random", "random"), {}))
    return P


def small_pass(problems, its, seed, built=None):
    """One pass over the problem list; returns seconds per problem.  built: models kept between passes (their
    device datasets stay resident, as a user's model object would)."""
    secs = []
    for i, (name, make, init, kw) in enumerate(problems):
        np.random.seed(seed + i), random.seed(seed + i)
        if built is not None and i in built:
            m = built[i]
        else:
            m = make()
            if built is not None:
                built[i] = m
        t0 = time.time()
        with quiet():
            m.initialise(*init)
            m.run(its, **kw)
        secs.append(time.time() - t0)
    return secs


# ------------------------------------------------------------------------------------------------------
def reference_models():
    import bench
    return bench.load_reference()


def cpu_cv(its_sample=15):
    """One (fold, grid point) fit of the reference's bnmtf_vb_optimised on the host, its_sample sweeps timed after the
    K-means start; seconds per fit at CV_ITS sweeps = start + CV_ITS x seconds per sweep."""
    ref = reference_models()
    if ref is None:
        return None
    R, folds_train, folds_test, jobs = cv_jobs()
    f, K, L = jobs[0]
    np.random.seed(0), random.seed(0)
    t0, c0 = time.time(), time.process_time()
    m = ref.bnmtf_vb_optimised(R, folds_train[f], K, L, priors3())
    with quiet():
        m.initialise("random", "kmeans")
    t_init = time.time() - t0
    t1 = time.time()
    with quiet():
        m.run(its_sample)
    t_sweep = (time.time() - t1) / its_sample
    busy = (time.process_time() - c0) / max(1e-9, time.time() - t0)
    fit_s = t_init + CV_ITS * t_sweep
    return {"seconds_per_fit": fit_s, "seconds_init": t_init, "seconds_per_sweep": t_sweep, "cores_busy": busy,
            "sample": "one (fold, K=%d, L=%d) fit of the reference's bnmtf_vb_optimised (baseline/_ref): K-means start %.2f s + %d "
                      "sweeps at %.4f s; a fit of %d sweeps derived from that" % (K, L, t_init, its_sample, t_sweep, CV_ITS)}


def cpu_small(its_sample=20):
    ref = reference_models()
    if ref is None:
        return None
    out = {}
    c0, w0 = time.process_time(), time.time()
    for name, make, init, kw in small_problems(ref):
        np.random.seed(0), random.seed(0)
        m = make()
        n = its_sample if "GDSC" in name else 3 * its_sample
        with quiet():
            m.initialise(*init)
            t0 = time.time()
            m.run(n, **kw)
        out[name] = n / (time.time() - t0)
    return out, (time.process_time() - c0) / max(1e-9, time.time() - w0)


def nmtf_parity(K, L, shape=(2048, 1024)):
    """GPU tri-factor classes (tcgen05 statistics forced) vs the reference's classes from the same seeded starts."""
    import bench
    import bnmtf_b200
    ref = reference_models()
    I, J = shape
    R, M = bench.sample_problem(I, J, max(K, L))
    R = np.abs(R) + 0.5
    out = {"problem": "%dx%d, K=%d, L=%d, 20%% missing, seeded 'random' starts, BNMTF_NMTF_STATS=umma" % (I, J, K, L),
           "checker": "reference classes (baseline/_ref)" if ref is not None else None}
    if ref is None:
        return out
    os.environ["BNMTF_NMTF_STATS"] = "umma"
    try:
        worst = 0.0
        for name, gcls, rcls, its in (("vb", bnmtf_b200.bnmtf_vb_optimised, ref.bnmtf_vb_optimised, 2),
                                      ("icm", bnmtf_b200.nmtf_icm, ref.nmtf_icm, 2)):
            res = []
            for cls in (gcls, rcls):
                np.random.seed(31), random.seed(31)
                m = cls(R, M, K, L, priors3())
                with quiet():
                    m.initialise("random", "random")
                    m.run(its) if name == "vb" else m.run(its, minimum_TN=0.1)
                res.append(m)
            g, r = res
            if name == "vb":
                assert g._engine().stats_impl == "umma"
                d = {"max_rel_factors": max(bench.rel_err(g.expF, r.expF), bench.rel_err(g.expS, r.expS), bench.rel_err(g.expG, r.expG)),
                     "max_rel_mse": max(abs(a / b - 1.0) for a, b in zip(g.all_performances["MSE"], r.all_performances["MSE"])),
                     "max_rel_exptau": abs(g.exptau / r.exptau - 1.0)}
            else:
                d = {"max_rel_factors": max(bench.rel_err(g.F, r.F), bench.rel_err(g.S, r.S), bench.rel_err(g.G, r.G)),
                     "max_rel_mse": max(abs(a / b - 1.0) for a, b in zip(g.all_performances["MSE"], r.all_performances["MSE"])),
                     "max_rel_tau": abs(g.tau / r.tau - 1.0)}
            out[name] = d
            worst = max(worst, *d.values())
        out["worst"], out["tolerance"], out["pass"] = worst, 1e-9, bool(worst <= 1e-9)
    finally:
        del os.environ["BNMTF_NMTF_STATS"]
    return out


def run_nmtf(args, rank, world, device, dist, barrier, max_over_ranks, sum_over_ranks, sampler, line):
    import torch
    import bench
    from bnmtf_b200 import _lib, bnmtf, engine
    I, J, K = args.rows, args.cols, 10
    R, bits, n_obs = bench.make_synthetic(I, J, K, device)
    ds = engine.Dataset.from_device(R, bits, I, J, n_obs=n_obs)
    models = {}
    for i, (mode, cls) in enumerate((("gibbs", bnmtf.bnmtf_gibbs_optimised), ("vb", bnmtf.bnmtf_vb_optimised))):
        m = cls.from_dataset(ds, K, K, priors3(), seed=1)
        np.random.seed(1 + i), random.seed(1 + i)
        m.initialise("random", "random")
        m._push()
        models[mode] = m
    engs = {k: m._engine() for k, m in models.items()}
    for e in engs.values():
        e.alloc_trace(args.warmup + args.steps + 8)
    for _ in range(args.warmup):
        for e in engs.values():
            e.sweep()
    barrier()
    sampler.start()
    l0 = _lib.launch_count[0]
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in engs}
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k, e in engs.items():
        ev[k][0].record()
        for _ in range(args.steps):
            e.sweep()
        ev[k][1].record()
    t1.record()
    barrier()
    total_ms = max_over_ranks(t0.elapsed_time(t1))
    launches = sum_over_ranks(_lib.launch_count[0] - l0)
    sampler.stop_flag = True
    ms = {k: ev[k][0].elapsed_time(ev[k][1]) for k in engs}
    sc = {k: e.scalars.cpu().numpy() for k, e in engs.items()}
    N = float(I) * J
    hbm, kind = bench.measured_peaks()
    sweep_s = total_ms / 1e3 / (2.0 * args.steps)
    stat = engs["vb"].metrics_mode == "stats"
    # F and G phases stream R once each (as BNMF); the metrics come from the G-phase statistics or from a third pass
    b_alg = 2.0 * N * 8.125 + (0.0 if stat else N * 8.125)
    line.update({"metric": "BNMTF Gibbs+VB sweeps/sec at %dx%d K=L=%d" % (I, J, K), "value": world * 2.0 * args.steps / (total_ms / 1e3),
                 "unit": "sweeps/s", "ms_per_step": total_ms / (2.0 * args.steps), "scaling": "weak",
                 "dtype": bench.DTYPE, "data": "synthetic",
                 "config": {"workload": "BNMTF Gibbs+VB sweep (F, S, G, tau, train metrics), %dx%d fp64, 20%% missing, K=L=%d; "
                                        "every rank its own chains" % (I, J, K),
                            "stats_kernels": engs["vb"].stats_impl, "gibbs_sweeps_per_s": args.steps / (ms["gibbs"] / 1e3),
                            "vb_sweeps_per_s": args.steps / (ms["vb"] / 1e3),
                            "train_MSE_after": {k: float(v[engine.S_MSE]) for k, v in sc.items()},
                            "metrics": engs["vb"].metrics_mode,
                            "l2": "inputs (2 x %.1f GiB of digit planes%s per sweep) far larger than L2"
                                  % (N * 6 / 2 ** 30, "" if stat else " + one 16 GiB pass for the metrics")},
                 "roofline": {"bound": "hbm", "kernel": "whole sweep (two statistics phases on the tcgen05 kernels%s)" % ("" if stat else " + the direct metrics pass"),
                              "achieved": b_alg / sweep_s / 1e9, "peak": hbm, "unit": "GB/s", "frac": b_alg / sweep_s / 1e9 / hbm,
                              "traffic": None, "algorithmic_bytes_per_sweep": b_alg, "peak_kind": kind},
                 "e2e": None, "gpu_launches": int(launches), "clocks": sampler.summary()})
    # end to end: run(1) through the class API (host state in and out)
    if not args.no_e2e:
        for m in models.values():
            m.run(1)
        barrier()
        w0 = time.time()
        for m in models.values():
            for _ in range(max(3, args.steps // 4)):
                m.run(1)
        barrier()
        dt = max_over_ranks(time.time() - w0)
        line["e2e"] = {"value": world * 2.0 * max(3, args.steps // 4) / dt, "unit": "sweeps/s",
                       "h2d_bytes_per_step": int((I + J) * K * 8 * 3.5), "d2h_bytes_per_step": int((I + J) * K * 8 * 3),
                       "note": "model.run(1) per step: factor state uploaded from host numpy, one sweep, state read back"}
    del models, engs, ds, R, bits
    torch.cuda.empty_cache()
    if rank == 0 and not args.no_parity:
        line["parity"] = nmtf_parity(6, 5)


# ------------------------------------------------------------------------------------------------------
def main(args, rank, world):
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    import torch
    import bnmtf_b200
    import bench
    from bnmtf_b200 import _lib
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from bnmtf_b200 import parallel
        parallel.init_process_group("nccl")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t)
        return float(t.item())

    sampler = bench.ClockSampler(local_rank)
    line = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
            "dtype": "f64 (fp64 kernels throughout: the toy / GDSC shapes use the fp64 statistics kernels of the tri-factor "
                     "engine and, for the two-factor models, the tcgen05 fixed-point ones)",
            "data": "the reference's own matrices (toy 100x80, GDSC 622x138) as stored in tests/golden/*.npz"}
    if args.workload == "nmtf":
        run_nmtf(args, rank, world, device, dist, barrier, max_over_ranks, sum_over_ranks, sampler, line)
    elif args.workload == "cv":
        R, folds_train, folds_test, jobs = cv_jobs()
        todo = jobs[:args.warmup + args.steps] if args.warmup + args.steps <= len(jobs) else \
            [jobs[i % len(jobs)] for i in range(args.warmup + args.steps)]
        warm, timed = todo[:args.warmup], todo[args.warmup:]
        for i, (f, K, L) in enumerate(warm):
            if i % world == rank:
                cv_fit(bnmtf_b200.bnmtf_vb_optimised, R, folds_train[f], folds_test[f], K, L, CV_ITS, 1000 + i)
        barrier()
        sampler.start()
        l0 = _lib.launch_count[0]
        t0 = time.time()
        mine = []
        for i, (f, K, L) in enumerate(timed):
            if i % world == rank:
                mine.append((f, K, L, cv_fit(bnmtf_b200.bnmtf_vb_optimised, R, folds_train[f], folds_test[f], K, L, CV_ITS, i)["MSE"]))
        torch.cuda.synchronize()
        dt_own = time.time() - t0
        barrier()
        dt = max_over_ranks(time.time() - t0)
        launches = sum_over_ranks(_lib.launch_count[0] - l0)
        sampler.stop_flag = True
        mse = sum_over_ranks(sum(x[3] for x in mine)) / max(1, len(timed))
        h2d = 3 * R.size * 8 + 0     # per fit: R, train mask, test mask uploaded from host numpy
        line.update({"metric": "VB-NMTF sweeps/sec over the (fold, K, L) fits of a 10-fold CV grid search on GDSC",
                     "value": len(timed) * CV_ITS / dt, "unit": "sweeps/s", "ms_per_step": 1e3 * dt / max(1, len(timed)),
                     "scaling": "strong",
                     "config": {"workload": "BASELINE.json config 5: VB-NMTF, GDSC 622x138, 10 folds x (K,L) in [5..10]^2 = 360 fits; "
                                            "the first %d of the fixed job list dealt round-robin to %d rank(s), %d sweeps per fit, "
                                            "K-means start with the assignment step on the device" % (len(timed), world, CV_ITS),
                                "fits_per_s": len(timed) / dt, "mean_test_MSE": mse, "slowest_rank_busy_s": dt,
                                "this_rank_busy_s": dt_own, "l2": "a fit's working set (1.4 MB) lives in L2; every fit re-uploads its data"},
                     "roofline": None,
                     "e2e": {"value": len(timed) * CV_ITS / dt, "unit": "sweeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64,
                             "note": "the timed region IS the class API: construct, initialise, run, predict -- host data in, metrics out"},
                     "gpu_launches": int(launches), "clocks": sampler.summary()})
        if rank == 0 and not args.no_cpu_baseline:
            c = cpu_cv()
            if c is not None:
                line["cpu_baseline"] = {"value": CV_ITS / c["seconds_per_fit"], "unit": "sweeps/s", "cores": max(1, int(round(c["cores_busy"]))),
                                        "kind": "reference", "sample": c["sample"], "fits_per_s": 1.0 / c["seconds_per_fit"],
                                        "cores_busy_measured": c["cores_busy"], "host_cores": os.cpu_count()}
    else:
        problems = small_problems(bnmtf_b200)
        built = {}
        for w in range(max(1, args.warmup)):
            small_pass(problems, SMALL_ITS, 100 + w, built)
        barrier()
        sampler.start()
        l0 = _lib.launch_count[0]
        t0 = time.time()
        secs = np.zeros(len(problems))
        for s in range(args.steps):
            secs += np.array(small_pass(problems, SMALL_ITS, 200 + s, built))
        torch.cuda.synchronize()
        barrier()
        dt = max_over_ranks(time.time() - t0)
        launches = sum_over_ranks(_lib.launch_count[0] - l0)
        sampler.stop_flag = True
        per = {p[0]: args.steps * SMALL_ITS / float(s) for p, s in zip(problems, secs)}
        line.update({"metric": "problem-iterations/sec over the reference's small timing experiments (toy 100x80, GDSC 622x138)",
                     "value": world * args.steps * SMALL_ITS * len(problems) / dt, "unit": "iterations/s",
                     "ms_per_step": 1e3 * dt / max(1, args.steps), "scaling": "weak",
                     "config": {"workload": "BASELINE.json configs 1-3: %d (model, matrix) problems, initialise + run(%d) each, back "
                                            "to back, every rank its own chains" % (len(problems), SMALL_ITS),
                                "iterations_per_s_per_problem_rank0": per, "l2": "working sets of 64 KB - 0.7 MB live in L2 / shared memory"},
                     "roofline": None,
                     "e2e": {"value": world * args.steps * SMALL_ITS * len(problems) / dt, "unit": "iterations/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0, "note": "the timed region is initialise() + run() through the class API, host state in and out"},
                     "gpu_launches": int(launches), "clocks": sampler.summary()})
        if rank == 0 and not args.no_cpu_baseline:
            c = cpu_small()
            if c is not None:
                per_cpu, busy = c
                tot = sum(1.0 / v for v in per_cpu.values())
                line["cpu_baseline"] = {"value": len(per_cpu) / tot, "unit": "iterations/s", "cores": max(1, int(round(busy))), "kind": "reference",
                                        "sample": "the reference's classes (baseline/_ref) on the same problems, 20-60 iterations each",
                                        "iterations_per_s_per_problem": per_cpu, "cores_busy_measured": busy, "host_cores": os.cpu_count(),
                                        "speedup_per_problem": {k: per[k] / per_cpu[k] for k in per_cpu}}
    if dist is not None:
        barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def run_reference(args):
    t0 = time.time()
    if args.workload == "cv":
        c = cpu_cv()
        if c is None:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing"}))
            return
        v = CV_ITS / c["seconds_per_fit"]
        line = {"impl": "reference", "metric": "VB-NMTF sweeps/sec over the (fold, K, L) fits of a 10-fold CV grid search on GDSC",
                "value": v, "unit": "sweeps/s", "ms_per_step": 1e3 * c["seconds_per_fit"], "scaling": "strong",
                "cpu_baseline": {"value": v, "unit": "sweeps/s", "cores": max(1, int(round(c["cores_busy"]))), "kind": "reference", "sample": c["sample"]}}
    else:
        c = cpu_small()
        if c is None:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing"}))
            return
        per_cpu, busy = c
        v = len(per_cpu) / sum(1.0 / x for x in per_cpu.values())
        line = {"impl": "reference", "metric": "problem-iterations/sec over the reference's small timing experiments (toy 100x80, GDSC 622x138)",
                "value": v, "unit": "iterations/s", "ms_per_step": 1e3 * SMALL_ITS * len(per_cpu) / v, "scaling": "weak",
                "config": {"iterations_per_s_per_problem": per_cpu},
                "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": max(1, int(round(busy))), "kind": "reference",
                                 "sample": "the reference's classes (baseline/_ref), 20-60 iterations per problem"}}
    line.update({"n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                 "dtype": "f64", "data": "the reference's own matrices",
                 "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "wall_s": time.time() - t0})
    print(json.dumps(line))
