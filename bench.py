#!/usr/bin/env python
"""Headline benchmark: BNMF Gibbs + VB update sweeps on synthetic 65536 x 32768 fp64, 20 % missing, K=20
(BASELINE.json config 4).  One JSON line on stdout; see DESIGN.md section "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows I --cols J --K K]

A "step" is one Gibbs sweep plus one VB sweep (all U columns, all V columns, tau, train metrics each);
value = sweeps per second over both.  N>1 (torchrun): rows of R / R^T are sharded across ranks
(bnmtf_b200/parallel.py) -- strong scaling of the same matrix.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0
# ncu dram bytes (read + write) per matrix entry of one k_rx_umma launch, by digit count (profiles/r01_ncu_summary_final.txt: 7, profiles/r01d_ncu_summary.txt: 6)
RX_TRAFFIC_PER_ENTRY = {7: 15.15e9 / 2.0 ** 31, 6: 12.98e9 / 2.0 ** 31}
GRAM_TRAFFIC_PER_LAUNCH = {7: 1.6e9, 6: 1.29e9}       # mean of the two phases, at the full 65536 x 32768 shape on one GPU
FP64_PEAK_TFLOPS = 37.1   # measured here: tools/microbench/fp64_pipes.cu -> profiles/r01_microbench_fp64_pipes.txt


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
def make_synthetic(I, J, K, device, seed=0, row_lo=0, row_hi=None, tile=4096):
    """U0,V0 ~ Exp(1) (numpy RandomState(seed)), R = U0 V0^T + N(0,1), M = rand >= 0.2, generated in
    row tiles directly into the library's device layout (rows [row_lo,row_hi) only).  torch is the data generator
    here -- this is input synthesis, outside every timed region."""
    import torch
    from bnmtf_b200 import _lib
    from bnmtf_b200.engine import _ptr, _stream, ld_for
    rng = np.random.RandomState(seed)
    U0 = rng.exponential(1.0, size=(I, K))
    V0 = rng.exponential(1.0, size=(J, K))
    row_hi = I if row_hi is None else row_hi
    n = row_hi - row_lo
    ld = ld_for(J)
    R = torch.zeros((n, ld), dtype=torch.float64, device=device)
    bits = torch.zeros((n, ld // 32), dtype=torch.int32, device=device)
    V0d = torch.from_numpy(V0).to(device)
    n_obs = 0
    for t0 in range(0, n, tile):
        t1 = min(n, t0 + tile)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + (row_lo + t0))       # tile-keyed: identical data for any sharding
        U0d = torch.from_numpy(U0[row_lo + t0:row_lo + t1]).to(device)
        Rt = U0d @ V0d.T + torch.randn((t1 - t0, J), dtype=torch.float64, device=device, generator=g)
        Mt = (torch.rand((t1 - t0, J), dtype=torch.float64, device=device, generator=g) >= 0.2).to(torch.float64)
        n_obs += float(Mt.sum())
        _lib.call("bnmtf_pack_dataset_f64", _ptr(Rt), _ptr(Mt), t1 - t0, J, ld, R[t0:].data_ptr(), bits[t0:].data_ptr(), _stream())
        torch.cuda.synchronize()
        del Rt, Mt
    return R, bits, n_obs


def build_roofline(engs, prof, I, J, K, n_obs, sweep_s, world=1):
    """Roofline block of the JSON line.  The dominant kernel of a sweep is the per-row Gram
    (tcgen05 int8 tensor pipe); the R-streaming kernel is reported beside it against the HBM roof.  world > 1: the
    per-launch figures are per rank (1/world of the entries), the whole-sweep HBM figure is against world x peak."""
    hbm_peak, hbm_kind = measured_peaks()
    N = float(I) * J / world
    n_obs = n_obs / world
    e0 = next(iter(engs.values()))
    # the Gram kernel is timed inside a long, power-capped run: the sustained cuBLAS figure is its denominator
    # (B200_PROFILING.md); the burst figure is reported beside it
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            pk = json.load(fh)
        bf16, bf16_burst = float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])), float(pk["bf16_tflops"])
        bf16_kind = "measured, sustained"
    except Exception:
        bf16, bf16_burst, bf16_kind = 1400.0, 1590.0, "fallback, sustained"
    b_alg_sweep = 2.0 * N * world * 8.125
    out = {"kernel_ms": prof,
           "sweep_hbm": {"algorithmic_bytes_per_sweep": b_alg_sweep, "achieved_gbs": b_alg_sweep / sweep_s / 1e9,
                         "peak_gbs": hbm_peak * world, "peak_kind": hbm_kind,
                         "frac": b_alg_sweep / sweep_s / 1e9 / (hbm_peak * world)}}
    from bnmtf_b200 import _lib
    digits = _lib.call("bnmtf_fixed_point_digits")       # bytes per fixed-point image (6 by default)
    out["fixed_point_digits"] = digits
    if e0.gram == "umma":
        # exact fixed-point Gram: 0/1 selection matrix (rows x cols) times the int8 digit slices of K(K+1)/2 (+K
        # variance, VB) (+K column-sum, metrics phase) product columns; algorithmic ops = 2 * rows * cols * digit columns
        ops, ach = {}, {}
        for k, e in engs.items():
            nc = K * (K + 1) // 2 + (K if e.vb else 0)
            ops[k] = 2.0 * N * (nc + 0.5 * (K if e.metrics_mode == "stats" else 0)) * digits
            ach[k] = ops[k] / (prof[k]["stats_gram"] * 1e-3) / 1e12
        mean_ach = sum(ach.values()) / len(ach)
        peak = 2.0 * bf16
        out.update({"bound": "tensor", "kernel": "k_gram_umma (tcgen05.mma kind::i8, TMEM accumulators, TMA-fed)",
                    "achieved": mean_ach, "peak": peak, "unit": "TFLOP/s", "frac": mean_ach / peak, "traffic": (GRAM_TRAFFIC_PER_LAUNCH.get(digits) / world if GRAM_TRAFFIC_PER_LAUNCH.get(digits) else None),
                    "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (%s): the int8 tensor rate of B200 is "
                                   "twice the bf16 rate; ops are int8 multiply-adds x 2; the kernel is timed inside the "
                                   "power-capped sweep loop" % bf16_kind,
                    "frac_of_burst_peak": mean_ach / (2.0 * bf16_burst),
                    "traffic_source": "ncu dram__bytes_read+write per launch at the full shape on one GPU, profiles/r01d_ncu_summary.txt",
                    "per_mode_achieved_tops": ach})
    else:
        miss = N - n_obs
        pairs = K * (K + 1) / 2.0
        fl = {"gibbs": miss * (pairs + K) * 2.0, "vb": miss * (pairs + 2 * K) * 2.0}
        ach = {k: fl[k] / (prof[k]["stats_gram"] * 1e-3) / 1e12 for k in engs}
        mean_ach = sum(ach.values()) / len(ach)
        out.update({"bound": "tensor", "kernel": "k_stats_gram (fp64 DMMA, mma.sync.m8n8k4.f64)", "achieved": mean_ach,
                    "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": mean_ach / FP64_PEAK_TFLOPS, "traffic": None,
                    "peak_source": "fp64 DMMA peak measured by tools/microbench/fp64_pipes.cu"})
    # the HBM-bound kernel: streams R once per phase (`digits` digit-plane bytes per entry for the tcgen05 kernel,
    # 8 + 1/8 bytes for the fp64 kernel)
    bpe = float(digits) if e0.rx == "umma" else 8.125
    rx_ms = sum(prof[k]["stats_rx"] for k in engs) / len(engs)
    rx_gbs = N * bpe / (rx_ms * 1e-3) / 1e9
    out["hbm_kernel"] = {"bound": "hbm", "kernel": "k_rx_umma (tcgen05 digit planes)" if e0.rx == "umma" else "k_stats_rx (fp64 DMMA)",
                         "achieved": rx_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": rx_gbs / hbm_peak,
                         "algorithmic_bytes_per_launch": N * bpe, "peak_kind": hbm_kind + " (copy bandwidth)",
                         "traffic": RX_TRAFFIC_PER_ENTRY.get(digits, 0) * N if (e0.rx == "umma" and digits in RX_TRAFFIC_PER_ENTRY) else None,
                         "traffic_source": "ncu dram__bytes_read+write per launch / entries, profiles/r01d_ncu_summary.txt"}
    return out


def cpu_baseline(K, seed=0, budget_s=8.0, sizes=((1024, 512), (2048, 1024), (4096, 2048))):
    """Time the CPU oracle (numpy restatement of the reference's sweep, same per-column full-GEMM cost structure) on
    bounded samples of the same synthetic workload and extrapolate linearly in I*J to the full shape.  A size is
    started while less than budget_s seconds have been used: the last one (4096 x 2048, ~10-20 s for the two sweeps
    on 8-16 cores) brings the sample to the 10-30 s the bench contract asks for."""
    from oracle import bnmtf_oracle as orc
    try:
        import threadpoolctl
        cores = max(i.get("num_threads", 1) for i in threadpoolctl.threadpool_info()) if threadpoolctl.threadpool_info() else 1
    except Exception:
        cores = os.cpu_count() or 1
    per_elem = {}
    used = 0.0
    desc = []
    for (I, J) in sizes:
        if used > budget_s:
            break
        rng = np.random.RandomState(seed)
        U0, V0 = rng.exponential(1.0, size=(I, K)), rng.exponential(1.0, size=(J, K))
        R = U0 @ V0.T + rng.normal(size=(I, J))
        M = (rng.rand(I, J) >= 0.2).astype(float)
        pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}
        for mode in ("gibbs", "vb"):
            o = orc.OracleBNMF(R, M, K, pri, mode=mode, seed=seed)
            if mode == "vb":
                o.init_vb(1.0 / o.lambdaU, 1.0 / o.lambdaV)
            else:
                o.set_state(1.0 / o.lambdaU, 1.0 / o.lambdaV)
            t0 = time.time()
            o.sweep()
            dt = time.time() - t0
            used += dt
            per_elem[(mode, I * J)] = dt / (I * J)
        desc.append("%dx%d" % (I, J))
    return per_elem, cores, desc, used


def run_reference(args):
    I, J, K = args.rows, args.cols, args.K
    t0 = time.time()
    per_elem, cores, desc, used = cpu_baseline(K)
    biggest = max(n for (_, n) in per_elem)
    s_gibbs, s_vb = per_elem[("gibbs", biggest)] * I * J, per_elem[("vb", biggest)] * I * J
    value = 2.0 / (s_gibbs + s_vb)
    line = {"impl": "reference", "metric": "BNMF Gibbs+VB sweeps/sec at %dx%d K=%d" % (I, J, K), "value": value,
            "unit": "sweeps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BNMF Gibbs+VB sweep, %dx%d fp64, 20%% missing, K=%d" % (I, J, K),
                       "note": "oracle port of the reference's numpy sweep; the full shape does not fit host RAM, "
                               "value is the linear extrapolation in I*J from the largest sample"},
            "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": cores, "kind": "port",
                             "sample": "one Gibbs + one VB sweep at " + ", ".join(desc) + "; extrapolated linearly in I*J",
                             "seconds_per_sweep_extrapolated": {"gibbs": s_gibbs, "vb": s_vb}},
            "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows", type=int, default=65536)
    ap.add_argument("--cols", type=int, default=32768)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import torch
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from bnmtf_b200 import _lib, bnmf, engine
    I, J, K = args.rows, args.cols, args.K
    pri = {"alpha": 1.0, "beta": 1.0, "lambdaU": 0.1, "lambdaV": 0.1}

    if world > 1:
        from bnmtf_b200 import parallel
        return parallel.bench_sharded(args, rank, world, device, ClockSampler, measured_peaks, build_roofline)

    R, bits, n_obs = make_synthetic(I, J, K, device)
    ds = engine.Dataset.from_device(R, bits, I, J, n_obs=n_obs)
    torch.cuda.synchronize()
    models = {}
    for mode, cls in (("gibbs", bnmf.bnmf_gibbs_optimised), ("vb", bnmf.bnmf_vb_optimised)):
        m = cls.from_dataset(ds, K, pri, seed=1)
        m.initialise("exp")
        models[mode] = m
    engs = {k: m._engine() for k, m in models.items()}
    for m in models.values():
        m._push()
    # ---- device-resident timing: state stays in HBM, K sweeps of each sampler back to back -------------------
    for e in engs.values():
        e.alloc_trace(args.warmup + args.steps + 64)
    for _ in range(args.warmup):
        for e in engs.values():
            e.sweep()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in engs}
    launches0 = _lib.launch_count[0]
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for k, e in engs.items():
        ev[k][0].record()
        for _ in range(args.steps):
            e.sweep()
        ev[k][1].record()
    t_all1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count[0] - launches0
    ms = {k: ev[k][0].elapsed_time(ev[k][1]) for k in engs}
    total_ms = t_all0.elapsed_time(t_all1)
    value = 2.0 * args.steps / (total_ms / 1e3)
    mse = {k: float(e.scalars.cpu()[engine.S_MSE]) for k, e in engs.items()}

    # ---- per-kernel timings for the roofline block -----------------------------------------------------------
    prof = {k: e.profile_sweep(reps=2) for k, e in engs.items()}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- end to end through the public class API: host state in, host state out, every step -------------------
    e2e = None
    if not args.no_e2e:
        for m in models.values():
            m.run(1)
        torch.cuda.synchronize()
        t0 = time.time()
        for m in models.values():
            for _ in range(args.steps):
                m.run(1)
        torch.cuda.synchronize()
        dt = time.time() - t0
        fU, fV = I * K * 8, J * K * 8
        h2d = ((2 * (fU + fV)) + (5 * (fU + fV))) / 2.0        # gibbs: U,V,lambdaU,lambdaV; vb: exp,var,mu,tau,lambda
        d2h = ((2 * (fU + fV)) + (4 * (fU + fV))) / 2.0        # gibbs: state + the kept sample; vb: exp,var,mu,tau
        e2e = {"value": 2.0 * args.steps / dt, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h),
               "note": "model.run(1) per step: factor state DMA'd from the model's page-locked host arrays, one sweep, state + "
                       "trace read back into them; the priors come from pageable numpy; R itself (16 GiB) is resident like "
                       "a dataset"}

    # ---- roofline ------------------------------------------------------------------------------------------------
    roofline = build_roofline(engs, prof, I, J, K, n_obs, total_ms / 1e3 / (2.0 * args.steps))
    N = float(I) * J
    line = {"metric": "BNMF Gibbs+VB sweeps/sec at %dx%d K=%d" % (I, J, K), "value": value, "unit": "sweeps/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / (2.0 * args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BNMF Gibbs+VB sweep, %dx%d fp64, 20%% missing, K=%d" % (I, J, K),
                       "l2": "inputs (2 x %.1f GiB) far larger than L2" % (N * 8 / 2 ** 30),
                       "gibbs_sweeps_per_s": args.steps / (ms["gibbs"] / 1e3), "vb_sweeps_per_s": args.steps / (ms["vb"] / 1e3),
                       "train_mse_after": mse, "observed_fraction": n_obs / N},
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary()}
    if not args.no_cpu_baseline:
        per_elem, cores, desc, used = cpu_baseline(K)
        biggest = max(n for (_, n) in per_elem)
        s_g, s_v = per_elem[("gibbs", biggest)] * N, per_elem[("vb", biggest)] * N
        line["cpu_baseline"] = {"value": 2.0 / (s_g + s_v), "unit": "sweeps/s", "cores": cores, "kind": "port",
                                "sample": "one Gibbs + one VB sweep of the oracle at " + ", ".join(desc) +
                                          " (%.1f s of CPU); extrapolated linearly in I*J -- the full shape needs >62 GB" % used}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
