"""Per-phase CUDA-event times of one large tri-factor sweep (bench.py --workload nmtf shape): which of the calls of
BNMTFEngine._sweep_eager the sweep time goes to.  Usage: python tools/nmtf_phase_times.py [rows cols K] > out.json"""
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import bench
    from bench_replicas import priors3
    from bnmtf_b200 import bnmtf, engine
    I, J, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (65536, 32768, 10)
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    R, bits, n_obs = bench.make_synthetic(I, J, K, device)
    ds = engine.Dataset.from_device(R, bits, I, J, n_obs=n_obs)
    out = {"shape": [I, J, K]}
    for mode, cls in (("gibbs", bnmtf.bnmtf_gibbs_optimised), ("vb", bnmtf.bnmtf_vb_optimised)):
        m = cls.from_dataset(ds, K, K, priors3(), seed=1)
        np.random.seed(1), random.seed(1)
        m.initialise("random", "random")
        m._push()
        e = m._engine()
        e.alloc_trace(64)
        engine.thread_flags.no_graph = True
        for _ in range(2):
            e.sweep()
        steps = (("stats_rows", e.stats_rows), ("phase_S", e.phase_S), ("phase_F", e.phase_F), ("stats_cols", e.stats_cols),
                 ("phase_G", e.phase_G)) if e.vb else \
                (("stats_rows", e.stats_rows), ("phase_F", e.phase_F), ("phase_S", e.phase_S), ("stats_cols", e.stats_cols),
                 ("phase_G", e.phase_G))
        steps += ((("vb_extra", e.vb_extra), ("vb_terms", e.vb_terms)) if e.vb else ())
        steps += (("phase_S_reduction_only", lambda: e.phase_S(order=[], apply=False, use_iter=False)),)
        steps += (("metrics", e.metrics_from_stats if e.metrics_mode == "stats" else e.metrics), ("finish", e.finish))
        acc = {n: 0.0 for n, _ in steps}
        reps = 3
        for _ in range(reps):
            for n, fn in steps:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                fn()
                b.record()
                b.synchronize()
                acc[n] += a.elapsed_time(b) / reps
        acc["sum"] = sum(v for k, v in acc.items() if k != "phase_S_reduction_only")
        out[mode] = {k: round(v, 4) for k, v in acc.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
