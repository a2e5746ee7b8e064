"""One line per bench.py JSON line (stdin -> stdout): value, sustained, per-kernel times alone and inside the sweep."""
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    km = d.get("roofline", {}).get("kernel_ms", {})
    g = km.get("gibbs", {})
    print("value %.1f  sustained %s  ms/step %.3f | gibbs: rx %.2f gram %.2f solve %.2f  in-sweep rx %.2f gram %.2f phase %.2f | mse %s" % (
        d["value"], d["config"].get("sustained", {}).get("sweeps_per_s") if isinstance(d["config"].get("sustained"), dict) else d["config"].get("sustained"),
        d["ms_per_step"], g.get("stats_rx", 0), g.get("stats_gram", 0), g.get("row_solve", 0), g.get("stats_rx_in_sweep", 0), g.get("stats_gram_in_sweep", 0),
        g.get("stats_phase_in_sweep", 0), d["config"].get("state_after_timed_region", {}).get("gibbs", {}).get("train_MSE") if isinstance(d["config"].get("state_after_timed_region"), dict) else None))
