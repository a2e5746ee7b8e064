"""One line per JSON result of tools/check_gram_umma.py (stdin -> stdout): shape, form, time, errors."""
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.strip()[:300]); continue
    print(d["case"], d["shape"], "vb", d["vb"], "pol", d["pol"], "sums", d["sums"], "nseg", d["nseg"], "tile", d["tile"], "umma_ms %.4f" % d.get("umma_ms", -1), "err %.2e" % d["umma_vs_ref"], "cnt", d["count_err"], "sv %.1e" % d.get("sv_vs_ref", 0), "sums %.1e" % d.get("sums_vs_ref", 0))
