"""Print (not assert) where the VB tri-factorisation on the device stands against the reference's golden trajectory:
(1) the reference itself re-run on THIS host (baseline/_ref) against the goldens made in the build container -- the
reference's own host-to-host reproducibility; (2) the device trajectory, error per sweep; (3) one device sweep from
the oracle's state, error per quantity with the truncation point x = -mu sqrt(tau) of the worst entry."""
import contextlib, io, os, random, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def G(name):
    return dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-2 * max(1e-300, float(np.max(np.abs(b)))))))


def pri(g):
    lam = float(g["lambda"])
    return {"alpha": 1.0, "beta": 1.0, "lambdaF": lam, "lambdaS": lam, "lambdaG": lam}


def reference_rerun(name):
    import bench
    ref = bench.load_reference()
    if ref is None:
        print("no baseline/_ref: skipping the reference re-run")
        return
    g = G(name)
    K, L, its = int(g["K"]), int(g["L"]), int(g["its"])
    np.random.seed(0), random.seed(0)
    m = ref.bnmtf_vb_optimised(g["R"], g["M"], K, L, pri(g))
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialise("random", "random")
    print("%s reference re-run on this host: init muS identical: %s" % (name, np.array_equal(m.muS, g["init_muS"])))
    mse, tau = [], []
    for it in range(its):
        with contextlib.redirect_stdout(io.StringIO()):
            m.run(1)
        mse.append(m.predict(g["M"])["MSE"]), tau.append(m.exptau)
        if it in (0, 1, 4, 9, 19, its - 1):
            print("   sweep %2d: MSE %.1e exptau %.1e" % (it + 1, abs(mse[-1] / g["trace_MSE"][it] - 1), abs(tau[-1] / g["trace_exptau"][it] - 1)))
    print("   final: expF %.1e expS %.1e expG %.1e varS %.1e" % tuple(rel(getattr(m, k), g["final_" + k]) for k in ("expF", "expS", "expG", "varS")))


def device(name):
    import bnmtf_b200
    from oracle import bnmtf_oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_bnmtf_gpu import vb_from_golden, _set_vb_state
    g = G(name)
    K, L, its = int(g["K"]), int(g["L"]), int(g["its"])
    m = vb_from_golden(g)
    eng = m._push()
    eng.alloc_trace(its)
    for it in range(its):
        order = {"S": [int(k) * L + int(l) for k, l in g["order_S"][it]], "F": [int(x) for x in g["order_F"][it]],
                 "G": [int(x) for x in g["order_G"][it]]}
        eng.sweep(order=order)
    tr = eng.trace.cpu().numpy()[:its]
    print("%s device trajectory vs golden:" % name)
    for it in (0, 1, 4, 9, 19, its - 1):
        if it < its:
            print("   sweep %2d: MSE %.1e exptau %.1e" % (it + 1, abs(tr[it, 1] / g["trace_MSE"][it] - 1), abs(tr[it, 0] / g["trace_exptau"][it] - 1)))
    m._pull(eng)
    print("   final: expF %.1e expS %.1e expG %.1e varS %.1e" % tuple(rel(getattr(m, k), g["final_" + k]) for k in ("expF", "expS", "expG", "varS")))
    # one sweep from the oracle's state
    m = vb_from_golden(g)
    o = orc.OracleBNMTF(g["R"], g["M"], K, L, pri(g), mode="vb")
    o.init_vb(g["init_muF"], g["init_muS"], g["init_muG"], {"F": g["init_tauF"], "S": g["init_tauS"], "G": g["init_tauG"]})
    for it in range(min(4, its)):
        _set_vb_state(m, o)
        eng = m._push()
        eng.alloc_trace(1)
        oS = [tuple(int(v) for v in x) for x in g["order_S"][it]]
        order = {"S": oS, "F": [int(x) for x in g["order_F"][it]], "G": [int(x) for x in g["order_G"][it]]}
        o.sweep(order=order)
        eng.sweep(order={"S": [k * L + l for k, l in oS], "F": order["F"], "G": order["G"]})
        m._pull(eng)
        line = "   one sweep from the oracle state, it %d:" % it
        for k in "SFG":
            for q in ("mu", "tau", "exp", "var"):
                a, b = getattr(m, q + k), getattr(o, (q + k) if q != "exp" else k)
                e = np.abs(a - b) / (np.abs(b) + 1e-2 * np.abs(b).max())
                idx = np.unravel_index(np.argmax(e), e.shape)
                x = -getattr(o, "mu" + k)[idx] * np.sqrt(getattr(o, "tau" + k)[idx])
                line += " %s%s %.0e(x=%.0f)" % (q, k, e.max(), x)
        print(line)


if __name__ == "__main__":
    for name in ("toy_bnmtf_vb", "gdsc_bnmtf_vb"):
        reference_rerun(name)
        device(name)
