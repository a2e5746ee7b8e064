#!/usr/bin/env python
"""GPU check + timing of the tcgen05 R.X kernel (bnmtf_stats_rx_umma_f64) against the fp64 DMMA kernel
(bnmtf_stats_rx_f64) and a torch fp64 matmul.  Development tool, not part of the product path.

    python tools/check_rx_umma.py            # every case, each in its own process with a timeout
    python tools/check_rx_umma.py <case#>
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAXCTAS = int(os.environ.get("MAXCTAS", "0"))
sys.path.insert(0, ROOT)

# rows, cols, K, nseg, kind of data, timing reps
CASES = [
    (100, 80, 10, 1, "exp", 0),
    (129, 65, 5, 1, "signedR", 0),
    (300, 1000, 20, 2, "exp", 0),
    (300, 5000, 20, 1, "signedR", 0),          # > 4096 columns: two accumulator periods
    (1000, 9000, 32, 2, "wide", 0),            # rows with very different magnitudes
    (700, 3000, 17, 3, "negX", 0),             # negative factor entry -> gated fp64 fallback
    (2000, 20000, 20, 1, "exp", 0),            # 5 periods in one segment
    (65536, 32768, 20, 3, "exp", 3),
    (32768, 65536, 20, 6, "exp", 3),
    (65536, 32768, 20, 1, "exp", 3),
    (65536, 32768, 16, 3, "exp", 3),
]


def run_case(idx):
    import torch
    from bnmtf_b200 import _lib
    from bnmtf_b200.engine import _ptr, _stream, ld_for, kp_for
    rows, cols, K, nseg, kind, reps = CASES[idx]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(99 + idx)
    ld = ld_for(cols)
    KP = kp_for(K)
    X = -torch.log(torch.rand((cols, K), dtype=torch.float64, device=dev, generator=g))
    if kind == "negX":
        X[5, 3] = -0.25
    n_alloc = ld + 8
    Xp = torch.zeros((n_alloc, KP), dtype=torch.float64, device=dev)
    _lib.call("bnmtf_pad_factor_f64", _ptr(X), 0, cols, K, n_alloc, _ptr(Xp), 0, _stream())
    R = torch.zeros((rows, ld), dtype=torch.float64, device=dev)
    bits = torch.zeros((rows, ld // 32), dtype=torch.int32, device=dev)
    chunk = 4096
    U0 = -torch.log(torch.rand((rows, K), dtype=torch.float64, device=dev, generator=g))
    ref = None
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        Rt = U0[r0:r1] @ X.abs().T + torch.randn((r1 - r0, cols), dtype=torch.float64, device=dev, generator=g)
        if kind == "signedR":
            Rt = Rt - Rt.mean()
        if kind == "wide":
            Rt = Rt * torch.exp(30 * torch.randn((r1 - r0, 1), dtype=torch.float64, device=dev, generator=g))
        Mt = (torch.rand((r1 - r0, cols), dtype=torch.float32, device=dev, generator=g) < 0.8).to(torch.float64)
        _lib.call("bnmtf_pack_dataset_f64", _ptr(Rt), _ptr(Mt), r1 - r0, cols, ld, R[r0:].data_ptr(), bits[r0:].data_ptr(), _stream())
        torch.cuda.synchronize()
        if r0 == 0:
            ref = (Rt * Mt) @ X
            refscale = (Rt * Mt).abs() @ X.abs()
        del Rt, Mt
    nr = ref.shape[0]
    nbytes = _lib.call("bnmtf_rx_planes_bytes", rows, ld)
    pbuf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    pptr = (pbuf.data_ptr() + 1023) // 1024 * 1024
    rscale = torch.empty(rows, dtype=torch.float64, device=dev)
    rexp = torch.empty(rows, dtype=torch.int32, device=dev)
    t0 = time.time()
    _lib.call("bnmtf_rx_planes_pack_f64", _ptr(R), _ptr(bits), rows, ld, pptr, _ptr(rscale), _ptr(rexp), 0, _stream())
    torch.cuda.synchronize()
    out = {"case": idx, "shape": [rows, cols, K], "nseg": nseg, "kind": kind, "pack_s": round(time.time() - t0, 3)}
    wsb = _lib.call("bnmtf_rx_umma_workspace_bytes", K, ld)
    ws = torch.zeros(wsb + 1024, dtype=torch.uint8, device=dev)
    wsp = (ws.data_ptr() + 1023) // 1024 * 1024
    O1 = torch.full((nseg * rows, KP), float("nan"), dtype=torch.float64, device=dev)
    O0 = torch.full((nseg * rows, KP), float("nan"), dtype=torch.float64, device=dev)

    def umma():
        _lib.call("bnmtf_stats_rx_umma_f64", pptr, _ptr(rscale), _ptr(R), _ptr(bits), rows, ld, cols, _ptr(Xp), K, nseg,
                  MAXCTAS, _ptr(O1), wsp, wsb, _stream())

    def dmma():
        _lib.call("bnmtf_stats_rx_f64", _ptr(R), _ptr(bits), rows, ld, _ptr(Xp), K, nseg, _ptr(O0), _stream())
    umma()
    dmma()
    torch.cuda.synchronize()
    got1 = O1.view(nseg, rows, KP).sum(0)[:nr, :K]
    got0 = O0.view(nseg, rows, KP).sum(0)[:nr, :K]
    # error relative to sum |r||x| (the natural scale of the rounding error of either method)
    out["umma_vs_ref"] = float(((got1 - ref).abs() / refscale.clamp_min(1e-300)).max())
    out["dmma_vs_ref"] = float(((got0 - ref).abs() / refscale.clamp_min(1e-300)).max())
    out["umma_vs_dmma"] = float(((got1 - got0).abs() / refscale.clamp_min(1e-300)).max())
    if reps:
        for name, fn in (("umma_ms", umma), ("dmma_ms", dmma)):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            out[name] = e0.elapsed_time(e1) / reps
    print(json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
        return
    for i in range(len(CASES)):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True, timeout=240)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            if r.returncode != 0:
                tail += " | rc=%d %s" % (r.returncode, r.stderr.strip()[-400:].replace("\n", " / "))
        except subprocess.TimeoutExpired:
            tail = json.dumps({"case": i, "timeout": True})
        print("%s   [%.0fs]" % (tail, time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
