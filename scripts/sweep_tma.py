"""Sweep the launch shape of the TMA time-chunked path on the cfg-2 workload (64 ch x 60 s @ 48 kHz).
Prints ms per pass and the per-kernel event times for every (channel groups, chunk groups, bands per CTA)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
C = int(os.environ.get("C", 64)); N = int(os.environ.get("N", 2880000))
dts = {"f64": torch.float64, "f32": torch.float32}
CG = [int(v) for v in os.environ.get("CG", "1,4,8").split(",")]
GR = [int(v) for v in os.environ.get("GR", "3,4,6,8").split(",")]
NB = [int(v) for v in os.environ.get("NB", "7,11").split(",")]
NBA = [int(v) for v in os.environ.get("NBA", "0").split(",")]
for name in os.environ.get("DT", "f64").split(","):
    dt = dts[name]
    x = torch.randn(C, N, device="cuda", dtype=dt) * 0.1
    y = torch.empty(C, 33, N, device="cuda", dtype=dt)
    esz = 8 if name == "f64" else 4
    for cg, gr, nb, nba in itertools.product(CG, GR, NB, NBA):
        os.environ["PFB_TMA_CGROUPS"] = str(cg); os.environ["PFB_TMA_GROUPS"] = str(gr); os.environ["PFB_TMA_NB"] = str(nb)
        if nba: os.environ["PFB_TMA_NB_A"] = str(nba)
        else: os.environ.pop("PFB_TMA_NB_A", None)
        bank = p.Bank(sos)
        for it in range(2):
            p.bank_filter(bank, x, None, mode="scan", out=y)
        torch.cuda.synchronize()
        its = 5
        bank.timing_begin(its)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(its):
            p.bank_filter(bank, x, None, mode="scan", out=y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / its
        kt = bank.timing_end()
        bs = C * 33 * N / (ms * 1e-3)
        print("%s cg=%2d groups=%d nb=%2d nba=%2d: %7.3f ms  %.3e bs/s  step_frac %.3f | A %.2f carry %.2f B %.2f total %.2f" % (
            name, cg, gr, nb, nba, ms, bs, bs * (esz + esz / 33) / 6551e9, kt["prepass"], kt["carry"], kt["output"], kt["total"]), flush=True)
        del bank
