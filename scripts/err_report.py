"""Per-band relative L2 error of every GPU arithmetic variant against the CPU oracle (float64 reference
arithmetic), 1/3-octave bank @48 kHz: the time-chunked scan's re-association error is reported separately
from the FMA-contraction error of the sequential kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from oracle import oracle
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
rng = np.random.default_rng(5); C, n = 2, int(os.environ.get("N", 960000))
x = rng.standard_normal((C, n)); y0, s0 = oracle.sos_bank(x, sos)
xt = torch.from_numpy(x).cuda()
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
ENV = ("PFB_NO_SCALED", "PFB_NO_TMA", "PFB_WARM_TOL", "PFB_NO_HFORM", "PFB_TMA_J")
CASES = (("sequential, exact (separately rounded)", {}, dict(mode="seq", exact=True)),
         ("sequential, FMA", {}, dict(mode="seq")),
         ("TMA, 1 chunk (sequential, FMA, gain-normalised)", {"PFB_TMA_J": "1"}, dict(mode="scan")),
         ("TMA scan J=64, GEMM pre-pass, gain-normalised (default arithmetic)", {"PFB_TMA_J": "64"}, dict(mode="scan")),
         ("TMA scan J=256, GEMM pre-pass, gain-normalised", {"PFB_TMA_J": "256"}, dict(mode="scan")),
         ("TMA scan J=64, recurrence pre-pass", {"PFB_TMA_J": "64", "PFB_NO_HFORM": "1"}, dict(mode="scan")),
         ("TMA scan J=64, plain (not gain-normalised)", {"PFB_TMA_J": "64", "PFB_NO_SCALED": "1"}, dict(mode="scan")),
         ("TMA scan J=64, untruncated pre-pass (tol 0)", {"PFB_TMA_J": "64", "PFB_WARM_TOL": "0"}, dict(mode="scan")),
         ("cp.async scan kernel", {"PFB_NO_TMA": "1"}, dict(mode="scan")))
ref_seq = None
for name, env, kw in CASES:
    for k in ENV: os.environ.pop(k, None)
    os.environ.update(env)
    bank = p.Bank(sos)
    y1, s1 = p.bank_filter(bank, xt, **kw); y1 = y1.cpu().numpy()
    errs = [rel(y1[:, b], y0[:, b]) for b in range(33)]
    line = "%-68s worst %.2e (band %2d)  median %.2e  states %.2e" % (name, max(errs), int(np.argmax(errs)), float(np.median(errs)), rel(s1.cpu().numpy(), s0))
    if name.startswith("sequential, FMA"): ref_seq = y1
    elif ref_seq is not None and "scan" in name:
        d = [rel(y1[:, b], ref_seq[:, b]) for b in range(33)]
        line += "   | vs sequential FMA kernel (re-association only): worst %.2e" % max(d)
    print(line, flush=True)
