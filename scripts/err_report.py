"""Per-band relative L2 error of the GPU paths vs the CPU oracle (float64), 1/3-octave bank @48 kHz."""
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
for name, env, kw in (("seq fma", {}, dict(mode="seq")), ("scan tma scaled", {}, dict(mode="scan")),
                      ("scan tma plain", {"PFB_NO_SCALED": "1"}, dict(mode="scan")),
                      ("scan cp.async", {"PFB_NO_TMA": "1"}, dict(mode="scan")),
                      ("scan tma exact prepass", {"PFB_WARM_TOL": "0"}, dict(mode="scan"))):
    for k in ("PFB_NO_SCALED", "PFB_NO_TMA", "PFB_WARM_TOL"): os.environ.pop(k, None)
    os.environ.update(env)
    bank = p.Bank(sos)
    y1, s1 = p.bank_filter(bank, xt, **kw); y1 = y1.cpu().numpy()
    errs = [rel(y1[:, b], y0[:, b]) for b in range(33)]
    print("%-24s worst %.2e (band %d)  median %.2e  states %.2e  %s" % (name, max(errs), int(np.argmax(errs)), float(np.median(errs)), rel(s1.cpu().numpy(), s0), bank.last_launch()))
