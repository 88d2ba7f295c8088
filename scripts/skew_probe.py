"""cfg-2 float64 headline step with start offsets between the band warps of an SM sub-partition (PFB_TMA_SKEW clocks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
bank = p.Bank(sos)
C, N, B = 64, 2880000, 33
x = torch.randn(C, N, device="cuda", dtype=torch.float64) * 0.1
y = torch.empty(C, B, N, device="cuda", dtype=torch.float64)
for skew in [0, 500, 1000, 1500, 2000, 3000, 4000, 6000, 0]:
    os.environ["PFB_TMA_SKEW"] = str(skew)
    res = []
    for it in range(7):
        bank.timing_begin(1)
        p.bank_filter(bank, x, None, out=y)
        torch.cuda.synchronize()
        km = bank.timing_end()
        if it >= 2:
            res.append(km["output"])
    print("skew %5d: output pass ms min %.3f median %.3f  (%s)" % (skew, min(res), sorted(res)[len(res) // 2], " ".join("%.2f" % r for r in res)), flush=True)
