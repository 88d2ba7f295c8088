"""Error of the time-chunked gammatone path vs cascade order and chunk count (why orders above four are not chunked)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt
from oracle import oracle
rng = np.random.default_rng(0)
C, n = 16, 100002
x = rng.standard_normal((C, n))
for order in (2, 4, 5, 6, 8):
    gfb = p.GammatoneFilterbank(samplerate=16000, order=order, startband=-8, endband=8, density=0.7)
    for J in ("1", "2", "8", "32"):
        os.environ["PFB_FOS_J"] = J
        y, st = gt.fos_bank(torch.from_numpy(x).cuda(), gfb._coef_rows, order)
        errs = []
        for i in range(len(gfb._coeffs)):
            b, a = gfb._coeffs[i]
            y0, _ = oracle.fosfilter(b, a, order, x[0])
            errs.append(float(np.linalg.norm(y[0, i].cpu().numpy() - y0) / np.linalg.norm(y0)))
        print("order %d  J=%2s  worst band error %.2e (band %d, |c| = %.5f)" % (order, J, max(errs), int(np.argmax(errs)),
              abs(complex(gfb._coef_rows[int(np.argmax(errs))][1], gfb._coef_rows[int(np.argmax(errs))][2]))))
