import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt
from oracle import oracle
rng = np.random.default_rng(0)
n, order = 100002, 8
gfb = p.GammatoneFilterbank(samplerate=16000, order=order, startband=-11, endband=10, density=0.7)
B = len(gfb._coeffs)
os.environ["PFB_FOS_J"] = "128"
for C in (1, 2, 16):
    x = rng.standard_normal((C, n))
    xt = torch.from_numpy(x).cuda()
    ys = [gt.fos_bank(xt, gfb._coef_rows, order)[0] for _ in range(4)]
    same = [bool(torch.equal(ys[0], y)) for y in ys[1:]]
    ref = np.stack([oracle.fosfilter(*gfb._coeffs[i], order, x[0])[0] for i in range(B)])
    err = np.linalg.norm(ys[0][0].cpu().numpy() - ref, axis=1) / np.linalg.norm(ref, axis=1)
    print("C=%d: runs bitwise equal to run 0: %s; channel 0 worst band error %.1e (band %d)" % (C, same, err.max(), int(err.argmax())))
