"""Host-visible latency of one gammatone analyze call (tables cached after the first call)."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt
for order in (4, 8):
    gfb = p.GammatoneFilterbank(samplerate=16000, order=order, startband=-12, endband=12, density=0.25)
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((16, 160000))).cuda()
    ts = []
    for i in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y, st = gt.fos_bank(x, gfb._coef_rows, order)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("order %d, %d bands, 16 x 10 s: call latency ms first %.2f, then %s" % (order, len(gfb._coeffs), ts[0], " ".join("%.2f" % t for t in ts[1:])))
