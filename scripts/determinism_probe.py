"""Run-to-run bitwise determinism of the chunked paths (same inputs, 4 runs each)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt, sosfiltering as sf
rng = np.random.default_rng(0)

def check(name, fn, reps=4):
    ys = [fn() for _ in range(reps)]
    torch.cuda.synchronize()
    same = [bool(torch.equal(ys[0], y)) for y in ys[1:]]
    worst = max(float((ys[0] - y).abs().max()) for y in ys[1:])
    print("%-52s equal to run 0: %s  max |diff| %.2e" % (name, same, worst), flush=True)

ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
for C, n in (() if "gammatone" in sys.argv else ((64, 480000), (16, 2880000), (3, 1000001))):
    for dt in (torch.float64, torch.float32):
        x = torch.from_numpy(rng.standard_normal((C, n))).cuda().to(dt)
        for env in ({}, {"PFB_NO_HFORM": "1"}):
            os.environ.update(env)
            ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
            check("sos bank C=%d n=%d %s %s" % (C, n, str(dt)[6:], env or ""), lambda: ofb.filter_mimo_c(x)[0])
            for k in env: del os.environ[k]
        del x
for order in (1, 2, 4, 8):
    gfb = p.GammatoneFilterbank(samplerate=16000, order=order, startband=-11, endband=10, density=0.7)
    for C in (2, 16, 64):
        x = torch.from_numpy(rng.standard_normal((C, 100002))).cuda()
        for J in ("0", "32", "128"):
            os.environ["PFB_FOS_J"] = J
            check("gammatone order %d C=%d J=%s" % (order, C, J), lambda: gt.fos_bank(x, gfb._coef_rows, order)[0])
