import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt
rng = np.random.default_rng(0)
n, order = 100002, int(os.environ.get("ORDER", "8"))
gfb = p.GammatoneFilterbank(samplerate=16000, order=order, startband=-11, endband=10, density=0.7)
x = torch.from_numpy(rng.standard_normal((16, n))).cuda()
for _ in range(int(os.environ.get("REPS", "1"))):
    gt.fos_bank(x, gfb._coef_rows, order)
torch.cuda.synchronize()
