import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import pyfilterbank_b200 as pfb
from pyfilterbank_b200 import _lib
g = np.load('/root/repo/tests/golden/designs.npz')
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
C, n, B, K = 64, 2880000, 33, 4
bank = pfb.Bank(sos)
xh = torch.randn((C, n), dtype=torch.float64).pin_memory()
yh = torch.empty((C, B, n), dtype=torch.float64, pin_memory=True)
sth = np.zeros((C, B, K, 2))
def run():
    sth[:] = 0
    _lib.check(_lib.lib().pfb_bank_filter_host(bank._h, 1, xh.data_ptr(), n, yh.data_ptr(), n, sth.ctypes.data, n, C, 0))
for mb in sys.argv[1:]:
    os.environ["PFB_HOST_BLOCK_MB"] = mb
    run()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
    print("PFB_HOST_BLOCK_MB=%s: %s ms  -> %.2f GB/s D2H" % (mb, ["%.0f" % (t * 1e3) for t in ts], 48.66 / min(ts)))
