"""pfb_bank_filter_host on float32 signals (32 channels x 60 s, pinned buffers): float64 arithmetic against per-band
arithmetic (PFB_ARITH_AUTO), alternating, on one box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as pfb
from pyfilterbank_b200 import _lib
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
C, n, B, K = 32, 2880000, 33, 4
bank = pfb.Bank(sos)
xh = (torch.randn((C, n), dtype=torch.float32) * 0.1).pin_memory()
yh = torch.empty((C, B, n), dtype=torch.float32, pin_memory=True)
sth = np.zeros((C, B, K, 2))
def run(flags):
    sth[:] = 0
    _lib.check(_lib.lib().pfb_bank_filter_host(bank._h, 0, xh.data_ptr(), n, yh.data_ptr(), n, sth.ctypes.data, n, C, flags))
for rep in range(3):
    for name, flags in (("arith_f64", _lib.ARITH_F64), ("arith_auto", _lib.ARITH_F64 | _lib.ARITH_AUTO), ("arith_f32", 0)):
        if name == "arith_f32":
            sth = np.zeros((C, B, K, 2), dtype=np.float32)
        else:
            sth = np.zeros((C, B, K, 2))
        run(flags)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); run(flags); ts.append(time.perf_counter() - t0)
        print("%-10s %s ms -> %.1f GB/s D2H, launch %s" % (name, ["%.0f" % (t * 1e3) for t in ts], C * B * n * 4 / min(ts) / 1e9, bank.last_launch()), flush=True)
