"""Scratch timing of the cfg-2 kernel (64 ch x 60 s @ 48 kHz, 1/3-octave bank)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
bank = p.get_bank(sos)
C = int(os.environ.get("C", 64)); N = int(os.environ.get("N", 2880000))
for dt in (torch.float64, torch.float32):
    x = torch.randn(C, N, device="cuda", dtype=dt) * 0.1
    y = torch.empty(C, 33, N, device="cuda", dtype=dt)
    for mode, kw in (("scan", {}), ("seq", {})) + ((("scan", {"arith64": True}),) if dt == torch.float32 else ()):
        st = None
        for it in range(3):
            st = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _, st = p.bank_filter(bank, x, None, mode=mode, out=y, **kw); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        bs = C * 33 * N / (ms * 1e-3)
        esz = 8 if dt == torch.float64 else 4
        print(dt, mode, kw, "ms %.2f  band-samples/s %.3e  GB/s %.0f  frac %.3f" % (ms, bs, bs * (esz + esz / 33) / 1e9, bs * (esz + esz / 33) / 6551e9), bank.last_launch(), flush=True)
