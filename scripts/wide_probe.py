"""Wide (512-byte fragment) vs narrow output pass over several (channels, samples) shapes, float64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
for C, N in ((2, 2880000), (8, 480000), (8, 2880000), (16, 960000), (32, 2880000), (128, 1440000), (256, 480000), (1, 441000)):
    x = torch.randn(C, N, device="cuda", dtype=torch.float64) * 0.1
    y = torch.empty(C, 33, N, device="cuda", dtype=torch.float64)
    res = []
    for wide in ("0", "1", None):
        if wide is None: os.environ.pop("PFB_TMA_WIDE", None)
        else: os.environ["PFB_TMA_WIDE"] = wide
        bank = p.Bank(sos)
        for _ in range(2): p.bank_filter(bank, x, None, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): p.bank_filter(bank, x, None, out=y)
        e1.record(); torch.cuda.synchronize()
        info = bank.last_launch()
        res.append("%s: %.3f ms (J=%d%s)" % ({"0": "narrow", "1": "wide", None: "auto"}[wide], e0.elapsed_time(e1) / 5,
                                              round(N / max(info["chunk"], 1)), " wide" if info["aligned"] == 3 else ""))
    print("C=%4d N=%8d  %s" % (C, N, "   ".join(res)), flush=True)
    del x, y
