"""cfg 5 (fused weighted bank, 1024 channels) against the row pitch of input and output: is the run-to-run spread of this
case (14.4 / 17.8 / 47 ms on different visits) a matter of where rows fall in HBM?"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import splweighting as spl
ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
bank = p.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(33, 4, 6))
b, a = spl.a_weighting_coeffs_design(48000)
C, n = 1024, 240000
for dtype in (torch.float64, torch.float32):
    for pitch in (240000, 240000 + 32, 240000 + 512, 240000 + 2048, 262144):
        for rep in range(2):
            xb = (torch.randn(C, pitch, device="cuda", dtype=torch.float64) * 0.1).to(dtype)
            yb = torch.empty(C, 33, pitch, device="cuda", dtype=dtype)
            x, y = xb[:, :n], yb[:, :, :n]
            for _ in range(1):
                p.bank_filter_weighted(bank, x, b, a, out=y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                p.bank_filter_weighted(bank, x, b, a, out=y)
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps({"dtype": str(dtype)[6:], "row_pitch_samples": pitch, "rep": rep, "ms": round(e0.elapsed_time(e1) / 3, 2),
                              "y_ptr_mod_2MB": yb.data_ptr() % (2 << 20), "path": bank.last_launch()["aligned"]}), flush=True)
            del xb, yb, x, y
            torch.cuda.empty_cache()
