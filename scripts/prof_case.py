"""Small fixed case for ncu: 1/3-octave bank, C channels x N samples, one scan-mode launch per dtype."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
bank = p.get_bank(sos)
C = int(os.environ.get("C", 64)); N = int(os.environ.get("N", 360000))
dts = {"f64": torch.float64, "f32": torch.float32}
for name in os.environ.get("DT", "f64").split(","):
    x = torch.randn(C, N, device="cuda", dtype=dts[name]) * 0.1
    y = torch.empty(C, 33, N, device="cuda", dtype=dts[name])
    for it in range(int(os.environ.get("IT", 2))):
        p.bank_filter(bank, x, None, mode=os.environ.get("MODE", "scan"), out=y)
    torch.cuda.synchronize()
print("ok")
