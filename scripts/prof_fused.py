"""Small fixed cases for ncu captures: cfg5 (weighting warp), cfg3 (decimating band), cfg2 (headline), cfg4 (gammatone),
levels (band-level epilogue), tf (stand-alone weighting kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import splweighting as spl
case = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
if case.startswith("cfg5"):
    dtype = torch.float32 if case.endswith("f32") else torch.float64
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
    bank = p.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(33, 4, 6))
    b, a = spl.a_weighting_coeffs_design(48000)
    C, n = 1024, 48000
    x = (torch.randn(C, n, device="cuda", dtype=torch.float64) * 0.1).to(dtype)
    y = torch.empty(C, 33, n, device="cuda", dtype=dtype)
    for _ in range(3):
        p.bank_filter_weighted(bank, x, b, a, out=y)
elif case == "cfg4":
    from pyfilterbank_b200 import gammatone as gt
    gfb = gt.GammatoneFilterbank(samplerate=16000, order=4, startband=-14, endband=18, density=0.5)
    x = torch.randn(512, 80000, device="cuda", dtype=torch.float64) * 0.1
    for _ in range(2):
        gt.fos_bank(x, gfb._coef_rows, 4)
elif case == "levels":
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
    bank = p.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(33, 4, 6))
    x = torch.randn(64, 1440000, device="cuda", dtype=torch.float64) * 0.1
    for _ in range(2):
        p.bank_levels(bank, x, 4800)
elif case == "tf":
    b, a = spl.a_weighting_coeffs_design(48000)
    x = torch.randn(1024, 48000, device="cuda", dtype=torch.float64) * 0.1
    for _ in range(2):
        spl.tf_filter(b, a, x)
elif case in ("cfg2auto", "cfg2fd"):      # float32 signals: per-band arithmetic / float64 arithmetic everywhere
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
    s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
    bank = p.get_bank(sos)
    x = torch.randn(64, 1440000, device="cuda", dtype=torch.float32) * 0.1
    y = torch.empty(64, 33, 1440000, device="cuda", dtype=torch.float32)
    for _ in range(2):
        p.bank_filter(bank, x, None, out=y, arith64="auto" if case == "cfg2auto" else True)
elif case == "cfg3":
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51)
    x = (torch.randn(1024, 48000, device="cuda", dtype=torch.float64) * 0.1).t()
    for _ in range(2):
        ofb.filter_decimated(x)
else:
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
    s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
    bank = p.get_bank(sos)
    x = torch.randn(64, 1440000, device="cuda", dtype=torch.float64) * 0.1
    y = torch.empty(64, 33, 1440000, device="cuda", dtype=torch.float64)
    for _ in range(2):
        p.bank_filter(bank, x, None, out=y)
torch.cuda.synchronize()
print("ok")
