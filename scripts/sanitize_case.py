"""Small shapes through every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt, splweighting as spl
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
rng = np.random.default_rng(0)
for J in ("1", "4", "32", "64"):
    os.environ["PFB_TMA_J"] = J
    bank = p.Bank(sos)
    for C, n in ((3, 9000), (37, 4608)):
        x = torch.from_numpy(rng.standard_normal((C, n))).cuda()
        p.bank_filter(bank, x)                                    # TMA f64 (wide for J >= 32)
        p.bank_filter(bank, x.float())                            # TMA f32
        p.bank_filter(bank, x.float(), arith64=True)              # TMA f32 signals, f64 arithmetic
        p.bank_levels(bank, x[:, :4608].contiguous(), 512)        # band levels
os.environ.pop("PFB_TMA_J")
bank = p.Bank(sos)
x = torch.from_numpy(rng.standard_normal((5, 3001))).cuda()       # unaligned rows: cp.async kernels
p.bank_filter(bank, x, mode="scan"); p.bank_filter(bank, x, mode="seq"); p.bank_filter(bank, x, mode="seq", exact=True)
p.bank_filter(bank, x[:, :3000].contiguous(), mode="scan", decimate=1)
ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
ofb.filter_decimated(rng.standard_normal((4096, 2)))
ofb.filter_mimo_c(rng.standard_normal((2048, 2)), weighting="A")
gfb = p.GammatoneFilterbank(samplerate=16000, order=4, startband=-14, endband=18, density=0.5)
for J in ("1", "8", "32"):
    os.environ["PFB_FOS_J"] = J
    y, st = gt.fos_bank(torch.from_numpy(rng.standard_normal((5, 8000 + 6))).cuda(), gfb._coef_rows, 4)
gfb.synthesize_batch(y)
torch.cuda.synchronize()
print("sanitize case done")
