"""Random (channels, samples, bands, sections, dtype) shapes through the automatic planner against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200.butterworth import butter_sos
from oracle import oracle
rng = np.random.default_rng(int(os.environ.get("SEED", 0)))
def rel(a, b):
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))
def random_bank(B, K):
    rows = []
    for b in range(B):
        lo = 10 ** rng.uniform(-3.3, -0.7)
        hi = min(0.49, lo * 2 ** rng.uniform(0.2, 1.0))
        kind = rng.choice(["bandpass", "lowpass", "highpass"])
        if kind == "bandpass" and K % 2 == 0:
            sos = butter_sos("bandpass", K, lo, hi)             # K sections
        elif kind == "lowpass":
            sos = butter_sos("lowpass", 2 * K, hi)
        else:
            sos = butter_sos("highpass", 2 * K, lo)
        rows.append(sos[:K])
    return np.stack(rows)
t_end = time.time() + float(os.environ.get("SECONDS", 60))
cases = worst = 0
modes = {}
while time.time() < t_end:
    K = int(rng.choice([1, 2, 3, 4, 6, 8])); B = int(rng.integers(1, 40))
    C = int(rng.choice([1, 2, 3, 5, 8, 17, 32, 33, 64, 70])); n = int(rng.choice([1, 17, 130, 1000, 4096, 5000, 20000, 65536, 100003, 200000]))
    if C * B * n > 4e8: n = max(1, int(4e8 // (C * B)))
    aligned = rng.random() < 0.7
    if aligned: n = max(4, n // 4 * 4)
    sos = random_bank(B, K)
    x = rng.standard_normal((C, n))
    st0 = 0.01 * rng.standard_normal((C, B, K, 2))
    y0, s0 = oracle.sos_bank(x, sos, st0)
    bank = p.Bank(sos)
    y1, s1 = p.bank_filter(bank, torch.from_numpy(x).cuda(), torch.from_numpy(st0.copy()).cuda())
    info = bank.last_launch()
    y1 = y1.cpu().numpy()
    err = max(rel(y1[:, b], y0[:, b]) for b in range(B))
    key = (info["mode"], info["aligned"]); modes[key] = modes.get(key, 0) + 1
    if not (err <= 1e-9 and rel(s1.cpu().numpy(), s0) <= 1e-7 and np.isfinite(y1).all()):
        print("FAIL K=%d B=%d C=%d n=%d err=%.2e states=%.2e %s" % (K, B, C, n, err, rel(s1.cpu().numpy(), s0), info)); sys.exit(1)
    worst = max(worst, err); cases += 1
    if rng.random() < 0.3:   # float32 signals with float64 arithmetic on the same shape
        xf = x.astype(np.float32); yd, _ = oracle.sos_bank(xf.astype(np.float64), sos, st0)
        y2, _ = p.bank_filter(bank, torch.from_numpy(xf).cuda(), torch.from_numpy(st0.copy()).cuda(), arith64=True)
        e2 = max(rel(y2.cpu().numpy()[:, b], yd[:, b]) for b in range(B))
        if not e2 <= 1e-6:
            print("FAIL fd K=%d B=%d C=%d n=%d err=%.2e" % (K, B, C, n, e2)); sys.exit(1)
print("fuzz ok: %d cases, worst band error %.2e, launch kinds (mode, path) %s" % (cases, worst, modes))
