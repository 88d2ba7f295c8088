"""Targeted parity sweep of the wide float64 output pass (per-lane bulk copies of 512-byte row fragments): shapes that the
planner sends there (>= 25 channels, >= 9 bands, K <= 4, long signals), ragged ends, carried states, guard bands."""
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
        lo = 10 ** rng.uniform(-3.0, -0.7); hi = min(0.49, lo * 2 ** rng.uniform(0.2, 1.0))
        sos = butter_sos("bandpass", K, lo, hi) if K % 2 == 0 else butter_sos("lowpass", 2 * K, hi)
        rows.append(sos[:K])
    return np.stack(rows)
t_end = time.time() + float(os.environ.get("SECONDS", 60)); cases = wide = 0; worst = 0.0
while time.time() < t_end:
    K = int(rng.choice([1, 2, 3, 4])); B = int(rng.choice([9, 11, 12, 23, 33])); C = int(rng.choice([25, 32, 33, 64, 70]))
    n = int(rng.choice([131072, 200000, 262144 + 4 * int(rng.integers(0, 64)), 300000]))
    if C * B * n > 6e8: n = int(6e8 // (C * B)) // 4 * 4
    sos = random_bank(B, K); x = rng.standard_normal((C, n)); st0 = 0.01 * rng.standard_normal((C, B, K, 2))
    pad = 64
    ybuf = torch.full((C, B, n + 2 * pad), 777.0, device="cuda", dtype=torch.float64)      # guard bands around every row
    y = ybuf[:, :, pad:pad + n]
    bank = p.Bank(sos)
    cut = int(rng.integers(1, 8)) * (n // 8) // 4 * 4 if rng.random() < 0.5 else n
    xt = torch.from_numpy(x).cuda(); st = torch.from_numpy(st0.copy()).cuda()
    _, st = p.bank_filter(bank, xt[:, :cut], st, out=y[:, :, :cut]); info = bank.last_launch()
    if cut < n: _, st = p.bank_filter(bank, xt[:, cut:], st, out=y[:, :, cut:])
    wide += info["aligned"] == 3
    y0, s0 = oracle.sos_bank(x, sos, st0)
    y1 = y.cpu().numpy()
    err = max(rel(y1[:, b], y0[:, b]) for b in range(B))
    guards_ok = bool((ybuf[:, :, :pad] == 777.0).all() and (ybuf[:, :, pad + n:] == 777.0).all())
    if not (err <= 1e-9 and rel(st.cpu().numpy(), s0) <= 1e-7 and guards_ok):
        print("FAIL K=%d B=%d C=%d n=%d cut=%d err=%.2e guards=%s %s" % (K, B, C, n, cut, err, guards_ok, info)); sys.exit(1)
    worst = max(worst, err); cases += 1
print("wide sweep ok: %d cases (%d first blocks on the wide kernel), worst band error %.2e" % (cases, wide, worst))
