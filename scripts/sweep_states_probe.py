"""Final states after a log-sine sweep: the time-chunked path with and without pre-pass truncation and the sequential
kernel against the oracle (sosfilt.c arithmetic) and against an extended-precision evaluation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from oracle import oracle
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
B, K = sos.shape[:2]
C, n, fs = 2, 480000, 48000.0
t = np.arange(n) / fs
T, f0, f1 = n / fs, 10.0, 0.45 * fs
k = (f1 / f0) ** (1.0 / T)
phase = 2 * np.pi * f0 * (k ** t - 1.0) / np.log(k)
x = 0.5 * np.sin(phase[None, :] + 2 * np.pi * np.arange(C)[:, None] / C)
y0, s0 = oracle.sos_bank(x, sos)
# extended precision (80-bit) evaluation of band 0, channel 0
xl = x[0].astype(np.longdouble); c = sos[0].astype(np.longdouble)
w = np.zeros((K, 2), dtype=np.longdouble)
sig = xl.copy()
for kk in range(K):
    b0, b1, b2, _, a1, a2 = c[kk]
    w1 = w2 = np.longdouble(0)
    out = np.empty_like(sig)
    for i in range(n):
        w0 = sig[i] - a1 * w1 - a2 * w2
        out[i] = b0 * w0 + b1 * w1 + b2 * w2
        w2 = w1; w1 = w0
    w[kk] = (w1, w2); sig = out
print("band 0 ch 0 states, 80-bit:", np.array(w, dtype=np.float64).ravel())
print("oracle (double, sosfilt.c order):", s0[0, 0].ravel())
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
xt = torch.from_numpy(x).cuda()
VARIANTS = [("scan default", {}, None, {}), ("scan warm_tol=0", {}, 0.0, {}), ("seq fma", {"mode": "seq"}, None, {}), ("seq exact", {"mode": "seq", "exact": True}, None, {}),
            ("scan no H-form", {}, None, {"PFB_NO_HFORM": "1"}), ("scan unscaled", {}, None, {"PFB_NO_SCALED": "1"}),
            ("scan J=1", {}, None, {"PFB_TMA_J": "1"}), ("scan J=2", {}, None, {"PFB_TMA_J": "2"}), ("scan J=32", {}, None, {"PFB_TMA_J": "32"}),
            ("scan J=2 unscaled", {}, None, {"PFB_TMA_J": "2", "PFB_NO_SCALED": "1"}),
            ("scan J=2 no H unscaled", {}, None, {"PFB_TMA_J": "2", "PFB_NO_SCALED": "1", "PFB_NO_HFORM": "1"})]
for name, kw, tol, env in VARIANTS:
    for kname in ("PFB_NO_HFORM", "PFB_NO_SCALED", "PFB_TMA_J"):
        os.environ.pop(kname, None)
    os.environ.update(env)
    bank = p.Bank(sos)
    if tol is not None:
        bank.set_warm_tol(tol)
    y1, s1 = p.bank_filter(bank, xt, **kw)
    s1 = s1.cpu().numpy(); y1 = y1.cpu().numpy()
    print("%-22s states vs oracle %.2e, band0 vs 80-bit %.2e (oracle vs 80-bit %.2e); outputs worst band %.2e; band-0 states %s" % (
        name, rel(s1, s0), rel(s1[0, 0], np.array(w, dtype=np.float64)), rel(s0[0, 0], np.array(w, dtype=np.float64)),
        max(rel(y1[:, b], y0[:, b]) for b in range(B)), s1[0, 0, 2]))
