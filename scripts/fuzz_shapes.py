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
        # per-band arithmetic (PFB_ARITH_AUTO): whatever the calibration decides for this random bank, every band
        # stays within the float32 bar of 1e-4 against the double result
        y3, s3 = p.bank_filter(bank, torch.from_numpy(xf).cuda(), torch.from_numpy(st0.copy()).cuda(), arith64="auto")
        y3 = y3.cpu().numpy()
        e3 = max(rel(y3[:, b], yd[:, b]) for b in range(B))
        auto_cases = globals().get("auto_cases", 0) + 1
        auto_worst = max(globals().get("auto_worst", 0.0), e3)
        if not (e3 <= 1e-4 and np.isfinite(y3).all() and s3.dtype == torch.float64):
            print("FAIL auto K=%d B=%d C=%d n=%d err=%.2e f32 bands %d %s" % (K, B, C, n, e3, bank.f32_bands(), bank.last_launch())); sys.exit(1)
# gammatone banks: random batch sizes, lengths and orders through the planner
from pyfilterbank_b200 import gammatone as gt
t_end = time.time() + float(os.environ.get("SECONDS_FOS", 30))
fos_cases = 0
while time.time() < t_end:
    order = int(rng.choice([1, 2, 3, 4, 5, 6, 8])); nb = int(rng.integers(1, 30))
    C = int(rng.choice([1, 2, 5, 16, 33, 64])); n = int(rng.choice([1, 16, 100, 4096, 8000, 30000, 100002]))
    try:
        gfb = p.GammatoneFilterbank(samplerate=16000, order=order, startband=-nb // 2, endband=nb - nb // 2 - 1, density=0.7)
    except IndexError:   # like the reference, the constructor's delay estimate needs the impulse peak inside 20 ms
        continue
    x = rng.standard_normal((C, n))
    y, st = gt.fos_bank(torch.from_numpy(x).cuda(), gfb._coef_rows, order)
    for i in sorted({0, len(gfb._coeffs) - 1}):
        b, a = gfb._coeffs[i]
        for c in sorted({0, C - 1}):
            y0, s0 = oracle.fosfilter(b, a, order, x[c])
            e = rel(y[c, i].cpu().numpy(), y0)
            if not (e <= 1e-10 and rel(st[c, i].cpu().numpy(), s0) <= 1e-9):
                print("FAIL fos order=%d B=%d C=%d n=%d err=%.2e" % (order, len(gfb._coeffs), C, n, e)); sys.exit(1)
    fos_cases += 1
# band levels on random block structures
t_end = time.time() + float(os.environ.get("SECONDS_LEV", 30))
lev_cases = 0
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s48 = g["third_48k_o4_sosmat"]; sos48 = np.ascontiguousarray(s48.T).reshape(s48.shape[1], -1, 6)
bank48 = p.Bank(sos48)
while time.time() < t_end:
    C = int(rng.choice([1, 3, 8, 40])); bl = 32 * int(rng.choice([1, 2, 5, 20, 150])); nblk = int(rng.choice([1, 2, 7, 30, 100]))
    n = bl * nblk
    if C * 33 * n > 2e8: continue
    x = rng.standard_normal((C, n))
    y0, _ = oracle.sos_bank(x, sos48)
    e0 = (y0.reshape(C, 33, nblk, bl) ** 2).sum(axis=3)
    e, _ = p.bank_levels(bank48, torch.from_numpy(x).cuda(), bl)
    e = e.cpu().numpy()
    scale = e0.max(axis=2, keepdims=True)
    if not np.max(np.abs(e - e0) / scale) <= 2e-9:
        print("FAIL levels C=%d bl=%d nblk=%d err=%.2e" % (C, bl, nblk, np.max(np.abs(e - e0) / scale))); sys.exit(1)
    lev_cases += 1
# fused callers: weighting prefilter + bank (fused kernel forced or planner's choice), decimating band path
from pyfilterbank_b200 import splweighting as spl
t_end = time.time() + float(os.environ.get("SECONDS_FUSED", 30))
fused_cases = dec_cases = 0
ofb48 = p.FractionalOctaveFilterbank(sample_rate=48000)
while time.time() < t_end:
    C = int(rng.choice([1, 2, 31, 32, 33, 70, 129])); n = int(rng.choice([130, 1000, 4096, 5001, 20000, 65536]))
    wname = str(rng.choice(["A", "B", "C"]))
    b, a = spl._weighting_coeff_design_funsd[wname](48000)
    os.environ["PFB_WEIGHT_FUSED"] = str(int(rng.integers(0, 2)))
    x = rng.standard_normal((n, C))
    chans = sorted({0, C - 1})
    _, y0, _ = oracle.weighted_bank(x[:, chans], b, a, ofb48.sosmat)
    cut = int(rng.integers(1, n))
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    ya, st, wz = p.bank_filter_weighted(bank48, xt[:, :cut].contiguous(), b, a)
    yb, st, wz = p.bank_filter_weighted(bank48, xt[:, cut:].contiguous(), b, a, st, wz)
    yy = torch.cat([ya, yb], dim=2)[chans].cpu().numpy()
    e = max(rel(yy[:, k].T, y0[:, k]) for k in range(33))
    # Bar 5e-9 here (1e-9 in the fixed-seed tests): behind the bit-identical prefilter the lowest bands carry almost no
    # signal (A: -63 dB at 12.5 Hz) while their section states still see the broadband input, so FMA against separately
    # rounded arithmetic differs by up to ~2e-9 there on some draws; the reference's own float64 result is 5e-5 away
    # from the 80-bit evaluation of the same recurrences in those bands (profiles/r02_weighted_conditioning.txt).
    if not e <= 5e-9:
        print("FAIL weighted %s C=%d n=%d cut=%d fused=%s err=%.2e" % (wname, C, n, cut, os.environ["PFB_WEIGHT_FUSED"], e)); sys.exit(1)
    fused_cases += 1
    if rng.random() < 0.5:
        C = int(rng.choice([1, 3, 32, 40])); n = int(rng.choice([1, 17, 1000, 4096, 20001, 100000]))
        x = rng.standard_normal((n, C))
        chans = sorted({0, C - 1})
        ref = oracle.decimated_bank(x[:, chans], ofb48.decimation_plan(), ofb48.aa_sos, 4)
        cut = int(rng.integers(0, n + 1))
        ba, sd = ofb48.filter_decimated(torch.from_numpy(x[:cut]).cuda()) if cut else ([None] * 33, None)
        bb, sd = ofb48.filter_decimated(torch.from_numpy(x[cut:]).cuda(), sd) if cut < n else ([None] * 33, sd)
        for k in range(33):
            parts = [t[:, chans].cpu().numpy() for t in (ba[k], bb[k]) if t is not None]
            yk = np.concatenate(parts)
            if yk.shape != ref[k].shape or not rel(yk, ref[k]) <= 1e-9:
                print("FAIL decimated C=%d n=%d cut=%d band=%d err=%.2e" % (C, n, cut, k, rel(yk, ref[k]) if yk.shape == ref[k].shape else -1)); sys.exit(1)
        dec_cases += 1
os.environ.pop("PFB_WEIGHT_FUSED", None)
print("gammatone cases %d, band-level cases %d, weighted cases %d, decimated cases %d" % (fos_cases, lev_cases, fused_cases, dec_cases))
print("per-band arithmetic cases %d, worst band error %.2e" % (globals().get("auto_cases", 0), globals().get("auto_worst", 0.0)))
print("fuzz ok: %d cases, worst band error %.2e, launch kinds (mode, path) %s" % (cases, worst, modes))
