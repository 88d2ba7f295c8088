"""Throughput of the band-level epilogue (energies per block, no band signals) on the cfg-2 workload."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
C, N = 64, 2880000
for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
    bank = p.Bank(sos)
    x = torch.randn(C, N, device="cuda", dtype=dt) * 0.1
    for bl in (4800, 48000, N):
        for _ in range(2): p.bank_levels(bank, x, bl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): e, _ = p.bank_levels(bank, x, bl)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"case": "band levels, 64 ch x 60 s @48 kHz, 33 bands, %s, block %d" % (name, bl), "ms": ms,
                          "band_samples_per_s": C * 33 * N / (ms * 1e-3), "launch": bank.last_launch()}), flush=True)
