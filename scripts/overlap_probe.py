"""Do two independent passes of the bank overlap when issued on two streams?  (Upper bound for any
pre-pass / output-pass pipelining across batches.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "designs.npz"))
s = g["third_48k_o4_sosmat"]; sos = np.ascontiguousarray(s.T).reshape(s.shape[1], -1, 6)
C, N = 64, 2880000
banks = [p.Bank(sos), p.Bank(sos)]
xs = [torch.randn(C, N, device="cuda", dtype=torch.float64) * 0.1 for _ in range(2)]
ys = [torch.empty(C, 33, N, device="cuda", dtype=torch.float64) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(concurrent, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        for i in range(2):
            if concurrent:
                streams[i].wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(streams[i]):
                    p.bank_filter(banks[i], xs[i], None, out=ys[i])
            else:
                p.bank_filter(banks[i], xs[i], None, out=ys[i])
    if concurrent:
        for st in streams: torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * iters)
run(False, 2); run(True, 2)
for name, c in (("sequential", False), ("two streams", True), ("sequential", False), ("two streams", True)):
    print("%-12s %.3f ms per pass" % (name, run(c, 5)))
