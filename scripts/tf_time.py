"""Weighting kernel (tf_seq_kernel) timing over batch shapes: clocks per sample of one warp."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyfilterbank_b200 import splweighting as spl
b, a = spl.a_weighting_coeffs_design(48000)
for C, n in ((1024, 240000), (64, 240000), (8192, 48000)):
    x = torch.randn(C, n, device="cuda", dtype=torch.float64)
    for _ in range(2): spl.tf_filter(b, a, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); spl.tf_filter(b, a, x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("C=%d n=%d: %.2f ms  %.0f clk/sample/warp" % (C, n, ms, ms * 1e-3 * 1.9e9 / n))
