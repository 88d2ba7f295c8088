"""PCIe probe for the host-buffer path: plain pinned D2H/H2D copies vs pfb_bank_filter_host."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
n = 1 << 29   # 4 GiB of float64
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
for name, fn in (("D2H pinned contiguous 4 GiB", lambda: h.copy_(d, non_blocking=True)),
                 ("H2D pinned contiguous 4 GiB", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%s: %.1f GB/s" % (name, n * 8 / dt / 1e9))
# 2-D D2H like the block drain: 2112 rows of 508 KB into a host buffer with a 23 MB pitch
rows, width, pitch = 2112, 63552 * 8, 2880000 * 8
hb = torch.empty(rows * pitch // 8, dtype=torch.float64, pin_memory=True)
db = torch.empty(rows * width // 8, dtype=torch.float64, device="cuda")
hv = hb.view(rows, pitch // 8)[:, : width // 8]
dv = db.view(rows, width // 8)
hv.copy_(dv, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4): hv.copy_(dv, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
print("D2H 2-D (2112 rows x 508 KB, host pitch 23 MB): %.1f GB/s" % (rows * width / dt / 1e9))
