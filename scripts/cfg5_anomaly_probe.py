"""Hunting the intermittent slow runs of the fused weighted bank (47 ms instead of 14-18 on one visit, 76 instead of 19 on
another): repeat the cfg 5 shards with fresh tensors, clock sampling on, and print every repetition."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import splweighting as spl, gammatone as gt
import pynvml
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)
ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
bank = p.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(33, 4, 6))
b, a = spl.a_weighting_coeffs_design(48000)
def clk(): return pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM), hex(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(H)), round(pynvml.nvmlDeviceGetPowerUsage(H) / 1000)
def other_load():
    gfb = gt.GammatoneFilterbank(samplerate=16000, order=4, startband=-14, endband=18, density=0.5)
    x = torch.randn((512, 160000), device="cuda", dtype=torch.float64) * 0.1
    gt.fos_bank(x, gfb._coef_rows, 4); torch.cuda.synchronize(); del x; torch.cuda.empty_cache()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    for C, nblk, tdt in ((8192, 24000, torch.float32), (8192, 24000, torch.float64), (1024, 240000, torch.float64)):
        if rep % 2: other_load()
        x = torch.randn((C, nblk), device="cuda", dtype=tdt) * 0.1
        y = torch.empty((C, 33, nblk), device="cuda", dtype=tdt)
        st = torch.zeros((C, 33, 4, 2), device="cuda", dtype=tdt); wz = torch.zeros((C, 6), device="cuda", dtype=torch.float64)
        times = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); p.bank_filter_weighted(bank, x, b, a, st, wz, out=y); e1.record()
            c = clk(); torch.cuda.synchronize(); times.append((round(e0.elapsed_time(e1), 2), c))
        print(rep, C, nblk, str(tdt)[6:], times, flush=True)
        del x, y, st, wz; torch.cuda.empty_cache()
