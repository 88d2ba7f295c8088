"""Device-resident throughput of the other BASELINE.json configs (parity-test cases, not bench.py lines),
on bounded blocks that fit one GPU.  One JSON line per case: band-samples/s (input-rate band-samples for
the decimated mode), algorithmic bytes, fraction of the measured HBM peak.

    python scripts/bench_configs.py [case ...]     cases: cfg1 cfg3_full cfg3_dec cfg4 cfg5 cfg5_f32
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pyfilterbank_b200 as p
from pyfilterbank_b200 import gammatone as gt, splweighting as spl

PEAK = 6551.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


LAST_CLOCK = {}


def _sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        return {"sm_mhz": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                "power_w": pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                "throttle": hex(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))}
    except Exception as e:   # the clock is context, not the measurement
        return {"error": repr(e)[:80]}


def timeit(fn, iters=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    LAST_CLOCK.update(_sm_clock())        # sampled while the last kernels are still running
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, units, alg_bytes, extra=None):
    d = {"case": name, "ms": ms, "band_samples_per_s": units / (ms * 1e-3), "algorithmic_GB": alg_bytes / 1e9,
         "achieved_GBps": alg_bytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / PEAK}
    if extra:
        d.update(extra)
    d["clock_during_run"] = dict(LAST_CLOCK)
    print(json.dumps(d), flush=True)


def cfg1():
    # default bank (fs 44100), 1 channel x 10 s, through the reference-facing call with numpy in/out
    ofb = p.FractionalOctaveFilterbank()
    x = np.random.default_rng(0).standard_normal(441000)
    ofb.filter_mimo_c(x)
    t0 = time.perf_counter(); its = 5
    for _ in range(its):
        y, st = ofb.filter_mimo_c(x)
    dt = (time.perf_counter() - t0) / its
    xt = torch.from_numpy(x).cuda().reshape(-1, 1)
    ms = timeit(lambda: ofb.filter_mimo_c(xt))
    report("cfg1 default bank 1 ch x 10 s f64 (numpy in/out, incl. H2D/D2H)", dt * 1e3, 441000 * 33, 441000 * 8 * 34,
           {"device_resident_ms": ms, "device_resident_band_samples_per_s": 441000 * 33 / (ms * 1e-3)})


def cfg3_full(C=1024, seconds=1):
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51)
    B, K, n = ofb.num_bands, 6, 48000 * seconds
    bank = p.Bank(np.ascontiguousarray(ofb.sosmat.T).reshape(B, K, 6))
    x = torch.randn(C, n, device="cuda", dtype=torch.float64) * 0.1
    y = torch.empty(C, B, n, device="cuda", dtype=torch.float64)
    ms = timeit(lambda: p.bank_filter(bank, x, None, out=y))
    report("cfg3 1/12-oct order 6 full rate, %d ch x %d s f64" % (C, seconds),
           ms, C * B * n, C * n * 8 * (B + 1),
           {"launch": bank.last_launch()})


def cfg3_dec(C=1024, seconds=4):
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51)
    n = 48000 * seconds
    x = (torch.randn(C, n, device="cuda", dtype=torch.float64) * 0.1).t()          # (N, C) view
    load = float((1.0 / ofb.decimation_factors).sum())
    ms = timeit(lambda: ofb.filter_decimated(x), iters=2, warm=1)
    alg = C * n * 8 * (1 + load)
    report("cfg3 1/12-oct order 6 DECIMATED, %d ch x %d s f64%s" % (C, seconds, " (unfused)" if os.environ.get("PFB_DEC_UNFUSED") else ""),
           ms, C * ofb.num_bands * n, alg, {"band_equivalents": load, "bands": ofb.num_bands, "dec": ofb._dec_bank().info()})


def cfg3_stream(C=1024, block_s=3, blocks=10):
    """configs[2] at its full duration: 1024 channels x 30 s streamed as 10 blocks of 3 s with carried states
    (anti-alias, decimator phase and band states); the 0.43 TB of decimated bands are produced and dropped."""
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51)
    n = 48000 * block_s
    x = (torch.randn(C, n, device="cuda", dtype=torch.float64) * 0.1).t()
    load = float((1.0 / ofb.decimation_factors).sum())
    st = None
    for _ in range(2):
        _, st = ofb.filter_decimated(x, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(blocks):
        out, st = ofb.filter_decimated(x, st)
        del out
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    report("cfg3 1/12-oct order 6 DECIMATED, %d ch x %d s streamed in %d-s blocks with carried states" % (C, block_s * blocks, block_s),
           ms, C * ofb.num_bands * n * blocks, C * n * blocks * 8 * (1 + load), {"band_equivalents": load})


def cfg4(C=512, seconds=10):
    gfb = gt.GammatoneFilterbank(samplerate=16000, order=4, startband=-14, endband=18, density=0.5)
    B, n = len(gfb.centerfrequencies), 16000 * seconds
    x = torch.randn(C, n, device="cuda", dtype=torch.float64) * 0.1
    ms = timeit(lambda: gt.fos_bank(x, gfb._coef_rows, 4), iters=2, warm=1)
    report("cfg4 gammatone order 4, %d bands, %d signals x %d s @16 kHz, c128 out" % (B, C, seconds), ms, C * B * n,
           C * n * (8 + 16 * B))


def cfg5(C=1024, seconds=5, dtype=torch.float64, fused=None):
    """A-weighting + 1/3-octave bank through pfb_bank_filter_weighted (one call).  fused: None = planner's choice,
    True / False force the fused kernel / the weighting kernel + chunked bank."""
    ofb = p.FractionalOctaveFilterbank(sample_rate=48000)
    B, K, n = 33, 4, 48000 * seconds
    bank = p.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(B, K, 6))
    b, a = spl.a_weighting_coeffs_design(48000)
    x = (torch.randn(C, n, device="cuda", dtype=torch.float64) * 0.1).to(dtype)
    y = torch.empty(C, B, n, device="cuda", dtype=dtype)
    esz = 8 if dtype == torch.float64 else 4
    if fused is None:
        os.environ.pop("PFB_WEIGHT_FUSED", None)
    else:
        os.environ["PFB_WEIGHT_FUSED"] = "1" if fused else "0"
    ms = timeit(lambda: p.bank_filter_weighted(bank, x, b, a, out=y), iters=3, warm=1)
    info = bank.last_launch()
    os.environ.pop("PFB_WEIGHT_FUSED", None)
    msw = timeit(lambda: spl.tf_filter(b, a, x.to(torch.float64)), iters=2, warm=1) if C <= 1024 else None
    report("cfg5 A-weighting + 1/3-oct bank, %d ch x %d s, %s, fused=%s nb=%s" % (
               C, seconds, "f64" if esz == 8 else "f32", fused, os.environ.get("PFB_TMA_NB", "auto")) +
           (" J=1" if os.environ.get("PFB_TMA_J") else ""), ms,
           C * B * n, C * n * esz * (B + 1), {"weighting_kernel_alone_ms": msw, "launch": info})


def cfg5_sweep():
    """Bands per CTA of the fused weighted bank (PFB_TMA_NB <= 9) and larger shards."""
    for dtype in (torch.float64, torch.float32):
        cfg5(dtype=dtype)                                  # planner's choice
        cfg5(dtype=dtype, fused=False)
        for nb in (5, 7, 9):
            os.environ["PFB_TMA_NB"] = str(nb)
            cfg5(dtype=dtype, fused=True)
            torch.cuda.empty_cache()
        os.environ.pop("PFB_TMA_NB", None)
        for C in (2048, 4096):
            cfg5(C=C, seconds=5 if C <= 2048 else 2, dtype=dtype)
            torch.cuda.empty_cache()


def cfg3_unfused(C=1024, seconds=3):
    os.environ["PFB_DEC_UNFUSED"] = "1"
    cfg3_dec(C, seconds)
    os.environ.pop("PFB_DEC_UNFUSED", None)


CASES = {"cfg1": cfg1, "cfg3_full": cfg3_full, "cfg3_dec": cfg3_dec, "cfg3_stream": cfg3_stream, "cfg4": cfg4, "cfg5": cfg5,
         "cfg5_f32": lambda: cfg5(dtype=torch.float32), "cfg5_sweep": cfg5_sweep, "cfg3_unfused": cfg3_unfused}
if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
        torch.cuda.empty_cache()
