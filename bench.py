#!/usr/bin/env python
"""bench.py -- channel*band samples/s of the 1/3-octave MIMO filter bank on N B200 GPUs.

Workload (BASELINE.json configs[1]): 1/3-octave Butterworth bank at 48 kHz (33 bands x 4 biquads,
the FractionalOctaveFilterbank defaults at fs=48000), 64 channels x 60 s of white noise PER GPU,
float64 (default) or float32.  One step = one pass of the bank over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl ours|reference]

`value` is timed with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max
over ranks).  `e2e` is the same pass through the C ABI on HOST buffers (pfb_bank_filter_host:
pinned host input, H2D, kernels, D2H of every band signal inside the timed region).
`--impl reference` times the reference's own C (oracle/_ref, built from the unmodified sosfilt.c)
on the host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
SECONDS = 60
CHANNELS_PER_GPU = 64
METRIC = "channel*band samples/s (1/3-oct MIMO bank)"
UNIT = "band-samples/s"


def bank_sos():
    from pyfilterbank_b200.octbank import FractionalOctaveFilterbank
    import pyfilterbank_b200.octbank as ob
    # design only (numpy); constructing the class does not touch the GPU
    cf, edges = ob.frequencies_fractional_octaves(-19, 13, 1000.0, 3.0)
    sosmat = ob.design_sosmat_band_passes(4, edges, FS, 0.01)
    B = sosmat.shape[1]
    return np.ascontiguousarray(sosmat.T).reshape(B, 4, 6)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread every 5 ms
    (nvidia_ml_py), falling back to an `nvidia-smi -lms` subprocess."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []          # (time, sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.thread = None
        self.smmax = None
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"

            def loop():
                while not self.stop_flag:
                    try:
                        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.time(), float(clk), int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "smi"

            def read():
                for line in self.proc.stdout:
                    f = [v.strip() for v in line.split(",")]
                    try:
                        self.smmax = float(f[1])
                        self.samples.append((time.time(), float(f[0]), int(f[2], 16)))
                    except Exception:
                        pass
            self.thread = threading.Thread(target=read, daemon=True)
            self.thread.start()
        except Exception:
            self.mode = None

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.mode == "smi":
            time.sleep(0.1)
            self.proc.terminate()
        elif self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.smmax, "reasons": ["no clock samples (%s)" % self.mode],
                    "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:   # region shorter than the sampling period: nearest samples
            inside = sorted(self.samples, key=lambda s: abs(s[0] - 0.5 * (t0 + t1)))[:3]
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = sorted(n for n, bit in self.BAD.items() if mask & bit)
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": self.smmax, "reasons": reasons,
                "samples": len(inside), "source": self.mode}


def bind_to_gpu_numa_node(gpu_index):
    """Several ranks share one host: run this rank (and first-touch its pinned buffers) on the CPU cores NVML reports
    as local to its GPU, so the host side of the PCIe copies stays on the GPU's NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu_index]) if vis else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def cpu_reference_run(steps, warmup, sos, cores=None, seconds=None):
    """Reference C (`sosfilter_double`, sosfilt.c:31-54, via ctypes) over all host cores:
    work items are (channel, band) cascades of a bounded sample of the workload."""
    import ctypes
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as o
    o.build()
    kind = "reference" if o.have_ref() else "port"
    fn = o.ref().sosfilter_double if kind == "reference" else o.lib().oracle_sosfilter_double
    cores = cores or os.cpu_count() or 1
    B, K = sos.shape[:2]
    seconds = seconds or 20
    n = FS * seconds
    C = cores                     # one channel per core so every thread has B cascades of work
    rng = np.random.default_rng(1234)
    x = 0.1 * rng.standard_normal((C, n))
    dp = ctypes.POINTER(ctypes.c_double)
    sos_rows = [np.ascontiguousarray(sos[b]).reshape(-1) for b in range(B)]

    def work(c):
        buf = np.empty(n)
        st = np.zeros(2 * K)
        for b in range(B):
            np.copyto(buf, x[c])
            st[:] = 0
            if kind == "reference":
                fn(buf.ctypes.data_as(dp), n, sos_rows[b].ctypes.data_as(dp), K, st.ctypes.data_as(dp))
            else:
                fn(buf.ctypes.data_as(dp), n, sos_rows[b].ctypes.data_as(dp), K, st.ctypes.data_as(dp))
        return buf[-1]

    times = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            list(ex.map(work, range(C)))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    per_step = float(np.mean(times))
    value = C * B * n / per_step
    t0 = time.perf_counter()
    work(0)                                   # one channel's 33 cascades on one core, for the per-core figure
    single = B * n / (time.perf_counter() - t0)
    sample = "%d channels x %d bands x %d s @ %d Hz float64 per step, %s via ctypes, %d threads" % (
        C, B, seconds, FS, "oracle/_ref sosfilter_double (unmodified sosfilt.c, gcc -O3)"
        if kind == "reference" else "oracle port", cores)
    return value, per_step, {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "single_core_value": single}


def cpu_extra_baselines(sos, seconds=2):
    """The other CPU figures SURVEY.md 8(d) asks for, single core, bounded samples: the reference's C through a
    restatement of its Python wrapper (sosfiltering.py:111-163: three buffer copies around the call), and the
    scipy.signal.lfilter calls the reference makes for the gammatone bank (gammatone.py:232-236) and the
    A-weighting prefilter (splweighting.py:80)."""
    out = {}
    try:
        from oracle import oracle as o
        B, K = sos.shape[:2]
        n = FS * seconds
        x = 0.1 * np.random.default_rng(1).standard_normal(n)
        t0 = time.perf_counter()
        for b in range(B):
            o.sosfilter_double_c(x, sos[b], use_ref=o.have_ref())
        dt = time.perf_counter() - t0
        out["python_wrapper"] = {"value": B * n / dt, "unit": UNIT, "cores": 1,
                                 "sample": "33 x sosfilter_double_c (wrapper copies + %s C) on %d s @ %d Hz, 1 channel" % (
                                     "oracle/_ref" if o.have_ref() else "oracle port", seconds, FS)}
    except Exception as e:   # the checker is optional for the bench line
        out["python_wrapper"] = {"error": repr(e)[:200]}
    try:
        from scipy.signal import lfilter
        from pyfilterbank_b200 import gammatone as gt, splweighting as spl
        gfb_coeffs = list(gt.design_filtbank_coeffs(16000, 4, gt.frequencies_gammatone_bank(-14, 18, 1000.0, 0.5),
                                                    bandwidth_factor=1.0))
        n = 16000 * 2
        x = np.random.default_rng(2).standard_normal(n)
        t0 = time.perf_counter()
        for b, a in gfb_coeffs:
            sig = x
            bb = b
            for i in range(4):
                sig, _ = lfilter(bb, a, sig, zi=[0j])
                bb = np.ones_like(b)
        dt = time.perf_counter() - t0
        out["gammatone_scipy"] = {"value": len(gfb_coeffs) * n / dt, "unit": UNIT, "cores": 1,
                                  "sample": "%d bands x 4 lfilter passes (gammatone.py:232-236) on 2 s @ 16 kHz" % len(gfb_coeffs)}
        b, a = spl.a_weighting_coeffs_design(FS)
        n = FS * 5
        x = np.random.default_rng(3).standard_normal(n)
        t0 = time.perf_counter()
        lfilter(b, a, x)
        dt = time.perf_counter() - t0
        out["a_weighting_scipy"] = {"value": n / dt, "unit": "channel-samples/s", "cores": 1,
                                    "sample": "lfilter(b, a, x) (splweighting.py:80) on 5 s @ 48 kHz"}
    except Exception as e:
        out["scipy"] = {"error": repr(e)[:200]}
    return out


LAST_RUNS = {}


def event_time(torch, fn, iters, warm, reps=2):
    """Mean time of `iters` calls between CUDA events after `warm` untimed calls, measured `reps` times with the SM clock
    and throttle reasons sampled meanwhile; returns the fastest repetition and leaves all of them in LAST_RUNS (these
    extra configs run once each, right after other heavy work: a single disturbed measurement would otherwise stand)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    runs = []
    for _ in range(reps):
        sampler = ClockSampler(torch.cuda.current_device())
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        clk = sampler.stop(t0, t1)
        runs.append({"ms": e0.elapsed_time(e1) / iters, "sm_mhz": clk.get("sm_mhz"), "reasons": clk.get("reasons")})
    LAST_RUNS.clear()
    LAST_RUNS["runs"] = runs
    return min(r["ms"] for r in runs)


def other_configs(torch, pfb, dist, dev, rank, world, peak, with_cpu=True):
    """Per-GPU shards of BASELINE.json configs[2..4] (extra keys; the headline is configs[1]): channels / signals
    are sharded over the ranks, long signals stream in time blocks with carried states, inputs are generated on the
    device.  Each entry: band-samples/s over all ranks (max time over ranks) and the fraction of the HBM roofline
    of its algorithmic bytes."""
    from pyfilterbank_b200 import gammatone as gt, splweighting as spl
    out = {}

    def finish(name, ms, units_per_rank, alg_bytes_per_rank, extra):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        d = {"value": world * units_per_rank / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
             "roofline_frac": None if alg_bytes_per_rank is None else alg_bytes_per_rank / (ms * 1e-3) / 1e9 / peak,
             "algorithmic_GB_per_gpu": None if alg_bytes_per_rank is None else alg_bytes_per_rank / 1e9}
        d.update(extra)
        d["runs"] = list(LAST_RUNS.get("runs", []))     # every repetition (this rank) with its clocks; ms_per_step = fastest
        out[name] = d

    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    try:   # cfg 5: A-weighting fused with the 1/3-octave bank, 8192 channels sharded, streamed in 0.5 s blocks
        ofb = pfb.FractionalOctaveFilterbank(sample_rate=FS)
        B, K = ofb.num_bands, 4
        bank = pfb.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(B, K, 6))
        b, a = spl.a_weighting_coeffs_design(FS)
        C = 8192 // world
        nblk, blocks = FS // 2, 3
        for name, tdt, esz in (("cfg5_f32", torch.float32, 4), ("cfg5_f64", torch.float64, 8)):
            x = torch.randn((C, nblk), generator=gen, device=dev, dtype=tdt) * 0.1
            y = torch.empty((C, B, nblk), device=dev, dtype=tdt)
            st = torch.zeros((C, B, K, 2), device=dev, dtype=tdt)
            wz = torch.zeros((C, 6), device=dev, dtype=torch.float64)

            def run():
                for _ in range(blocks):      # consecutive blocks of the stream: states and delay values carried
                    pfb.bank_filter_weighted(bank, x, b, a, st, wz, out=y)
            ms = event_time(torch, run, 2, 2)
            finish(name, ms, C * B * nblk * blocks, C * nblk * blocks * esz * (B + 1),
                   {"workload": "A-weighting + 1/3-octave bank, %d of 8192 channels per GPU, %d blocks of 0.5 s @ 48 kHz "
                                "with carried states, %s" % (C, blocks, "float32" if esz == 4 else "float64"),
                    "path": bank.last_launch()["aligned"]})
            if esz == 4:   # the monitoring form of cfg 5: A-weighted band energies per 0.1 s, no band signal written
                st.zero_()
                wz.zero_()
                del y

                def run_lev():
                    for _ in range(blocks):
                        pfb.bank_levels(bank, x, 4800, st, weighting=(b, a), wz=wz)
                ms = event_time(torch, run_lev, 2, 2)
                finish("cfg5_f32_levels", ms, C * B * nblk * blocks, None,
                       {"workload": "A-weighting + 1/3-octave bank + band energies per 0.1 s in one kernel "
                                    "(pfb_bank_levels_weighted), %d of 8192 channels per GPU, %d blocks of 0.5 s, float32" % (C, blocks),
                        "bound": "FP32 / FP64 issue: only the input (%.2f GB per GPU) is read, no band signal is written, so "
                                 "there is no HBM roofline for it" % (C * nblk * blocks * esz / 1e9)})
                y = None
            del x, y, st, wz
            torch.cuda.empty_cache()
    except Exception as e:
        out["cfg5_error"] = repr(e)[:300]
    try:   # cfg 3: 1/12-octave order-6 bank with per-band decimation, 1024 channels sharded, 3 s blocks
        ofb = pfb.FractionalOctaveFilterbank(sample_rate=FS, order=6, nth_oct=12.0, start_band=-67, end_band=51)
        C = 1024 // world
        nblk = 3 * FS
        load = float((1.0 / ofb.decimation_factors).sum())
        x = (torch.randn((C, nblk), generator=gen, device=dev, dtype=torch.float64) * 0.1).t()   # (N, C) view
        state = [None]

        def run():
            bands, state[0] = ofb.filter_decimated(x, state[0])
            del bands
        ms = event_time(torch, run, 2, 2)
        info = ofb._dec_bank().info()
        finish("cfg3_decimated", ms, C * ofb.num_bands * nblk, C * nblk * 8 * (1 + load),
               {"workload": "1/12-octave order-6 bank, per-band decimation, %d of 1024 channels per GPU, 3 s blocks "
                            "@ 48 kHz with carried states, float64" % C,
                "band_equivalents": load, "bands": ofb.num_bands, "kernels_per_block": info["last_kernels"],
                "fused_stages": info["last_fused_stages"]})
        del x, state
        torch.cuda.empty_cache()
    except Exception as e:
        out["cfg3_error"] = repr(e)[:300]
    if rank == 0:
        try:   # cfg 1: the reference's own CPU-runnable case through the reference-facing call, numpy in / numpy out
            ofb1 = pfb.FractionalOctaveFilterbank()                         # fs 44100, order 4, 1/3 octave, 33 bands
            x1 = np.random.default_rng(0).standard_normal(441000)
            for _ in range(3):
                ofb1.filter_mimo_c(x1)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                y1, _ = ofb1.filter_mimo_c(x1)
                ts.append(time.perf_counter() - t0)
            xt = torch.from_numpy(x1).to(dev).reshape(-1, 1)
            ms_dev = event_time(torch, lambda: ofb1.filter_mimo_c(xt), 3, 2)
            d = {"value": 441000 * 33 / min(ts), "unit": UNIT, "ms_per_call": min(ts) * 1e3, "ms_per_call_all": [t * 1e3 for t in ts],
                 "device_resident_ms": ms_dev, "h2d_bytes": 441000 * 8, "d2h_bytes": 441000 * 33 * 8,
                 "workload": "FractionalOctaveFilterbank() default (fs 44100, order 4, 33 bands), 1 channel x 10 s white noise, "
                             "float64, filter_mimo_c(numpy) -> numpy: host-to-device copy, kernels, device-to-host copy of the "
                             "(N, B, C) result included (rank 0 only)"}
            try:   # the same call on the CPU checker (one core, like the reference's own wrapper + C loop)
                if not with_cpu:
                    raise RuntimeError("--no-cpu")
                from oracle import oracle as o
                t0 = time.perf_counter()
                o.sosfilter_double_mimo_c(x1, ofb1.sosmat, use_ref=o.have_ref())
                d["cpu_same_call_ms"] = (time.perf_counter() - t0) * 1e3
                d["cpu_kind"] = "oracle wrapper restatement + " + ("oracle/_ref sosfilter_double_mimo" if o.have_ref() else "oracle port")
            except Exception as e:
                d["cpu_same_call_ms"] = None
                d["cpu_error"] = repr(e)[:200]
            out["cfg1_default_bank"] = d
            del xt
        except Exception as e:
            out["cfg1_error"] = repr(e)[:300]
    try:   # cfg 4: gammatone, 64 ERB bands, 512 signals x 10 s @ 16 kHz sharded over the ranks
        gfb = gt.GammatoneFilterbank(samplerate=16000, order=4, startband=-14, endband=18, density=0.5)
        Bg, C, n = len(gfb.centerfrequencies), 512 // world, 160000
        x = torch.randn((C, n), generator=gen, device=dev, dtype=torch.float64) * 0.1
        ms = event_time(torch, lambda: gt.fos_bank(x, gfb._coef_rows, 4), 2, 2)
        finish("cfg4_gammatone", ms, C * Bg * n, C * n * (8 + 16 * Bg),
               {"workload": "gammatone order 4, %d ERB bands, %d of 512 signals x 10 s @ 16 kHz per GPU, complex128 out" % (Bg, C)})
        del x
        torch.cuda.empty_cache()
    except Exception as e:
        out["cfg4_error"] = repr(e)[:300]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=int, default=SECONDS)
    ap.add_argument("--channels", type=int, default=CHANNELS_PER_GPU)
    ap.add_argument("--signal", default="noise", choices=["noise", "sweep"],
                    help="synthetic input: white noise (sigma 0.1) or a log-sine sweep 10 Hz -> 0.45 fs, amplitude 0.5, "
                         "per-channel phase offset (SURVEY.md 8d)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra `configs` block (cfg 3 / 4 / 5 shards)")
    ap.add_argument("--no-f32", action="store_true", help="skip the float32 lines of the default (float64) run")
    ap.add_argument("--e2e-channels", type=int, default=32,
                    help="channels per GPU of the host-buffer (e2e) measurement, the same at every GPU count")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    sos = bank_sos()
    B, K = sos.shape[:2]
    n = FS * args.seconds
    C = args.channels
    esz = 8 if args.dtype == "f64" else 4
    config = {"workload": "1/3-octave filter_mimo_c bank, %d channels x %d s @ %d Hz per GPU, %s" % (
                  C, args.seconds, FS, "float64" if esz == 8 else "float32"),
              "bands": B, "sections_per_band": K, "channels_per_gpu": C, "samples": n,
              "sharding": "channels, no collective on the filtering path",
              "l2": "inputs (%.2f GB) and outputs (%.1f GB) exceed the 126 MB L2; no flush needed" % (
                  C * n * esz / 1e9, C * B * n * esz / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        value, per_step, cpu = cpu_reference_run(args.steps, max(args.warmup, 1), sos)
        config["cpu_sample"] = ("bounded sample of the workload, scaled linearly (the work is exactly linear in channels "
                                "and samples; sosfilter_double_mimo's 32-bit indices cannot address the full workload): "
                                + cpu["sample"])
        cpu.update(cpu_extra_baselines(sos))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": cpu,
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import pyfilterbank_b200 as pfb
    from pyfilterbank_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    # stdout carries exactly one JSON line: everything libraries print meanwhile (NCCL's version banner, ...) goes
    # to stderr; the real stdout is restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tdt = torch.float64 if esz == 8 else torch.float32
    bank = pfb.Bank(sos)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    if args.signal == "noise":
        x = torch.randn((C, n), generator=gen, device=dev, dtype=tdt) * 0.1
    else:   # log-sine sweep, the same on every channel up to a phase offset 2 pi c / C_total
        t = torch.arange(n, device=dev, dtype=torch.float64) / FS
        f0, f1, T = 10.0, 0.45 * FS, n / FS
        k = (f1 / f0) ** (1.0 / T)
        phase = 2 * np.pi * f0 * (k ** t - 1.0) / np.log(k)
        ch = (torch.arange(C, device=dev, dtype=torch.float64) + rank * C) / (world * C)
        x = (0.5 * torch.sin(phase[None, :] + 2 * np.pi * ch[:, None])).to(tdt)
        del t, phase
    y = torch.empty((C, B, n), device=dev, dtype=tdt)
    # float32 signals: the headline runs on float64 states with the arithmetic chosen per band (PFB_ARITH_AUTO: float32
    # where the library's calibration shows it within 1e-5 of float64, float64 elsewhere) -- the fastest arithmetic that
    # meets the 1e-4 tolerance in every band; float64 everywhere and pure float32 are reported beside it (`f32` block)
    arith64 = "auto" if esz == 4 else False
    states = torch.zeros((C, B, K, 2), device=dev, dtype=torch.float64 if arith64 else tdt)

    def step():
        states.zero_()
        pfb.bank_filter(bank, x, states, out=y, arith64=arith64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-kernel CUDA events inside the library, recorded on the launching stream during the timed steps
    bank.timing_begin(args.steps)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    kt = bank.timing_end()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    info = bank.last_launch()
    ms_per_step = ms / args.steps
    total_units = world * C * B * n
    value = total_units / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (the output pass: reads every input sample, writes every band
    # sample).  algorithmic bytes per launch = (B + 1) * C * n * element size (s_out + s_in / B
    # per band-sample); duration = that kernel's mean launch time from the library's CUDA events over the
    # timed steps.  `step_frac` is the same bytes over the WHOLE step (pre-pass + carry + output pass).
    alg_bytes = C * n * esz * (B + 1)
    peak, peak_src = peaks()
    tma = info["aligned"] in (2, 3)
    kname = ("sos_tma_wide_kernel (output pass, 512-byte fragments)" if info["aligned"] == 3 else
             "sos_tma_kernel<PASS=1> (output pass)" if tma else
             "sos_scan_kernel" if info["mode"] == 2 else "sos_seq_kernel")
    kms = kt["output"] if kt["calls"] else ms_per_step
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    traffic = None
    traffic_source = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and not arith64:
        with open(tpath) as f:
            tj = json.load(f)
        # dram bytes per band-sample from the committed ncu capture, scaled to this launch (NOT measured in this run)
        per_unit = tj.get(args.dtype, {}).get("dram_bytes_per_band_sample")
        if per_unit:
            traffic = per_unit * C * B * n
            traffic_source = "profiles/traffic.json (one committed ncu --set full capture of this kernel, scaled per band-sample)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src, "kernel": kname,
                "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                "step_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernels_ms": {k: kt[k] for k in ("prepass", "carry", "output", "total")},
                "kernel_share_of_step": kms / ms_per_step}

    def allmax(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    # optional gather of a bounded slab of band outputs, timed separately (never inside `value`)
    gather = None
    if world > 1:
        from pyfilterbank_b200.sharding import gather_bands
        ns = FS  # one second of every band of every local channel
        slab = y[:, :, :ns].contiguous()
        gather_bands(slab, world * C)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = gather_bands(slab, world * C)
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1)
        t = torch.tensor([gms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = float(t.item())
        gather = {"ms": gms, "bytes_received_per_rank": int(full.numel() * esz),
                  "GBps_per_rank": full.numel() * esz / (gms * 1e-3) / 1e9,
                  "what": "all_gather_into_tensor (NCCL) of 1 s of every band; reported separately"}
        del full, slab

    # float32 signals on the same workload, both arithmetics, each with its parity (worst band relative L2 against this
    # library's float64 path -- itself <= 1e-9 of the reference's sosfilt.c in tests/ -- on the float32-rounded input)
    f32 = None
    if not args.no_f32:
        x32 = x if esz == 4 else x.to(torch.float32)
        npar, cpar = min(n, 20 * FS), min(C, 2)
        xp = x32[:cpar, :npar].contiguous()
        yref, _ = pfb.bank_filter(bank, xp.to(torch.float64), None)
        den = yref.norm(dim=2)
        if esz == 8:
            del y
            torch.cuda.empty_cache()
            y = torch.empty((C, B, n), device=dev, dtype=torch.float32)
        f32 = {}
        for key, a64 in (("arith_f64", True), ("arith_auto", "auto"), ("arith_f32", False)):
            yp, _ = pfb.bank_filter(bank, xp, None, arith64=a64)
            err = ((yp.to(torch.float64) - yref).norm(dim=2) / den).max(dim=0).values.cpu().numpy()
            st32 = torch.zeros((C, B, K, 2), device=dev, dtype=torch.float64 if a64 else torch.float32)

            def step32():
                st32.zero_()
                pfb.bank_filter(bank, x32, st32, out=y, arith64=a64)
            ms32 = allmax(event_time(torch, step32, min(args.steps, 10), 3))
            f32[key] = {"value": world * C * B * n / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32,
                        "step_frac": C * n * 4 * (B + 1) / (ms32 * 1e-3) / 1e9 / peak,
                        "path": bank.last_launch()["aligned"],
                        "parity": {"vs": "float64 path of this library on the float32-rounded input, %d channels x %d s" % (cpar, npar // FS),
                                   "worst_band_rel_l2": float(err.max()), "bands_within_1e-4": int((err <= 1e-4).sum()),
                                   "bands": int(B), "tolerance": 1e-4}}
            if a64 == "auto":
                f32[key]["f32_bands"] = bank.f32_bands()
            del yp, st32
        f32["note"] = ("arith_auto = float32 signals, float64 states, arithmetic chosen per band (PFB_ARITH_AUTO: float32 "
                       "where the library's calibration shows it within 1e-5 of float64, float64 elsewhere): meets 1e-4 "
                       "in every band; arith_f64 = float32 signals, float64 coefficients / states / arithmetic (PFB_ARITH_F64): meets the "
                       "1e-4 tolerance in every band; arith_f32 = pure float32 arithmetic: like the reference's own "
                       "float path it misses 1e-4 below ~500 Hz (DF-II states of 1e12, SURVEY.md H3)")
        del yref, xp, den
        if esz == 8:
            del x32

    # end to end through the C ABI on host buffers
    e2e = None
    if not args.no_e2e:
        del y
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(5, args.steps))
        # The same per-GPU workload at every GPU count: a slice of `e2e-channels` channels of the workload (32: 24 GB
        # of pinned host memory for the band signals per rank; all ranks share one host).  The path is PCIe-bound, so
        # the rate does not depend on the slice size; `value` counts the band-samples actually moved.
        Ce = min(C, args.e2e_channels)
        xh = torch.empty((Ce, n), dtype=tdt, pin_memory=True)
        xh.copy_(x[:Ce].cpu())
        yh = torch.empty((Ce, B, n), dtype=tdt, pin_memory=True)
        sth = np.zeros((Ce, B, K, 2), dtype=np.float64)      # float32 signals run float64 arithmetic (see above)
        dtype_code = _lib.PFB_F64 if esz == 8 else _lib.PFB_F32
        e2e_flags = (_lib.ARITH_F64 | _lib.ARITH_AUTO) if esz == 4 else 0

        # what the link gives: plain pinned device-to-host copies on all ranks at once (no kernels, 2 GiB each)
        probe_n = min(yh.numel(), (2 << 30) // esz)
        dprobe = torch.empty(probe_n, device=dev, dtype=tdt)
        hprobe = yh.view(-1)[:probe_n]
        hprobe.copy_(dprobe, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            hprobe.copy_(dprobe, non_blocking=True)
        torch.cuda.synchronize()
        rate = 2 * probe_n * esz / (time.perf_counter() - t0) / 1e9
        del dprobe
        if world > 1:
            tr = torch.tensor([rate, -rate], device=dev, dtype=torch.float64)
            tsum = tr.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
            pcie = {"d2h_GBps_per_rank_min": float(-tr[1].item()), "d2h_GBps_per_rank_max": float(tr[0].item()),
                    "d2h_GBps_aggregate": float(tsum[0].item())}
        else:
            pcie = {"d2h_GBps_per_rank_min": rate, "d2h_GBps_per_rank_max": rate, "d2h_GBps_aggregate": rate}
        pcie["what"] = "plain pinned 2 GiB device-to-host copies, all ranks concurrently, no kernels: the host-side ceiling of e2e"

        def e2e_step():
            sth[:] = 0
            _lib.check(_lib.lib().pfb_bank_filter_host(bank._h, dtype_code, xh.data_ptr(), n, yh.data_ptr(), n,
                                                       sth.ctypes.data, n, Ce, e2e_flags))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * Ce * B * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(Ce * n * esz + sth.nbytes),
               "d2h_bytes_per_step": int(Ce * B * n * esz + sth.nbytes), "ms_per_step": dt * 1e3,
               "steps": e2e_steps, "channels_per_gpu": Ce,
               "config": "%d of the %d channels per GPU x %d s, the same slice at every GPU count" % (Ce, C, args.seconds),
               "pcie_probe": pcie,
               "link_frac": (Ce * (B + 1) * n * esz / dt / 1e9) / max(pcie["d2h_GBps_per_rank_min"], 1e-9),
               "api": "pfb_bank_filter_host (C ABI, pinned host buffers)",
               "checksum": float(yh[0, B // 2, :1000].double().abs().sum())}
        # the same host input reduced to band energies per 0.1 s inside the kernel (pfb_bank_levels_host): nothing of
        # size N*B*C crosses the bus, the path is bound by the input copy
        bl = 4800
        lev = np.empty((Ce, B, n // bl))

        def lev_step():
            sth[:] = 0
            _lib.check(_lib.lib().pfb_bank_levels_host(bank._h, dtype_code, xh.data_ptr(), n, lev.ctypes.data, n // bl,
                                                       sth.ctypes.data, n, Ce, bl, e2e_flags))
        lev_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            lev_step()
        dtl = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([dtl], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtl = float(t.item())
        e2e["band_levels"] = {"value": world * Ce * B * n / dtl, "unit": UNIT, "ms_per_step": dtl * 1e3,
                              "h2d_bytes_per_step": int(Ce * n * esz + sth.nbytes), "d2h_bytes_per_step": int(lev.nbytes + sth.nbytes),
                              "api": "pfb_bank_levels_host: band energies per %d samples instead of band signals" % bl}
        del xh, yh

    configs = None
    if not args.no_configs:
        try:
            del y
        except NameError:
            pass
        del x
        torch.cuda.empty_cache()
        configs = other_configs(torch, pfb, dist, dev, rank, world, peak, with_cpu=not args.no_cpu)

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        _, _, cpu = cpu_reference_run(2, 1, sos, seconds=10)
        cpu.update(cpu_extra_baselines(sos))
        config["cpu_sample"] = "cpu_baseline is timed on a bounded sample and scaled linearly: " + cpu["sample"]

    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": args.dtype if esz == 8 else "f32 signals, f64 states, f32 / f64 arithmetic per band",
                          "data": "synthetic (%s)" % ("white noise" if args.signal == "noise" else "log-sine sweep"),
                          "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                          "e2e_band_levels": None if e2e is None else e2e.get("band_levels"),
                          "f32": f32, "configs": configs,
                          "gpu_launches": int(info["kernels"]) * args.steps, "clocks": clocks,
                          "launch": info, "gather": gather}), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
