"""CPU tests of the host-side logic: coefficient design vs the reference's (golden designs),
the C-ABI library (loads, exports every declared symbol, fails loudly without a GPU), and the
channel-sharding helpers over a 2-rank gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden


def test_butter_sos_matches_reference():
    from pyfilterbank_b200 import butterworth as bw
    g = golden("designs")
    for band, args in {"lowpass": (4, 0.1), "highpass": (4, 0.2), "bandpass": (4, 0.1, 0.2),
                       "bandstop": (2, 0.15, 0.3)}.items():
        s = bw.butter_sos(band, *args)
        assert s.shape == g["butter_" + band].shape
        np.testing.assert_allclose(s, g["butter_" + band], rtol=1e-13, atol=1e-15)
        assert np.all(s[:, 3] == 1.0)
    with pytest.raises(Exception):
        bw.butter_sos("lowpass", 3, 0.1)        # odd pole count (butterworth.py:88-89)
    with pytest.raises(Exception):
        bw.butter_sos("lowpass", 4, 0.6)        # v1 >= 0.5 (butterworth.py:44-45)
    with pytest.raises(Exception):
        bw.butter_sos("bandpass", 4, 0.2, 0.1)  # v2 <= v1


def test_octbank_design_matches_reference():
    from pyfilterbank_b200 import octbank as ob
    g = golden("designs")
    for name, kw in {"third_48k_o4": dict(sample_rate=48000, order=4, nth_oct=3.0),
                     "twelfth_48k_o6": dict(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51),
                     "octave_44k_o2": dict(sample_rate=44100, order=2, nth_oct=1.0, start_band=-5, end_band=4)}.items():
        cf, edges = ob.frequencies_fractional_octaves(kw.get("start_band", -19), kw.get("end_band", 13), 1000.0,
                                                      kw["nth_oct"])
        assert np.array_equal(cf, g[name + "_cf"]) and np.array_equal(edges, g[name + "_edges"])
        sosmat = ob.design_sosmat_band_passes(kw["order"], edges, kw["sample_rate"], 0.01)
        np.testing.assert_allclose(sosmat, g[name + "_sosmat"], rtol=1e-13, atol=0)
    d = golden("octbank_default")
    cf, edges = ob.frequencies_fractional_octaves(-19, 13, 1000.0, 3.0)
    sosmat = ob.design_sosmat_band_passes(4, edges, 44100, 0.01)     # top edge clipped to 0.499 fs
    np.testing.assert_allclose(sosmat, d["sosmat"], rtol=1e-13, atol=0)
    # the reference's own design tests (tests/test_octbank.py:49-64)
    cf, edges = ob.frequencies_fractional_octaves(-20, 20, 1000.0, 3.0)
    assert cf[-20 - 1] == 1000.0 and len(cf) == 41 and len(edges) == 42
    assert np.all(np.isfinite(ob.design_sosmat_band_passes(2, edges, 44100)))


def test_bilinear_sos_errors():
    from pyfilterbank_b200.sosfiltering import bilinear_sos
    with pytest.raises(Exception):
        bilinear_sos(np.ones((2, 3)), np.ones((2, 2)))
    with pytest.raises(Exception):
        bilinear_sos(np.ones((2, 2)), np.zeros((2, 2)))


def test_abi_exports_every_declared_symbol():
    from pyfilterbank_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "pfb_sos.h")).read()
    names = re.findall(r"PFB_API\s+[\w\s\*]+?\b(\w+)\s*\(", header)
    assert {"sosfilter", "sosfilter_double", "sosfilter_double_mimo", "pfb_bank_create", "pfb_bank_filter",
            "pfb_bank_filter_host"} <= set(names)
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in L.pfb_version()


def test_tma_kernels_wait_converged():
    """Every kernel whose consumer warps wait on an mbarrier does so through the vote-terminated loop of
    mbar_wait_warp (pfb_tma.cuh): all lanes test the phase, a warp vote decides uniformly whether to poll again, so
    the warp stays one converged group.  The earlier forms failed on the GPU: one polling lane plus __syncwarp() left
    the warp split in two groups that issued the following arithmetic twice (profiles/r02l_warp_prof.txt); a
    branch-inside-asm spin made ptxas assume convergence that did not hold (profiles/r01h_fos_race.txt).
    The SASS must therefore contain a VOTE next to the phase checks, and no YIELD-ing single-lane spin in between."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from pyfilterbank_b200 import _lib
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    waits, votes, name = {}, {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
        elif "SYNCS.PHASECHK" in line:
            waits[name] = waits.get(name, 0) + 1
        elif "VOTE" in line:
            votes[name] = votes.get(name, 0) + 1
    assert len(waits) > 100                      # all instantiations of the TMA kernels
    bad = [k for k in waits if votes.get(k, 0) < 1]
    assert not bad, bad[:3]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pyfilterbank_b200 as pfb
    with pytest.raises(pfb.PfbError):
        pfb.sosfilter_double_c(np.zeros(8), np.array([[1, 0, 0, 1, 0, 0.]]))
    h = ctypes.c_void_p()
    from pyfilterbank_b200 import _lib
    rc = _lib.lib().pfb_bank_create(ctypes.byref(h), np.array([1, 0, 0, 1, 0, 0.]).ctypes.data, 1, 1, 1, 0)
    assert rc == -3 and b"no CPU fallback" in _lib.lib().pfb_last_error()


def test_design_and_adapter_abi_validate_arguments():
    """pfb_design_butter_sos / pfb_states_convert / pfb_bank_levels_weighted reject bad arguments before touching a device
    (the reference raises for the same inputs: butterworth.py:44-47, :89); without a device the design call reports
    PFB_ERR_NO_DEVICE instead of computing on the host."""
    import torch
    from pyfilterbank_b200 import _lib
    L = _lib.lib()
    v1, v2, out = np.array([0.1]), np.array([0.2]), np.empty((1, 4, 6))
    assert L.pfb_design_butter_sos(2, 3, v1.ctypes.data, v2.ctypes.data, 1, out.ctypes.data, None, None) == -1   # odd L
    assert L.pfb_design_butter_sos(7, 4, v1.ctypes.data, v2.ctypes.data, 1, out.ctypes.data, None, None) == -1   # kind
    assert L.pfb_design_butter_sos(2, 4, v1.ctypes.data, None, 1, out.ctypes.data, None, None) == -1             # band-pass needs v2
    assert L.pfb_design_butter_sos(2, 4, v1.ctypes.data, v2.ctypes.data, 1, None, None, None) == -1              # no output
    assert L.pfb_states_convert(None, out.ctypes.data, 1, 1, 4, 0, 1, None) == -1
    assert L.pfb_bank_levels_weighted(None, 1, None, 0, None, 0, None, None, 0, None, 0, None, 0, 0, 0, 0, None) == -1
    if not torch.cuda.is_available():
        assert L.pfb_design_butter_sos(2, 4, v1.ctypes.data, v2.ctypes.data, 1, out.ctypes.data, None, None) == -3
        assert b"no CUDA device" in L.pfb_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyfilterbank_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), (f, "the product path must not reference the CPU checker")


def test_shard_channels():
    from pyfilterbank_b200.sharding import shard_channels
    for C in (1, 7, 64, 8192):
        for W in (1, 2, 3, 8):
            spans = [shard_channels(C, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from pyfilterbank_b200.sharding import shard_channels, gather_bands
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
C, B, n = 5, 3, 16                      # ragged: rank 0 gets 3 channels, rank 1 gets 2
full = torch.arange(C * B * n, dtype=torch.float64).reshape(C, B, n)
c0, c1 = shard_channels(C, world, rank)
got = gather_bands(full[c0:c1].clone(), C)
assert torch.equal(got, full), "gathered slab differs"
got2 = gather_bands(full[c0:c1].clone())          # sizes discovered by a collective
assert torch.equal(got2, full)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gather_bands_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` needs no GPU: one JSON line on stdout with the driver's keys, timed on the
    unmodified reference C (oracle/_ref) or, where that is not built, on the port."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "band-samples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("1/3-octave filter_mimo_c bank, 64 channels x 60 s")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_flag_words_and_auto_arithmetic_abi():
    """Python flag words match include/pfb_sos.h; arith64='auto' asks for float64 states AND the per-band choice;
    pfb_bank_f32_bands validates its argument and needs a device (no host-side arithmetic in the product)."""
    import re
    import torch
    from pyfilterbank_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pfb_sos.h")).read()
    words = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define (PFB_[A-Z0-9_]+) (0x[0-9a-fA-F]+)", hdr)}
    assert words["PFB_EXACT"] == _lib.EXACT and words["PFB_ARITH_F64"] == _lib.ARITH_F64
    assert words["PFB_DECIMATE2"] == _lib.DECIMATE2 and words["PFB_ARITH_AUTO"] == _lib.ARITH_AUTO
    assert _lib.flags_of("auto", False, "auto") == _lib.ARITH_F64 | _lib.ARITH_AUTO
    assert _lib.flags_of("scan", False, True) == _lib.MODE_SCAN | _lib.ARITH_F64
    L = _lib.lib()
    assert L.pfb_bank_f32_bands(None) == -1
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert L.pfb_bank_create(ctypes.byref(h), np.array([1, 0, 0, 1, 0, 0.]).ctypes.data, 1, 1, 1, 0) == -3
