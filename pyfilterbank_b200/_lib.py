"""ctypes binding of libpfb_sos.so (include/pfb_sos.h).  There is no CPU fallback: if the
CUDA library is missing or no device is present, calls raise."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFB_SOS_LIB") or os.path.join(_HERE, "libpfb_sos.so")

PFB_F32, PFB_F64 = 0, 1
MODE_AUTO, MODE_SEQ, MODE_SCAN = 0, 1, 2
EXACT, ARITH_F64, DECIMATE2, DEC_PHASE1, ARITH_AUTO = 0x10, 0x20, 0x40, 0x80, 0x200
_MODES = {"auto": MODE_AUTO, "seq": MODE_SEQ, "scan": MODE_SCAN}


class PfbError(RuntimeError):
    pass


class LaunchInfo(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("kernels", ctypes.c_int), ("warps_per_cascade", ctypes.c_int),
                ("chunk", ctypes.c_int64), ("warm_total", ctypes.c_int64), ("aligned", ctypes.c_int)]


class StreamInfo(ctypes.Structure):
    _fields_ = [("blocks", ctypes.c_int64), ("samples", ctypes.c_int64), ("device_allocs_total", ctypes.c_int64),
                ("device_allocs_after_first_block", ctypes.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -m pyfilterbank_b200._build` "
                "(pyfilterbank_b200 has no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.pfb_bank_create.argtypes = [ctypes.POINTER(vp), vp, ci, ci, ci, ci]
        L.pfb_bank_create.restype = ci
        L.pfb_bank_destroy.argtypes = [vp]
        L.pfb_bank_destroy.restype = None
        L.pfb_bank_filter.argtypes = [vp, ci, vp, i64, vp, i64, vp, i64, ci, ci, vp]
        L.pfb_bank_filter.restype = ci
        L.pfb_bank_levels.argtypes = [vp, ci, vp, i64, vp, i64, vp, i64, ci, i64, ci, vp]
        L.pfb_bank_levels.restype = ci
        L.pfb_bank_levels_weighted.argtypes = [vp, ci, vp, i64, vp, i64, vp, vp, ci, vp, ci, vp, i64, ci, i64, ci, vp]
        L.pfb_bank_levels_weighted.restype = ci
        L.pfb_bank_levels_host.argtypes = [vp, ci, vp, i64, vp, i64, vp, i64, ci, i64, ci]
        L.pfb_bank_levels_host.restype = ci
        L.pfb_bank_filter_host.argtypes = [vp, ci, vp, i64, vp, i64, vp, i64, ci, ci]
        L.pfb_bank_filter_host.restype = ci
        L.pfb_bank_last_launch.argtypes = [vp, ctypes.POINTER(LaunchInfo)]
        L.pfb_bank_last_launch.restype = ci
        L.pfb_bank_timing_begin.argtypes = [vp, ci]
        L.pfb_bank_timing_begin.restype = ci
        L.pfb_bank_timing_end.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ci)]
        L.pfb_bank_timing_end.restype = ci
        L.pfb_bank_f32_bands.argtypes = [vp]
        L.pfb_bank_f32_bands.restype = ci
        L.pfb_bank_set_warm_tol.argtypes = [vp, ctypes.c_double]
        L.pfb_bank_set_warm_tol.restype = ci
        L.pfb_fos_bank_c128.argtypes = [vp, i64, vp, i64, vp, ci, ci, vp, i64, ci, vp]
        L.pfb_fos_bank_c128.restype = ci
        L.pfb_gt_synthesize_c128.argtypes = [vp, i64, vp, i64, vp, vp, ci, ci, i64, ci, vp]
        L.pfb_gt_synthesize_c128.restype = ci
        L.pfb_tf_filter_f64.argtypes = [vp, i64, vp, i64, vp, ci, vp, ci, vp, i64, ci, vp]
        L.pfb_tf_filter_f64.restype = ci
        L.pfb_bank_filter_weighted.argtypes = [vp, ci, vp, i64, vp, i64, vp, vp, ci, vp, ci, vp, i64, ci, ci, vp]
        L.pfb_bank_filter_weighted.restype = ci
        L.pfb_bank_release_stream.argtypes = [vp, vp]
        L.pfb_bank_release_stream.restype = ci
        L.pfb_device_allocs.restype = ctypes.c_longlong
        L.pfb_design_butter_sos.argtypes = [ci, ci, vp, vp, ci, vp, vp, vp]
        L.pfb_design_butter_sos.restype = ci
        L.pfb_states_convert.argtypes = [vp, vp, i64, ci, ci, ci, ci, vp]
        L.pfb_states_convert.restype = ci
        L.pfb_decbank_create.argtypes = [ctypes.POINTER(vp), ci, ctypes.POINTER(ci), ctypes.POINTER(vp), ci, vp, ci]
        L.pfb_decbank_create.restype = ci
        L.pfb_decbank_destroy.argtypes = [vp]
        L.pfb_decbank_destroy.restype = None
        L.pfb_decbank_filter.argtypes = [vp, vp, i64, ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(vp), vp,
                                         ctypes.POINTER(ci), i64, ci, ci, vp]
        L.pfb_decbank_filter.restype = ci
        L.pfb_decbank_info.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.pfb_decbank_info.restype = ci
        L.pfb_stream_create.argtypes = [ctypes.POINTER(vp), vp, ci, ci, i64, ci]
        L.pfb_stream_create.restype = ci
        L.pfb_stream_push.argtypes = [vp, vp, i64, vp, i64, i64]
        L.pfb_stream_push.restype = ci
        L.pfb_stream_sync.argtypes = [vp]
        L.pfb_stream_sync.restype = ci
        L.pfb_stream_wait.argtypes = [vp, i64]
        L.pfb_stream_wait.restype = ci
        L.pfb_stream_states.argtypes = [vp, vp, ci]
        L.pfb_stream_states.restype = ci
        L.pfb_stream_stats.argtypes = [vp, ctypes.POINTER(StreamInfo)]
        L.pfb_stream_stats.restype = ci
        L.pfb_stream_destroy.argtypes = [vp]
        L.pfb_stream_destroy.restype = None
        L.pfb_last_error.restype = ctypes.c_char_p
        L.pfb_last_status.restype = ci
        L.pfb_version.restype = ctypes.c_char_p
        fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
        L.sosfilter.argtypes = [fp, ci, fp, ci, fp]
        L.sosfilter_double.argtypes = [dp, ci, dp, ci, dp]
        L.sosfilter_double_mimo.argtypes = [dp, ci, ci, dp, ci, ci, dp]
        for f in (L.sosfilter, L.sosfilter_double, L.sosfilter_double_mimo):
            f.restype = None
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PfbError("libpfb_sos error %d: %s" % (rc, lib().pfb_last_error().decode()))


def flags_of(mode="auto", exact=False, arith64=False, decimate=None):
    f = _MODES[mode]
    if exact:
        f |= EXACT
    if arith64 == "auto":             # float64 states, arithmetic chosen per band (PFB_ARITH_AUTO)
        f |= ARITH_F64 | ARITH_AUTO
    elif arith64:
        f |= ARITH_F64
    if decimate is not None:          # phase 0 / 1 of the decimate-by-2 store
        f |= DECIMATE2 | (DEC_PHASE1 if decimate else 0)
    return f


class Bank:
    """Device-resident coefficient tables of one filter bank (pfb_bank_create).

    sos: (B, K, 6) shared by all channels, or (C, B, K, 6) per channel; rows are
    [b0 b1 b2 a0 a1 a2] and a0 is ignored like in sosfilt.c:39."""

    def __init__(self, sos):
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        if sos.ndim == 3:
            self.C_sos, per_channel = 1, 0
            self.B, self.K = sos.shape[0], sos.shape[1]
        elif sos.ndim == 4:
            self.C_sos, per_channel = sos.shape[0], 1
            self.B, self.K = sos.shape[1], sos.shape[2]
        else:
            raise ValueError("sos must be (B, K, 6) or (C, B, K, 6)")
        if sos.shape[-1] != 6:
            raise ValueError("sos rows must hold 6 coefficients")
        self.per_channel = per_channel
        h = ctypes.c_void_p()
        check(lib().pfb_bank_create(ctypes.byref(h), sos.ctypes.data, self.C_sos, self.B, self.K, per_channel))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.pfb_bank_destroy(h)

    def filter_ptr(self, dtype, x_ptr, ldx, y_ptr, ldy, states_ptr, n, C, flags=0, stream=0):
        check(lib().pfb_bank_filter(self._h, dtype, x_ptr, ldx, y_ptr, ldy, states_ptr, n, C, flags, stream))

    def filter_weighted_ptr(self, dtype, x_ptr, ldx, y_ptr, ldy, states_ptr, b, a, wz_ptr, n, C, flags=0, stream=0):
        """Weighting prefilter lfilter(b, a, .) + bank as one call (pfb_bank_filter_weighted); b, a numpy float64."""
        check(lib().pfb_bank_filter_weighted(self._h, dtype, x_ptr, ldx, y_ptr, ldy, states_ptr, b.ctypes.data, len(b),
                                             a.ctypes.data, len(a), wz_ptr, n, C, flags, stream))

    def levels_ptr(self, dtype, x_ptr, ldx, lev_ptr, ld_lev, states_ptr, n, C, block_len, flags=0, stream=0):
        check(lib().pfb_bank_levels(self._h, dtype, x_ptr, ldx, lev_ptr, ld_lev, states_ptr, n, C, block_len, flags, stream))

    def levels_weighted_ptr(self, dtype, x_ptr, ldx, lev_ptr, ld_lev, states_ptr, b, a, wz_ptr, n, C, block_len, flags=0, stream=0):
        b = np.ascontiguousarray(b, dtype=np.float64)
        a = np.ascontiguousarray(a, dtype=np.float64)
        check(lib().pfb_bank_levels_weighted(self._h, dtype, x_ptr, ldx, lev_ptr, ld_lev, states_ptr, b.ctypes.data, len(b),
                                             a.ctypes.data, len(a), wz_ptr, n, C, block_len, flags, stream))

    def levels_host(self, dtype, x, ldx, levels, ld_lev, states, n, C, block_len, flags=0):
        check(lib().pfb_bank_levels_host(self._h, dtype, x.ctypes.data, ldx, levels.ctypes.data, ld_lev,
                                         states.ctypes.data, n, C, block_len, flags))

    def filter_host(self, dtype, x, ldx, y, ldy, states, n, C, flags=0):
        check(lib().pfb_bank_filter_host(self._h, dtype, x.ctypes.data, ldx, y.ctypes.data, ldy,
                                         states.ctypes.data, n, C, flags))

    def f32_bands(self):
        """number of bands that arith64='auto' runs in float32 arithmetic (calibrated by the library on first use)"""
        n = lib().pfb_bank_f32_bands(self._h)
        if n < 0:
            check(n)
        return n

    def set_warm_tol(self, tol):
        """Relative-L2 budget of the truncated pre-pass of the time-chunked kernel (0: exact)."""
        check(lib().pfb_bank_set_warm_tol(self._h, float(tol)))

    def timing_begin(self, max_calls):
        check(lib().pfb_bank_timing_begin(self._h, int(max_calls)))

    def timing_end(self):
        """Mean device ms per call: {prepass, carry, output, total, calls} since timing_begin."""
        ms = (ctypes.c_double * 4)()
        calls = ctypes.c_int()
        check(lib().pfb_bank_timing_end(self._h, ms, ctypes.byref(calls)))
        return {"prepass": ms[0], "carry": ms[1], "output": ms[2], "total": ms[3], "calls": calls.value}

    def last_launch(self):
        info = LaunchInfo()
        check(lib().pfb_bank_last_launch(self._h, ctypes.byref(info)))
        return {k: getattr(info, k) for k, _ in LaunchInfo._fields_}


class DecBank:
    """Decimating band path as one C call (pfb_decbank_*): `stage_sos[m]` is (B_m, K, 6) or None for a stage without
    bands, `aa_sos` (Kaa, 6) the anti-alias cascade applied before every halving."""

    def __init__(self, stage_sos, K, aa_sos):
        self.nstages = len(stage_sos)
        self.K = int(K)
        self.nb = [0 if s is None else int(s.shape[0]) for s in stage_sos]
        self._keep = [None if s is None else np.ascontiguousarray(s, dtype=np.float64) for s in stage_sos]
        aa = np.ascontiguousarray(aa_sos, dtype=np.float64).reshape(-1, 6)
        nbands = (ctypes.c_int * self.nstages)(*self.nb)
        ptrs = (ctypes.c_void_p * self.nstages)(*[None if s is None else s.ctypes.data for s in self._keep])
        h = ctypes.c_void_p()
        check(lib().pfb_decbank_create(ctypes.byref(h), self.nstages, nbands, ptrs, self.K, aa.ctypes.data, aa.shape[0]))
        self._h = h
        self.Kp = self.info()["aa_sections_padded"]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.pfb_decbank_destroy(h)

    def info(self):
        v = [ctypes.c_int() for _ in range(4)]
        check(lib().pfb_decbank_info(self._h, *[ctypes.byref(x) for x in v]))
        return {"stages": v[0].value, "aa_sections_padded": v[1].value, "last_kernels": v[2].value,
                "last_fused_stages": v[3].value}

    def filter_ptr(self, x_ptr, ldx, y_ptrs, ldys, state_ptrs, aa_states_ptr, phases, n, C, flags=0, stream=0):
        """phases: list of ints (decimator phase per halving stage), updated in place."""
        S = self.nstages
        yp = (ctypes.c_void_p * S)(*y_ptrs)
        ld = (ctypes.c_int64 * S)(*ldys)
        sp = (ctypes.c_void_p * S)(*state_ptrs)
        ph = (ctypes.c_int * max(S - 1, 1))(*(list(phases) + [0])[:max(S - 1, 1)])
        check(lib().pfb_decbank_filter(self._h, x_ptr, ldx, yp, ld, sp, aa_states_ptr, ph, n, C, flags, stream))
        for i in range(S - 1):
            phases[i] = int(ph[i])


class Stream:
    """Streaming session over host buffers (pfb_stream_*): device-resident states, asynchronous pushes."""

    def __init__(self, bank, dtype, C, max_block, flags=0):
        self.bank = bank                      # keeps the bank alive
        h = ctypes.c_void_p()
        check(lib().pfb_stream_create(ctypes.byref(h), bank._h, dtype, C, max_block, flags))
        self._h = h

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.pfb_stream_destroy(h)

    __del__ = close

    def push(self, x_ptr, ldx, y_ptr, ldy, n):
        check(lib().pfb_stream_push(self._h, x_ptr, ldx, y_ptr, ldy, n))

    def sync(self):
        check(lib().pfb_stream_sync(self._h))

    def wait(self, block):
        check(lib().pfb_stream_wait(self._h, block))

    def get_states(self, out):
        check(lib().pfb_stream_states(self._h, out.ctypes.data, 0))
        return out

    def set_states(self, arr):
        check(lib().pfb_stream_states(self._h, arr.ctypes.data, 1))

    def stats(self):
        info = StreamInfo()
        check(lib().pfb_stream_stats(self._h, ctypes.byref(info)))
        return {k: getattr(info, k) for k, _ in StreamInfo._fields_}
