"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the CPU checker (`oracle/liboracle.so`, this repo's restatement of
pyfilterbank/sosfilt.c) and, when present, for the unmodified reference C file compiled into
`oracle/_ref/libsosfilt_ref.so`.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module.  Nothing under
`pyfilterbank_b200/` imports it; the product path has no CPU fallback.

Parity status: pinned.  The restatement is bit-exact against the reference's own C
(tests/test_oracle.py::test_restatement_equals_reference_c) and reproduces the golden vectors
made by importing the Python reference (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libsosfilt_ref.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_int = ctypes.c_int


def build(force=False):
    """Compile the checker with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_ORACLE_SO) or (
            os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "sos_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/pyfilterbank/sosfilt.c") and (force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_ORACLE_SO)
        L.oracle_sosfilter.argtypes = [_f32p, _i64, _f32p, _int, _f32p]
        L.oracle_sosfilter_double.argtypes = [_f64p, _i64, _f64p, _int, _f64p]
        L.oracle_sosfilter_double_mimo.argtypes = [_f64p, _i64, _int, _f64p, _int, _int, _f64p]
        L.oracle_sos_bank_f32.argtypes = [_f32p, _i64, _f32p, _i64, _i64, _int, _int, _int, _f32p, _int, _f32p]
        L.oracle_sos_bank_f64.argtypes = [_f64p, _i64, _f64p, _i64, _i64, _int, _int, _int, _f64p, _int, _f64p]
        L.oracle_fos_cascade_c128.argtypes = [_f64p, _i64, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_double, _int, _f64p, _f64p]
        L.oracle_tf_df2t_f64.argtypes = [_f64p, _f64p, _i64, _f64p, _f64p, _int, _f64p]
        for f in ("oracle_sosfilter", "oracle_sosfilter_double", "oracle_sosfilter_double_mimo",
                  "oracle_sos_bank_f32", "oracle_sos_bank_f64", "oracle_fos_cascade_c128",
                  "oracle_tf_df2t_f64"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(_REF_SO)


def ref():
    """The reference's own sosfilt.c, compiled unmodified (three symbols, build.py:5-7)."""
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        R = ctypes.CDLL(_REF_SO)
        R.sosfilter.argtypes = [_f32p, _int, _f32p, _int, _f32p]
        R.sosfilter_double.argtypes = [_f64p, _int, _f64p, _int, _f64p]
        R.sosfilter_double_mimo.argtypes = [_f64p, _int, _int, _f64p, _int, _int, _f64p]
        for f in ("sosfilter", "sosfilter_double", "sosfilter_double_mimo"):
            getattr(R, f).restype = None
        _ref = R
    return _ref


def _p(a, t):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------------------------------------
# Restatements of the reference's Python wrappers (sosfiltering.py:54-248).  They return fresh
# arrays and never mutate the caller's input, like the reference.
# ---------------------------------------------------------------------------------------------

def sosfilter_c(signal, sos, states=None, use_ref=False):
    """sosfiltering.py:54-108 (float32; coefficients and states are cast to float32 too)."""
    sig = np.array(signal, dtype=np.float32).flatten()
    s = np.array(sos, dtype=np.float32).flatten()
    nsamp = int(len(signal))
    ksos = s.size // 6
    st = np.zeros(2 * ksos, np.float32) if states is None else np.array(states, dtype=np.float32).flatten()
    if use_ref:
        ref().sosfilter(_p(sig, _f32p), nsamp, _p(s, _f32p), ksos, _p(st, _f32p))
    else:
        lib().oracle_sosfilter(_p(sig, _f32p), nsamp, _p(s, _f32p), ksos, _p(st, _f32p))
    return sig[:nsamp], st


def sosfilter_double_c(signal, sos, states=None, use_ref=False):
    """sosfiltering.py:111-163."""
    sig = np.array(signal, dtype=np.float64).flatten()
    s = np.array(sos, dtype=np.float64).flatten()
    nsamp = int(len(signal))
    ksos = s.size // 6
    st = np.zeros(2 * ksos) if states is None else np.array(states, dtype=np.float64).flatten()
    if use_ref:
        ref().sosfilter_double(_p(sig, _f64p), nsamp, _p(s, _f64p), ksos, _p(st, _f64p))
    else:
        lib().oracle_sosfilter_double(_p(sig, _f64p), nsamp, _p(s, _f64p), ksos, _p(st, _f64p))
    return sig[:nsamp], st


def sosfilter_double_mimo_c(signal, sos, states=None, use_ref=False):
    """sosfiltering.py:166-248, same marshalling: sos flattened 'F' and tiled per channel,
    states flattened 'F', signal tiled over bands and laid out [C][B][N]; returns
    (out (N, B, C) Fortran-ordered, states flat)."""
    signal = np.asarray(signal)
    sos = np.asarray(sos)
    nframes = int(signal.shape[0])
    nchan = int(signal.shape[1]) if signal.ndim > 1 else 1
    ksos = int(sos.shape[0] // 6)
    if sos.ndim > 1:
        kbands = int(sos.shape[1])
        if sos.ndim == 2 and nchan > 1:
            sos = np.tile(sos.flatten('F'), nchan)
    else:
        kbands = 1
    if states is None:
        states = np.zeros(nchan * kbands * ksos * 2)
    st = np.array(states, dtype=np.float64).flatten('F')
    s = np.array(sos, dtype=np.float64).flatten('F')
    x = np.array(signal, dtype=np.float64).reshape(nframes, nchan)
    buf = np.ascontiguousarray(np.repeat(x.T[:, None, :], kbands, axis=1)).reshape(-1)  # [C][B][N]
    if use_ref:
        ref().sosfilter_double_mimo(_p(buf, _f64p), nframes, nchan, _p(s, _f64p), ksos, kbands, _p(st, _f64p))
    else:
        lib().oracle_sosfilter_double_mimo(_p(buf, _f64p), nframes, nchan, _p(s, _f64p), ksos, kbands,
                                           _p(st, _f64p))
    return buf.reshape((nframes, kbands, nchan), order='F'), st


def sos_bank(x, sos, states=None, dtype=np.float64):
    """Bank form on [C][N] input: y [C][B][N].  sos [B][K][6] or [C][B][K][6]; states
    [C][B][K][2].  Per (c, b) the arithmetic is exactly the reference cascade."""
    x = np.ascontiguousarray(x, dtype=dtype)
    C, N = x.shape
    sos = np.ascontiguousarray(sos, dtype=dtype)
    per_channel = sos.ndim == 4
    B, K = (sos.shape[1], sos.shape[2]) if per_channel else (sos.shape[0], sos.shape[1])
    st = np.zeros((C, B, K, 2), dtype) if states is None else np.array(states, dtype=dtype).reshape(C, B, K, 2)
    y = np.empty((C, B, N), dtype)
    if dtype == np.float64:
        lib().oracle_sos_bank_f64(_p(x, _f64p), N, _p(y, _f64p), N, N, C, B, K, _p(sos, _f64p),
                                  int(per_channel), _p(st, _f64p))
    else:
        lib().oracle_sos_bank_f32(_p(x, _f32p), N, _p(y, _f32p), N, N, C, B, K, _p(sos, _f32p),
                                  int(per_channel), _p(st, _f32p))
    return y, st


def fosfilter(b, a, order, signal, states=None):
    """gammatone.py:196-237 for a real 1-D signal: returns (complex128 band, complex128 states)."""
    x = np.ascontiguousarray(np.real(signal), dtype=np.float64)
    n = x.shape[0]
    coef = complex(-np.asarray(a)[1])
    gain = float(np.real(np.asarray(b)[0]))
    st = np.zeros(order, np.complex128) if states is None else np.array(states, dtype=np.complex128)
    out = np.empty(n, np.complex128)
    stv = st.view(np.float64)
    lib().oracle_fos_cascade_c128(_p(x, _f64p), n, gain, coef.real, coef.imag, order,
                                  _p(out.view(np.float64), _f64p), _p(stv, _f64p))
    return out, st


def lfilter_df2t(b, a, x, zi=None):
    """scipy.signal.lfilter(b, a, x) for 1-D real x with a[0] == 1 (splweighting.py:80)."""
    b = np.asarray(b, np.float64)
    a = np.asarray(a, np.float64)
    nt = max(len(b), len(a))
    bb = np.zeros(nt); bb[:len(b)] = b / a[0]
    aa = np.zeros(nt); aa[:len(a)] = a / a[0]
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    z = np.zeros(max(nt - 1, 1)) if zi is None else np.array(zi, np.float64)
    lib().oracle_tf_df2t_f64(_p(x, _f64p), _p(y, _f64p), x.shape[0], _p(bb, _f64p), _p(aa, _f64p), nt, _p(z, _f64p))
    return y, z


def sosfilter_py(x, sos, states=None):
    """The reference's `sosfilter_py` (sosfiltering.py:330-379): `lfilter(b, a, x, 0, zi)` section by section, states a
    dict {section index: zi (2,)} of Direct-Form-II-Transposed delay values.  x 1-D float64."""
    sos = np.asarray(sos, np.float64).reshape(-1, 6)
    n = sos.shape[0]
    if states is None:
        states = {i: np.zeros(2) for i in range(n)}
    x = np.asarray(x, np.float64)
    for i in range(n):
        x, zi = lfilter_df2t(sos[i, :3], sos[i, 3:], x, states[i])
        states[i] = zi
    return x, states


# ---------------------------------------------------------------------------------------------
# Composed paths (no single reference function exists; composed from the pieces above exactly like
# tests/golden/make_golden_composed.py composes the reference's own functions).
# ---------------------------------------------------------------------------------------------

def decimated_bank(x, plan, aa_sos, order):
    """Decimating band path: x (N, C); plan = list of stages {'m', 'bands', 'sosmat'} (stage 0 first);
    stage m signal = [::2] of the anti-alias cascade `aa_sos` applied to the stage m-1 signal
    (sosfilter_double arithmetic).  Returns a list over bands of (N_b, C) arrays."""
    x = np.asarray(x, np.float64).reshape(len(x), -1)
    C = x.shape[1]
    nb = sum(len(s['bands']) for s in plan)
    out = [None] * nb
    for c in range(C):
        stage = x[:, c].copy()
        for st in plan:
            if st['m'] > 0:
                stage, _ = sosfilter_double_c(stage, aa_sos)
                stage = stage[::2].copy()
            for j, b in enumerate(st['bands']):
                y, _ = sosfilter_double_c(stage, st['sosmat'][:, j].reshape(order, 6))
                if out[b] is None:
                    out[b] = np.zeros((len(y), C))
                out[b][:, c] = y
    return out


def weighted_bank(x, b, a, sosmat):
    """lfilter(b, a, .) per channel (splweighting.py:80) followed by sosfilter_double_mimo_c
    (octbank.py:412-434).  x (N, C) -> (weighted (N, C), out (N, B, C), states flat)."""
    x = np.asarray(x, np.float64).reshape(len(x), -1)
    w = np.stack([lfilter_df2t(b, a, x[:, c])[0] for c in range(x.shape[1])], axis=1)
    y, st = sosfilter_double_mimo_c(w, sosmat)
    return w, y, st
