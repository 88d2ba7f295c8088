"""Butterworth filters in cascade (second-order-section) form, host side (numpy).

Same public functions as the reference's pyfilterbank/butterworth.py:16-146 (`butter_sos`,
`butter_analog_sos`), following Stearns, "Digital Signal Processing with Examples in MATLAB":
the analog filter is held as L/2 complex one-pole sections (d0*s + d1) / (c0*s + c1) whose
squared magnitudes become real biquads under the bilinear transform (`bilinear_sos`).  Coefficient
design runs once per bank on the host; its product, rows `[b0 b1 b2 a0 a1 a2]`, is what the CUDA
kernels consume.
"""
import numpy as np

from .sosfiltering import bilinear_sos

lowpass, highpass, bandpass, bandstop = 'lowpass', 'highpass', 'bandpass', 'bandstop'


def _split_quadratic(q2, q1, q0):
    """Roots of q2*r^2 + q1*r + q0 (arrays), '+' root first."""
    root = np.sqrt(q1 ** 2 - 4 * q2 * q0)
    return (-q1 + root) / (2 * q2), (-q1 - root) / (2 * q2)


def butter_analog_sos(band, L, w1, w2=0):
    """Analog Butterworth weights d, c (each (sections, 2), complex) for `band` in
    {'lowpass', 'highpass', 'bandpass', 'bandstop'}; L low-pass poles (even), critical
    frequencies w1 (< w2) in rad/s.  butterworth.py:61-146."""
    band = band.lower()
    if L % 2:
        raise Exception('Number of poles L must be even')
    if w1 <= 0:
        raise Exception('Frequency w1 must be in rad/s and >0')
    half = int(L / 2.0)
    wc = w1 if band in (lowpass, highpass) else w2 - w1
    k = np.arange(1, half + 1, dtype=np.double)
    poles = wc * np.exp(-1j * (2 * k + L - 1) * np.pi / (2 * L))

    # low-pass prototype sections: wc / (s - p)
    d = np.zeros((half, 2), dtype=complex)
    c = np.ones((half, 2), dtype=complex)
    d[:, 1] = wc
    c[:, 1] = -poles

    if band == lowpass:
        return d, c
    if band == highpass:            # s -> wc^2 / s
        d = d[:, ::-1].copy()
        c = c[:, ::-1].copy()
        c[:, 1] = c[:, 1] * wc ** 2
        return d, c
    if band == bandpass:            # s -> (s^2 + w1 w2) / s : every pole splits in two
        r1, r2 = _split_quadratic(c[:, 0], c[:, 1], c[:, 0] * w1 * w2)
        d_lo = np.stack([d[:, 1], np.zeros(half)], axis=1)            # wc * s
        d_hi = np.stack([np.zeros(half), np.ones(half)], axis=1)      # 1
        c_lo = np.stack([c[:, 0], -c[:, 0] * r1], axis=1)
        c_hi = np.stack([np.ones(half), -r2], axis=1)
        return np.concatenate([d_lo, d_hi]).astype(complex), np.concatenate([c_lo, c_hi]).astype(complex)
    if band == bandstop:            # s -> wc^2 s / (s^2 + w1 w2)
        def split(w):
            r1, r2 = _split_quadratic(w[:, 1], w[:, 0] * wc ** 2, w[:, 1] * w1 * w2)
            lo = np.stack([w[:, 1], -w[:, 1] * r1], axis=1)
            hi = np.stack([np.ones(half), -r2], axis=1)
            return np.concatenate([lo, hi]).astype(complex)
        return split(d), split(c)
    raise Exception("band must be one of 'lowpass', 'highpass', 'bandpass', 'bandstop'")


_KINDS = {lowpass: 0, highpass: 1, bandpass: 2, bandstop: 3}


def butter_sos_cuda(band, L, v1, v2=None, device_out=False):
    """`butter_sos` for B bands at once on the GPU (pfb_design_butter_sos, csrc/pfb_design.cu): v1, v2 arrays of
    critical frequencies in cycles per sample.  Returns (B, K, 6) rows in the reference's order (K = L for band-pass /
    band-stop, L / 2 otherwise) as a numpy array, or with device_out=True additionally as a CUDA tensor.  Raises
    PfbError for frequencies outside 0 < v1 < v2 < 0.5, like the reference's exceptions (butterworth.py:44-47)."""
    import torch
    from . import _lib
    kind = _KINDS[band.lower()]
    if L % 2:
        raise Exception('Number of poles L must be even')
    v1 = np.ascontiguousarray(np.atleast_1d(v1), dtype=np.float64)
    B = len(v1)
    v2a = None if kind < 2 else np.ascontiguousarray(np.atleast_1d(v2), dtype=np.float64)
    if v2a is not None and len(v2a) != B:
        raise ValueError("v1 and v2 must have the same length")
    K = L if kind >= 2 else L // 2
    out = np.empty((B, K, 6))
    dev_t = torch.empty((B, K, 6), dtype=torch.float64, device="cuda") if device_out else None
    _lib.check(_lib.lib().pfb_design_butter_sos(kind, L, v1.ctypes.data, None if v2a is None else v2a.ctypes.data, B,
                                                out.ctypes.data, None if dev_t is None else dev_t.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    return (out, dev_t) if device_out else out


def butter_sos(band, L, v1, v2=0.0):
    """Digital Butterworth cascade: rows [b0 b1 b2 a0 a1 a2] (a0 == 1), one per section; L/2
    sections for low/high-pass, L for band-pass/-stop.  v1, v2 are critical frequencies in
    cycles per sample (0 < v1 < v2 < 0.5).  butterworth.py:16-58."""
    if v1 <= 0 or v1 >= .5:
        raise Exception('Argument v1 must be >0.0 and <0.5')
    if v2 != 0 and (v2 <= v1 or v2 >= 0.5):
        raise Exception('Argument v2 must be >v1 and <0.5')
    if v2 == 0:
        d, c = butter_analog_sos(band, L, np.tan(np.pi * v1))
    else:
        d, c = butter_analog_sos(band, L, np.tan(np.pi * v1), np.tan(np.pi * v2))
    b, a = bilinear_sos(d, c)
    return np.concatenate((b, a), axis=1)[::-1].copy()   # the reference returns the sections reversed
