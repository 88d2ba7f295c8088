"""A-, B- and C-weighting of sound pressure level signals (IEC 61672) on the GPU.

Mirrors pyfilterbank/splweighting.py:61-211: `weight_signal(data, sample_rate, weighting)` and the
three coefficient designs.  The analog prototypes are products of first- and second-order pole
factors at the standard's corner frequencies, mapped to (b, a) with `scipy.signal.bilinear` exactly
as the reference does (scipy is the reference's own third-party dependency for this step); the
filtering itself -- `scipy.signal.lfilter(b, a, data)` in the reference -- runs in libpfb_sos.so as
the same direct form II transposed recurrence (pfb_tf_filter_f64), one row per GPU lane.
"""
import ctypes

import numpy as np
import torch
from scipy.signal import bilinear

from . import _lib
from .sosfiltering import _device

# corner frequencies (Hz) and 1 kHz normalisation gains (dB) of the weighting curves
_F1, _F2, _F3, _F4 = 20.598997, 107.65265, 737.86223, 12194.217
_F2_B = 158.5
_GAIN_DB = {'A': 1.9997, 'B': 0.17, 'C': 0.0619}


def _double_pole(f):
    w = 2 * np.pi * f
    return [1., 2 * w, w ** 2]


def _design(weighting, sample_rate):
    w4 = 2 * np.pi * _F4
    den = np.convolve(_double_pole(_F4), _double_pole(_F1))
    if weighting == 'A':
        den = np.convolve(np.convolve(den, [1., 2 * np.pi * _F3]), [1., 2 * np.pi * _F2])
        zeros_at_dc = 4
    elif weighting == 'B':
        den = np.convolve(den, [1., 2 * np.pi * _F2_B])
        zeros_at_dc = 3
    elif weighting == 'C':
        zeros_at_dc = 2
    else:
        raise KeyError(weighting)
    num = [w4 ** 2 * 10 ** (_GAIN_DB[weighting] / 20.0)] + [0.] * zeros_at_dc
    return bilinear(num, den, sample_rate)


def a_weighting_coeffs_design(sample_rate):
    """(b, a) of the digital A-weighting filter (7 taps each); splweighting.py:82-126."""
    return _design('A', sample_rate)


def b_weighting_coeffs_design(sample_rate):
    """(b, a) of the digital B-weighting filter (6 taps); splweighting.py:128-170."""
    return _design('B', sample_rate)


def c_weighting_coeffs_design(sample_rate):
    """(b, a) of the digital C-weighting filter (5 taps); splweighting.py:173-211."""
    return _design('C', sample_rate)


_weighting_coeff_design_funsd = {'A': a_weighting_coeffs_design, 'B': b_weighting_coeffs_design,
                                 'C': c_weighting_coeffs_design}


def tf_filter(b, a, x, zi=None, out=None):
    """lfilter(b, a, x) along the last axis of a CUDA float64 tensor; returns (y, zf) with zf
    (rows, ntaps-1) the final delay values (scipy's zi convention, updated in place when `zi` is given).
    A 2-D x may be a time slice of a larger tensor (row stride > length); `out` likewise."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    shape = x.shape
    n = shape[-1]
    x2 = x if (x.dim() == 2 and x.stride(1) == 1) else x.reshape(-1, n).contiguous()
    R = x2.shape[0]
    nz = max(len(b), len(a)) - 1
    z = torch.zeros((R, max(nz, 1)), dtype=torch.float64, device=x.device) if zi is None else zi
    y = torch.empty((R, n), dtype=torch.float64, device=x.device) if out is None else out
    if n > 0:
        ldx = x2.stride(0) if R > 1 else max(n, 1)
        ldy = y.stride(0) if R > 1 else max(n, 1)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().pfb_tf_filter_f64(x2.data_ptr(), ldx, y.data_ptr(), ldy, b.ctypes.data, len(b),
                                                     a.ctypes.data, len(a), z.data_ptr(), n, R,
                                                     torch.cuda.current_stream().cuda_stream))
    return (y.reshape(shape) if out is None else y), z


def weight_signal(data, sample_rate=44100, weighting='A'):
    """Filter `data` (numpy array or CUDA tensor, last axis = time) with the A-, B- or C-weighting
    filter; zero initial state, like the reference's `lfilter(b, a, data)` (splweighting.py:61-80)."""
    b, a = _weighting_coeff_design_funsd[weighting](sample_rate)
    if isinstance(data, torch.Tensor) and data.is_cuda:
        return tf_filter(b, a, data.to(torch.float64))[0]
    x = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).to(_device())
    from .sosfiltering import _to_host
    return _to_host(tf_filter(b, a, x)[0])
