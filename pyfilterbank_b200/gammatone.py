"""Gammatone filter bank analysis on the GPU (Hohmann 2002), with the reference's interface.

Mirrors pyfilterbank/gammatone.py: the ERB-scale helpers (:44-141), `design_filter` (:144-193),
`fosfilter` (:196-237) and `GammatoneFilterbank.analyze` (:272-316).  Every band is `order` passes
of a complex one-pole  y[n] = g*x[n] + c*y[n-1]  (gain in the first pass only); the reference runs
them through scipy.signal.lfilter, here all bands (and all signals of a batch) run in one launch of
libpfb_sos.so (pfb_fos_bank_c128).  States use scipy's convention (c times the stage's last output),
so they can be carried between blocks -- which the reference's own code cannot do (its `if not
states:` raises on an array, gammatone.py:229).

The synthesis side (`reanalyze`, `synthesize`, `delay`) is host-side bookkeeping around the same
filter and is not part of the accelerated path.
"""
import numpy as np
import torch
from numpy import pi

from . import _lib
from .sosfiltering import _device

_ERB_L = 24.7
_ERB_Q = 9.265


def erb_count(centerfrequency):
    """Number of equivalent rectangular bandwidths below `centerfrequency` (Hz)."""
    return 21.4 * np.log10(4.37 * 0.001 * centerfrequency + 1)


def erb_aud(centerfrequency):
    """Equivalent rectangular bandwidth of the auditory filter at `centerfrequency` (Hohmann eq. 13)."""
    return _ERB_L + centerfrequency / _ERB_Q


def hertz_to_erbscale(frequency):
    """Hz -> ERB scale (Hohmann eq. 16)."""
    return _ERB_Q * np.log(1 + frequency / (_ERB_L * _ERB_Q))


def erbscale_to_hertz(erb):
    """ERB scale -> Hz (Hohmann eq. 17)."""
    return (np.exp(erb / _ERB_Q) - 1) * _ERB_L * _ERB_Q


def frequencies_gammatone_bank(start_band, end_band, norm_freq, density):
    """Centre frequencies `density` ERB apart, from start_band to end_band (exclusive) ERB around
    norm_freq; gammatone.py:118-141."""
    return erbscale_to_hertz(np.arange(start_band, end_band, density) + hertz_to_erbscale(norm_freq))


def design_filter(sample_rate=44100, order=4, centerfrequency=1000.0, band_width=None, band_width_factor=1.0,
                  attenuation_half_bandwidth_db=-3):
    """(b, a) = ([gain], [1, -coef]) of one gammatone stage; gammatone.py:144-193."""
    if band_width:
        phi = pi * band_width / sample_rate
    elif band_width_factor:
        phi = pi * band_width_factor * erb_aud(centerfrequency) / sample_rate
    else:
        raise ValueError('You need to specify either `band_width` or `band_width_factor!`')
    alpha = 10 ** (0.1 * attenuation_half_bandwidth_db / order)
    p = (-2 + 2 * alpha * np.cos(phi)) / (1 - alpha)
    lambda_ = -p / 2 - np.sqrt(p * p / 4 - 1)
    beta = 2 * pi * centerfrequency / sample_rate
    coef = lambda_ * np.exp(1j * beta)
    factor = 2 * (1 - np.abs(coef)) ** order
    return np.array([factor]), np.array([1., -coef])


def design_filtbank_coeffs(samplerate, order, centerfrequencies, bandwidths=None, bandwidth_factor=None,
                           attenuation_half_bandwidth_db=-3):
    for i, cf in enumerate(centerfrequencies):
        bw, bwf = (bandwidths[i], None) if bandwidths else (None, bandwidth_factor)
        yield design_filter(samplerate, order, cf, band_width=bw, band_width_factor=bwf,
                            attenuation_half_bandwidth_db=attenuation_half_bandwidth_db)


_MAX_ORDER_PER_CALL = 8     # cascade length of one pfb_fos_bank_c128 launch; longer cascades are chained


def _fos_bank_real(x, coef, order, states):
    """One library call: B complex one-pole cascades of `order` <= 8 stages over every row of the REAL (C, N) x."""
    C, N = x.shape
    B = coef.shape[0]
    y = torch.empty((C, B, N), dtype=torch.complex128, device=x.device)
    if N > 0:
        x = x.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().pfb_fos_bank_c128(x.data_ptr(), N, y.data_ptr(), N, coef.ctypes.data, B, order,
                                                     states.data_ptr(), N, C,
                                                     torch.cuda.current_stream().cuda_stream))
    return y


def fos_bank(x, coef, order, states=None):
    """Run B complex one-pole cascades over every row of x.

    x (C, N) CUDA tensor, float64 or complex128; coef (B, 3) numpy rows (gain, Re c, Im c); states (C, B, order)
    complex128 CUDA tensor or None (updated in place).  Returns (bands (C, B, N) complex128, states).

    The kernels take real rows and cascades of up to 8 stages.  The cascade is linear over the complex numbers,
    so a complex input is filtered as F(Re x) + i F(Im x) (the carried-in states go with the real part), which is
    what `reanalyze` needs (gammatone.py:318-321 feeds complex bands to lfilter); cascades of more than 8 stages
    run as 8 stages followed by the remaining stages with gain 1 on that complex intermediate, one band at a
    time.  Both compositions differ from the reference's single complex recurrence only in the rounding of the
    sums of separately rounded parts (~1e-16 relative)."""
    C, N = x.shape
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    B = coef.shape[0]
    if states is None:
        states = torch.zeros((C, B, order), dtype=torch.complex128, device=x.device)
    if order <= _MAX_ORDER_PER_CALL:
        if not x.is_complex():
            return _fos_bank_real(x.to(torch.float64), coef, order, states), states
        zero = torch.zeros_like(states)
        y = _fos_bank_real(x.real.to(torch.float64), coef, order, states)
        y = y + 1j * _fos_bank_real(x.imag.to(torch.float64), coef, order, zero)
        states += 1j * zero
        return y, states
    # long cascade: stages [0, 8) on the input, then the rest with gain 1, band by band (each band has its own input)
    head = states[:, :, :_MAX_ORDER_PER_CALL].contiguous()
    y, head = fos_bank(x, coef, _MAX_ORDER_PER_CALL, head)
    states[:, :, :_MAX_ORDER_PER_CALL] = head
    k0 = _MAX_ORDER_PER_CALL
    while k0 < order:
        k = min(_MAX_ORDER_PER_CALL, order - k0)
        for b in range(B):
            st = states[:, b:b + 1, k0:k0 + k].contiguous()
            unit = np.array([[1.0, coef[b, 1], coef[b, 2]]])
            yb, st = fos_bank(y[:, b], unit, k, st)
            y[:, b] = yb[:, 0]
            states[:, b:b + 1, k0:k0 + k] = st
        k0 += k
    return y, states


def _coef_rows(coeffs):
    return np.array([[np.real(b[0]), np.real(-a[1]), np.imag(-a[1])] for b, a in coeffs], dtype=np.float64)


def fosfilter(b, a, order, signal, states=None):
    """`order` passes of the first-order section (b, a) over `signal`; returns (complex128 signal,
    complex128 states[order]); gammatone.py:196-237.  numpy in -> numpy out, CUDA tensor in ->
    CUDA tensors out."""
    on_device = isinstance(signal, torch.Tensor) and signal.is_cuda
    dev = signal.device if on_device else _device()
    x = _as_rows(signal, dev).reshape(1, -1)
    st = None
    if states is not None:
        st = torch.as_tensor(np.asarray(states.cpu() if isinstance(states, torch.Tensor) else states,
                                        dtype=np.complex128)).reshape(1, 1, order).to(dev)
    y, st = fos_bank(x, _coef_rows([(b, a)]), order, st)
    y, st = y.reshape(-1), st.reshape(order)
    return (y, st) if on_device else (y.cpu().numpy(), st.cpu().numpy())


def _as_rows(signal, dev):
    """numpy array or tensor -> CUDA float64 tensor, or complex128 when the input has a non-zero imaginary part
    (the reference's lfilter filters complex signals; a complex input with zero imaginary part takes the real path,
    like the impulse the constructor analyses)."""
    if isinstance(signal, torch.Tensor):
        t = signal.detach().to(dev)
        if t.is_complex():
            return t.to(torch.complex128) if bool((t.imag != 0).any()) else t.real.to(torch.float64)
        return t.to(torch.float64)
    a = np.asarray(signal)
    if np.iscomplexobj(a) and np.any(a.imag != 0):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)).to(dev)
    return torch.from_numpy(np.ascontiguousarray(np.real(a), dtype=np.float64)).to(dev)


def _create_impulse(num_samples):
    sig = np.zeros(int(num_samples)) + 0j
    sig[0] = 1.0
    return sig


class GammatoneFilterbank:
    """Same constructor as the reference (gammatone.py:272-311).  `analyze(signal, states=None)` is a
    generator of (band, states) tuples in ascending centre frequency; a 2-D signal (C, N) is treated
    as a batch and yields bands of shape (C, N)."""

    def __init__(self, samplerate=44100, order=4, startband=-12, endband=12, normfreq=1000.0, density=1.0,
                 bandwidth_factor=1.0, desired_delay_sec=0.02):
        self.samplerate = samplerate
        self.order = order
        self.centerfrequencies = frequencies_gammatone_bank(startband, endband, normfreq, density)
        self._coeffs = tuple(design_filtbank_coeffs(samplerate, order, self.centerfrequencies,
                                                    bandwidth_factor=bandwidth_factor))
        self._coef_rows = _coef_rows(self._coeffs)
        self.init_delay(desired_delay_sec)
        self.init_gains()

    def init_delay(self, desired_delay_sec):
        self.desired_delay_sec = desired_delay_sec
        self.desired_delay_samples = int(self.samplerate * desired_delay_sec)
        self.max_indices, self.slopes = self.estimate_max_indices_and_slopes(
            delay_samples=self.desired_delay_samples)
        self.delay_samples = self.desired_delay_samples - self.max_indices
        self.delay_memory = np.zeros((len(self.centerfrequencies), np.max(self.delay_samples)))

    def init_gains(self):
        self.gains = np.ones(len(self.centerfrequencies))

    def analyze(self, signal, states=None):
        on_device = isinstance(signal, torch.Tensor) and signal.is_cuda
        dev = signal.device if on_device else _device()
        x = _as_rows(signal, dev)
        batched = x.dim() == 2
        x = x.reshape(-1, x.shape[-1])
        C, B = x.shape[0], len(self._coeffs)
        st = None
        if states is not None and len(states):
            rows = [np.asarray(s.cpu() if isinstance(s, torch.Tensor) else s, dtype=np.complex128) for s in states]
            st = torch.as_tensor(np.stack(rows).reshape(B, C, self.order)).permute(1, 0, 2).contiguous().to(dev)
        y, st = fos_bank(x, self._coef_rows, self.order, st)
        if not on_device:
            from .sosfiltering import _to_host
            y, st = _to_host(y), st.cpu().numpy()
        for i in range(B):
            band = y[:, i] if batched else y[0, i]
            yield band, (st[:, i] if batched else st[0, i])

    def reanalyze(self, bands, states=None):
        """Every band through ITS OWN cascade once more (gammatone.py:318-321): `bands` is the sequence of complex
        (N,) bands `analyze` yields (or (C, N) batches), `states` the matching sequence of state arrays.  Generator
        of (band, states) like `analyze`."""
        for i, ((b, a), band) in enumerate(zip(self._coeffs, bands)):
            on_device = isinstance(band, torch.Tensor) and band.is_cuda
            dev = band.device if on_device else _device()
            x = _as_rows(band, dev)
            batched = x.dim() == 2
            x = x.reshape(-1, x.shape[-1])
            st = None
            if states is not None and len(states):
                s_i = states[i]
                s_i = np.asarray(s_i.cpu() if isinstance(s_i, torch.Tensor) else s_i, dtype=np.complex128)
                st = torch.as_tensor(s_i.reshape(x.shape[0], 1, self.order)).to(dev)
            y, st = fos_bank(x, self._coef_rows[i:i + 1], self.order, st)
            y, st = (y[:, 0], st[:, 0]) if batched else (y[0, 0], st[0, 0])
            yield (y, st) if on_device else (y.cpu().numpy(), st.cpu().numpy())

    # ------------------------------------------------------------------------------------------
    # synthesis side (gammatone.py:323-340)
    # ------------------------------------------------------------------------------------------
    def _synth_par(self):
        pf = np.abs(self.slopes) * 1j / self.slopes                      # gammatone.py:328
        self.phase_factors = pf
        return np.ascontiguousarray(np.stack([np.asarray(self.gains, dtype=np.float64),
                                              np.asarray(self.delay_samples, dtype=np.float64),
                                              pf.real, pf.imag], axis=1))

    def synthesize_batch(self, bands, delay_memory=None):
        """Fused delay + weighted sum of a batch: bands (C, B, N) complex128 CUDA tensor ->
        (out (C, N) complex128, delay_memory (C, B, Dmax) float64).  One kernel reads every band once
        (pfb_gt_synthesize_c128); `delay_memory` carries the delayed tails between blocks."""
        C, B, N = bands.shape
        dmax = max(int(np.max(self.delay_samples)), 1)
        if delay_memory is None:
            delay_memory = torch.zeros((C, B, dmax), dtype=torch.float64, device=bands.device)
        bands = bands.contiguous()
        out = torch.empty((C, N), dtype=torch.complex128, device=bands.device)
        par = self._synth_par()
        if N > 0:
            with torch.cuda.device(bands.device):
                _lib.check(_lib.lib().pfb_gt_synthesize_c128(bands.data_ptr(), N, out.data_ptr(), N, delay_memory.data_ptr(),
                                                             par.ctypes.data, B, dmax, N, C,
                                                             torch.cuda.current_stream().cuda_stream))
        return out, delay_memory

    def synthesize(self, bands):
        """Sum of the gain-weighted, delayed bands (gammatone.py:323-325) of ONE signal: `bands` is the
        sequence of (N,) complex bands `analyze` yields.  `self.delay_memory` (B, Dmax) is read and updated like
        in the reference.  Returns a real array unless a band has delay 0 (then complex, like the reference)."""
        bands = list(bands)
        on_device = isinstance(bands[0], torch.Tensor) and bands[0].is_cuda
        dev = bands[0].device if on_device else _device()
        if on_device:
            stack = torch.stack([b.to(torch.complex128) for b in bands]).unsqueeze(0)
        else:
            stack = torch.from_numpy(np.stack([np.asarray(b, dtype=np.complex128) for b in bands])).to(dev).unsqueeze(0)
        mem = torch.from_numpy(np.ascontiguousarray(self.delay_memory, dtype=np.float64)).to(dev).unsqueeze(0)
        if mem.shape[2] == 0:
            mem = torch.zeros((1, len(bands), 1), dtype=torch.float64, device=dev)
        out, mem = self.synthesize_batch(stack, mem)
        if self.delay_memory.shape[1] > 0:
            self.delay_memory = mem[0].cpu().numpy()
        y = out[0]
        if not np.any(np.asarray(self.delay_samples) == 0):
            y = y.real
        return y if on_device else y.cpu().numpy()

    def delay(self, bands):
        """Generator of the delayed real parts (gammatone.py:327-340); updates `self.delay_memory`.  Plain
        tensor slicing (this is glue, not a kernel: `synthesize` is the fused path)."""
        self.phase_factors = np.abs(self.slopes) * 1j / self.slopes
        for i, band in enumerate(bands):
            d = int(self.delay_samples[i])
            on_device = isinstance(band, torch.Tensor)
            re = band.real if on_device else np.real(band)
            if d == 0:
                yield re * complex(self.phase_factors[i])
            else:
                mem = self.delay_memory[i, :d]
                if on_device:
                    yield torch.cat([torch.from_numpy(mem.copy()).to(band.device), re[:-d]])
                    self.delay_memory[i, :d] = re[-d:].cpu().numpy()
                else:
                    yield np.concatenate((mem, re[:-d]), axis=0)
                    self.delay_memory[i, :d] = re[-d:]

    def estimate_max_indices_and_slopes(self, delay_samples=None):
        if not delay_samples:
            delay_samples = int(self.samplerate / 10)
        sig = _create_impulse(delay_samples)
        bands = list(zip(*self.analyze(sig)))[0]
        ibandmax = [np.argmax(np.abs(b[:delay_samples])) for b in bands]
        slopes = [b[i + 1] - b[i - 1] for (b, i) in zip(bands, ibandmax)]
        return np.array(ibandmax), np.array(slopes)
