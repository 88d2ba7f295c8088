"""Fractional-octave filter bank on the GPU with the reference's interface.

Mirrors pyfilterbank/octbank.py:75-173 (frequency grid and band-pass design, host/numpy) and
:210-476 (`FractionalOctaveFilterbank` with `filter` and `filter_mimo_c`).  Filtering runs in
libpfb_sos.so; `filterfun` names the arithmetic:

    'cuda'  (default; 'cffi' is accepted as an alias for drop-in use) -- FMA / time-chunked
            kernels, within 1e-9 relative L2 of the reference's sosfilt.c in float64
    'exact' -- sequential, separately rounded: bit-identical to sosfilt.c

There is no CPU path ('py' / 'cprototype' of the reference are not provided).

Decimating band path (`filter_decimated`, opt-in; the reference has no resampling anywhere): band b
runs at fs / 2^m_b.  The semantics are composed from reference pieces so that they can be checked
(SURVEY.md section 8 a6'):

    stage 0 signal = the input; stage m signal = the stage m-1 signal filtered by
    `butter_sos('lowpass', 8, 0.2)` (butterworth.py:16) with `sosfilter_double_c` arithmetic and
    then decimated `[::2]`;
    band b lives at the deepest stage m with  upper band edge * (1 + edge_correction) <= 0.2 * fs / 2^m
    (m <= max_stage); its band-pass is `design_sosmat_band_passes` at the stage's rate (octbank.py:132).
"""
import numpy as np
import torch

from .butterworth import butter_sos
from . import sosfiltering as _sf

standardized_nominal_frequencies = np.array([
    0.1, 0.125, 0.16, 0.2, 0.25, 0.315, 0.4, 0.5, 0.63, 0.8, 1, 1.25, 1.6, 2, 2.5, 3.15, 4, 5, 6.3, 8,
    10, 12.5, 16, 20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800,
    1000, 1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500, 16000, 20000])


def centerfreq_to_bandnum(center_freq, norm_freq, nth_oct):
    return nth_oct * np.log2(center_freq / norm_freq)


def find_nominal_freq(center_frequencies, nominal_frequencies):
    for f in center_frequencies:
        yield nominal_frequencies[np.argmin(np.abs(standardized_nominal_frequencies - f))]


def frequencies_fractional_octaves(start_band, end_band, norm_freq, nth_oct):
    """Centre frequencies norm_freq * 2^(k / nth_oct), k = start_band..end_band, and the
    len+1 band edges (geometric means of neighbours).  octbank.py:75-104."""
    k = np.arange(start_band - 1, end_band + 2)
    grid = norm_freq * 2.0 ** (k / float(nth_oct))
    return grid[1:-1], np.sqrt(grid[:-1] * grid[1:])


def to_normalized_frequencies(frequencies, sample_rate, clip=True):
    """Frequencies / sample_rate; edges at or above Nyquist are clipped to 0.499 and everything
    after the first clipped edge is dropped (clip=True) or just dropped.  octbank.py:107-129."""
    frequencies = np.asarray(frequencies, dtype=float)
    over = frequencies >= 0.5 * sample_rate
    if clip and over.any():
        first = int(np.argmax(over))
        out = frequencies[:first + 1].copy()
        out[first] = 0.499 * sample_rate
        return out / sample_rate
    return frequencies[~over] / sample_rate


def design_sosmat_band_passes(order, band_edges, sample_rate, edge_correction_percent=0.0, device=False):
    """(order*6, num_bands) matrix; column b holds band b's `order` biquads
    [b0 b1 b2 a0 a1 a2] flattened row-major.  octbank.py:132-173.
    device=True designs all bands in one CUDA kernel (butterworth.butter_sos_cuda) instead of the numpy loop."""
    num_bands = len(band_edges) - 1
    sosmat = np.zeros((order * 6, num_bands))
    edges = to_normalized_frequencies(band_edges, sample_rate)
    lo_scale = 1 - edge_correction_percent * 1e-2
    hi_scale = 1 + edge_correction_percent * 1e-2
    if device:
        from .butterworth import butter_sos_cuda
        nb = len(edges) - 1
        if nb > 0:
            sosmat[:, :nb] = butter_sos_cuda('bandpass', order, lo_scale * edges[:-1], hi_scale * edges[1:]).reshape(nb, -1).T
        return sosmat
    for i in range(len(edges) - 1):
        sosmat[:, i] = butter_sos('bandpass', order, lo_scale * edges[i], hi_scale * edges[i + 1]).flatten()
    return sosmat


class FractionalOctaveFilterbank:
    """Butterworth fractional-octave band-pass bank (octbank.py:210-476), same constructor
    arguments, properties and `filter` / `filter_mimo_c` signatures and return conventions."""

    def __init__(self, sample_rate=44100, order=4, nth_oct=3.0, norm_freq=1000.0, start_band=-19,
                 end_band=13, edge_correction_percent=0.01, filterfun='cuda', design='host'):
        if design not in ('host', 'device'):
            raise ValueError("design must be 'host' (numpy) or 'device' (CUDA kernel)")
        self._design = design
        self._sample_rate = sample_rate
        self._order = order
        self._nth_oct = nth_oct
        self._norm_freq = norm_freq
        self._start_band = start_band
        self._end_band = end_band
        self._edge_correction_percent = edge_correction_percent
        self._initialize_filter_bank()
        self.set_filterfun(filterfun)

    def _prop(name):   # noqa: N805  (every setter re-designs the bank, octbank.py:286-347)
        def get(self):
            return getattr(self, '_' + name)

        def set_(self, value):
            setattr(self, '_' + name, value)
            self._initialize_filter_bank()
        return property(get, set_)

    sample_rate = _prop('sample_rate')
    order = _prop('order')
    nth_oct = _prop('nth_oct')
    norm_freq = _prop('norm_freq')
    start_band = _prop('start_band')
    end_band = _prop('end_band')
    edge_correction_percent = _prop('edge_correction_percent')
    del _prop

    center_frequencies = property(lambda self: self._center_frequencies)
    band_edges = property(lambda self: self._band_edges)
    sosmat = property(lambda self: self._sosmat)
    num_bands = property(lambda self: len(self._center_frequencies))
    band_widths = property(lambda self: np.diff(self._band_edges))

    @property
    def effective_filter_lengths(self):
        return [int(l) for l in self.sample_rate * 3 // self.band_widths]

    def _initialize_filter_bank(self):
        self._center_frequencies, self._band_edges = frequencies_fractional_octaves(
            self.start_band, self.end_band, self.norm_freq, self.nth_oct)
        self._sosmat = design_sosmat_band_passes(self.order, self.band_edges, self.sample_rate,
                                                 self.edge_correction_percent, device=self._design == 'device')
        self._dec_plan = None

    # ------------------------------------------------------------------------------------------
    # decimating band path (new, opt-in)
    # ------------------------------------------------------------------------------------------
    AA_ORDER = 8          # low-pass poles of the anti-alias filter of every halving stage (4 biquads)
    AA_CUTOFF = 0.2       # its cutoff, normalised to the stage's INPUT rate
    EDGE_LIMIT = 0.2      # a band may run at rate f if its (corrected) upper edge is <= EDGE_LIMIT * f
    MAX_STAGE = 10

    def decimation_plan(self):
        """List of stages [{'m', 'rate', 'bands' (ascending band indices), 'sosmat' (6K, B_m)}],
        stage 0 first; every band appears in exactly one stage.  Also sets `decimation_factors`."""
        if self._dec_plan is not None:
            return self._dec_plan
        B = self.num_bands
        fs = float(self.sample_rate)
        hi = self.band_edges[1:] * (1 + self.edge_correction_percent * 1e-2)
        m_b = np.zeros(B, dtype=int)
        for b in range(B):
            m = 0
            while m < self.MAX_STAGE and hi[b] <= self.EDGE_LIMIT * fs / 2 ** (m + 1):
                m += 1
            m_b[b] = m
        stages = []
        for m in range(int(m_b.max()) + 1):
            idx = np.nonzero(m_b == m)[0]
            if len(idx) == 0:
                stages.append({'m': m, 'rate': fs / 2 ** m, 'bands': [], 'sosmat': None})
                continue
            lo, hi_i = int(idx[0]), int(idx[-1])
            assert len(idx) == hi_i - lo + 1           # contiguous because edges ascend
            sosmat = design_sosmat_band_passes(self.order, self.band_edges[lo:hi_i + 2], fs / 2 ** m,
                                               self.edge_correction_percent)
            stages.append({'m': m, 'rate': fs / 2 ** m, 'bands': list(range(lo, hi_i + 1)),
                           'sosmat': sosmat[:, :len(idx)]})
        self._dec_plan = stages
        self._dec_factors = 2 ** m_b
        return stages

    @property
    def decimation_factors(self):
        self.decimation_plan()
        return self._dec_factors

    @property
    def aa_sos(self):
        """(4, 6) anti-alias low-pass applied before every `[::2]`."""
        return butter_sos('lowpass', self.AA_ORDER, self.AA_CUTOFF)

    def _dec_bank(self):
        """The C-side decimating bank (pfb_decbank_*), built once per design and CUDA device."""
        dev = torch.cuda.current_device()
        cache = getattr(self, '_dec_banks', None)
        if cache is None or cache[0] is not self._dec_plan:
            cache = (self._dec_plan, {})
            self._dec_banks = cache
        if dev not in cache[1]:
            from ._lib import DecBank
            K = self.order
            stage_sos = [None if not st['bands'] else
                         np.ascontiguousarray(st['sosmat'].T).reshape(len(st['bands']), K, 6) for st in self._dec_plan]
            cache[1][dev] = DecBank(stage_sos, K, self.aa_sos)
        return cache[1][dev]

    def filter_decimated(self, x, states=None):
        """Per-band decimated analysis of x (N,) or (N, C) (numpy or CUDA tensor).

        Returns (bands, states): `bands[b]` is band b at rate fs / decimation_factors[b], shape
        (ceil-length N_b, C); `states` carries the anti-alias states, the decimator phases and the
        band-pass states, so consecutive blocks of any length concatenate to the one-shot result.
        The whole analysis is ONE call into the library (pfb_decbank_filter): per stage a single pass
        over the stage signal filters the stage's bands and, as one more band warp with a decimating
        store, the anti-alias cascade that produces the next stage's signal -- no band signal is ever
        written at a rate above its own.  `states['aa'][m]` is (C, Kp, 2): the anti-alias cascade is
        padded with pass-through sections to the bands' section count (their delay values are
        irrelevant to the output)."""
        plan = self.decimation_plan()
        on_device = isinstance(x, torch.Tensor) and x.is_cuda
        K = self.order
        dev = x.device if on_device else _sf._device()
        sig = _sf.channel_rows(x, dev)
        C, n = sig.shape
        M = len(plan) - 1
        with torch.cuda.device(dev):
            dec = self._dec_bank()
            Kp = dec.Kp
            if states is None:
                states = {'aa': [None] * M, 'phase': [0] * M, 'bands': [None] * len(plan)}
            aa = torch.zeros((max(M, 1), C, Kp, 2), dtype=torch.float64, device=dev)
            for m in range(M):
                if states['aa'][m] is not None:
                    aa[m] = torch.as_tensor(states['aa'][m], dtype=torch.float64, device=dev).reshape(C, Kp, 2)
            phases = [int(p) for p in states['phase']]
            new = {'aa': [], 'phase': [], 'bands': []}
            ys, ldys, y_ptrs, st_ptrs = [], [], [], []
            nm = n
            for m, st in enumerate(plan):
                Bm = len(st['bands'])
                if Bm:
                    pitch = (nm + 31) // 32 * 32          # 16-byte aligned rows keep the TMA path available
                    y = torch.empty((C, Bm, max(pitch, 32)), dtype=torch.float64, device=dev)
                    bst = states['bands'][m]
                    bst = (torch.zeros((C, Bm, K, 2), dtype=torch.float64, device=dev) if bst is None else
                           torch.as_tensor(bst, dtype=torch.float64, device=dev).reshape(C, Bm, K, 2).clone())
                    ys.append((y, nm))
                    ldys.append(y.stride(1))
                    y_ptrs.append(y.data_ptr())
                    st_ptrs.append(bst.data_ptr())
                    new['bands'].append(bst)
                else:
                    ys.append(None)
                    ldys.append(0)
                    y_ptrs.append(None)
                    st_ptrs.append(None)
                    new['bands'].append(None)
                if m < M:
                    nm = max(0, (nm - phases[m] + 1) // 2)
            flags = _sf._lib.flags_of("seq" if self._exact else "auto", self._exact)
            dec.filter_ptr(sig.data_ptr(), sig.stride(0) if C > 1 else max(n, 1), y_ptrs, ldys, st_ptrs, aa.data_ptr(),
                           phases, n, C, flags, torch.cuda.current_stream().cuda_stream)
        new['aa'] = [aa[m] for m in range(M)]
        new['phase'] = phases
        out = [None] * self.num_bands
        for st, yy in zip(plan, ys):
            if yy is None:
                continue
            y, nm = yy
            for j, b in enumerate(st['bands']):
                yb = y[:, j, :nm].t()                                     # (N_m, C) view
                out[b] = yb if on_device else yb.cpu().numpy()
        return out, new

    def set_filterfun(self, filterfun_name):
        name = filterfun_name.lower()
        self._py_states = False
        if name in ('cuda', 'cffi'):
            self._exact = False
        elif name == 'exact':
            self._exact = True
        elif name == 'py':
            # the STATE CONVENTION of the reference's 'py' filter function (sosfilter_py: dict {section: lfilter zi}
            # per band, octbank.py:402-404), computed by the CUDA kernels; not a CPU path
            self._exact = False
            self._py_states = True
        else:
            raise ValueError("filterfun must be 'cuda' (alias 'cffi'), 'exact' or 'py' (state convention of the "
                             "reference's lfilter path); this package has no CPU path")
        self.filterfun_name = name
        mode = "seq" if self._exact else "auto"
        if self._py_states:
            self.sosfilterfun = _sf.sosfilter_py
        else:
            self.sosfilterfun = lambda x, sos, states=None: _sf.sosfilter_double_c(
                x, sos, states, mode=mode, exact=self._exact)

    def filter_mimo_c(self, x, states=None, weighting=None, weighting_states=None):
        """(N,) or (N, C) signal -> ((N, B, C) Fortran-ordered bands, flat states); octbank.py:412-434.

        weighting='A' | 'B' | 'C' (opt-in, not in the reference's signature) runs the SPL weighting
        prefilter of splweighting.weight_signal (splweighting.py:61-80) on every channel on the device
        before the bank, so the weighted signal never visits the host; the call then returns
        (bands, states, weighting_states) and `weighting_states` ((C, ntaps-1) delay values in scipy's
        zi convention) can be passed back for block-wise processing."""
        if weighting is None:
            return _sf.sosfilter_double_mimo_c(x, self.sosmat, states, mode="seq" if self._exact else "auto",
                                               exact=self._exact)
        from . import splweighting as _spl
        b, a = _spl._weighting_coeff_design_funsd[weighting](self.sample_rate)
        on_device = isinstance(x, torch.Tensor) and x.is_cuda
        rows = _sf.channel_rows(x, x.device if on_device else _sf._device())   # (C, N)
        if weighting_states is not None and not isinstance(weighting_states, torch.Tensor):
            weighting_states = torch.from_numpy(np.array(weighting_states, dtype=np.float64)).to(rows.device)
        wst = None if weighting_states is None else weighting_states.clone()
        C, B, K = rows.shape[0], self.num_bands, self.order
        bank = _sf.get_bank(np.ascontiguousarray(self.sosmat.T).reshape(B, K, 6))
        if states is None:
            st = torch.zeros((C, B, K, 2), dtype=torch.float64, device=rows.device)
        else:
            st = states if isinstance(states, torch.Tensor) else torch.from_numpy(np.array(states, dtype=np.float64).flatten('F'))
            st = st.to(device=rows.device, dtype=torch.float64).reshape(C, B, K, 2).clone()
        mode = "seq" if self._exact else "auto"
        # ONE library call (pfb_bank_filter_weighted): with enough channels the prefilter runs as an extra warp of
        # the bank kernel on the shared-memory input tiles (the weighted signal never touches HBM); with few
        # channels the weighting kernel and the time-chunked bank run back to back inside the call
        yb, st, wst = _sf.bank_filter_weighted(bank, rows, b, a, st, wst, mode=mode, exact=self._exact)
        y, st = yb.permute(2, 1, 0), st.reshape(-1)
        if on_device:
            return y, st, wst
        return _sf._to_host(yb).transpose(2, 1, 0), st.cpu().numpy(), wst.cpu().numpy()

    def filter_levels(self, x, block_len=None, states=None, db=False, weighting=None, weighting_states=None):
        """Band energies without the band signals (opt-in; the reference's users compute
        `10*log10(sum(y**2))` from filter_mimo_c's output, octbank.py:262-263): x (N,) or (N, C) ->
        (energy (N // block_len, B, C), flat states like filter_mimo_c).  The squares are summed inside
        the filtering kernel, so nothing of size N*B*C is written.  N must be a multiple of block_len
        (None: the whole signal) and block_len of 32 samples.  db=True returns 10*log10(energy).
        weighting='A' | 'B' | 'C': the SPL weighting prefilter of splweighting.weight_signal (splweighting.py:61-80) runs
        inside the same kernel in front of the bank -- A-weighted band levels with nothing but the input read from
        memory; the call then returns (energy, states, weighting_states) like filter_mimo_c(weighting=...)."""
        on_device = isinstance(x, torch.Tensor) and x.is_cuda
        if on_device:
            sig = x.detach().to(torch.float64)
        else:
            sig = torch.from_numpy(np.array(x, dtype=np.float64)).to(_sf._device())
        n = sig.shape[0]
        rows = sig.reshape(n, -1).t().contiguous()
        C, B, K = rows.shape[0], self.num_bands, self.order
        bank = _sf.get_bank(np.ascontiguousarray(self.sosmat.T).reshape(B, K, 6))
        st = None
        if states is not None:
            st = states if isinstance(states, torch.Tensor) else torch.from_numpy(np.array(states, dtype=np.float64))
            st = st.to(device=rows.device, dtype=torch.float64).reshape(C, B, K, 2).clone()
        wst = None
        if weighting is not None:
            from . import splweighting as _spl
            wb, wa = _spl._weighting_coeff_design_funsd[weighting](self.sample_rate)
            if weighting_states is not None:
                wst = weighting_states if isinstance(weighting_states, torch.Tensor) else torch.from_numpy(
                    np.array(weighting_states, dtype=np.float64))
                wst = wst.to(device=rows.device, dtype=torch.float64).contiguous().clone()
            e, st, wst = _sf.bank_levels(bank, rows, block_len, st, weighting=(wb, wa), wz=wst)
        else:
            e, st = _sf.bank_levels(bank, rows, block_len, st)
        e = e.permute(2, 1, 0)
        if db:
            e = 10.0 * torch.log10(e)
        st = st.reshape(-1)
        if weighting is not None:
            return (e, st, wst) if on_device else (e.cpu().numpy(), st.cpu().numpy(), wst.cpu().numpy())
        if on_device:
            return e, st
        return e.cpu().numpy(), st.cpu().numpy()

    def filter(self, x, ffilt=False, states=None):
        """Single-channel filtering, octbank.py:436-476: returns (y (N, B) float64, dict
        {centre frequency: states (2K,)}).  The B per-band cascades of the reference's loop run as
        one bank launch; with ffilt the time-reversed signal is filtered and that result, reversed
        again, is filtered once more with the states chained through both passes, like the reference."""
        on_device = isinstance(x, torch.Tensor) and x.is_cuda
        n = len(x)
        B, K = self.num_bands, self.order
        cfs = self.center_frequencies
        st = None
        py = getattr(self, '_py_states', False)
        sos_rows = np.ascontiguousarray(self.sosmat.T).reshape(B, K, 6)
        if isinstance(states, dict):
            rows = []
            for f in cfs:
                s = states[f]
                if s is None:
                    rows.append(np.zeros(2 * K))
                elif py:                                  # {section index: zi (2,)}, sosfiltering.py:368-379
                    rows.append(np.concatenate([np.asarray(
                        s[i].detach().cpu().numpy() if isinstance(s[i], torch.Tensor) else s[i], dtype=np.float64).reshape(2)
                        for i in range(K)]))
                elif isinstance(s, torch.Tensor):
                    rows.append(s.detach().cpu().numpy().reshape(-1))
                else:
                    rows.append(np.asarray(s, dtype=np.float64).reshape(-1))
            st = np.stack(rows).reshape(1, B, K, 2)
        bank = _sf.get_bank(np.ascontiguousarray(self.sosmat.T).reshape(B, K, 6))
        if on_device:
            dev = x.device
            sig = x.detach().to(torch.float64).reshape(-1)[:n]
        else:
            dev = _sf._device()
            sig = torch.from_numpy(np.array(x, dtype=np.float64).flatten()[:n]).to(dev)
        st_t = None if st is None else torch.from_numpy(st).to(dev)
        if py and st_t is not None:
            _sf.states_convert(st_t, sos_rows, to_df2t=False)          # lfilter zi -> (w1, w2), on the device
        mode = "seq" if self._exact else "auto"
        if not ffilt:
            y, st_t = _sf.bank_filter(bank, sig.reshape(1, n), st_t, mode=mode, exact=self._exact)
            y = y[0]                                                        # (B, N)
        else:
            y1, st_t = _sf.bank_filter(bank, sig.flip(0).reshape(1, n), st_t, mode=mode, exact=self._exact)
            # second pass: every band filters ITS OWN reversed first-pass output -> B independent rows
            rows = y1[0].flip(1).contiguous()                               # (B, N)
            bank2 = _sf.get_bank(np.ascontiguousarray(self.sosmat.T).reshape(B, 1, K, 6))
            y2, st_t = _sf.bank_filter(bank2, rows, st_t.reshape(B, 1, K, 2), mode=mode, exact=self._exact)
            y = y2[:, 0]
        if py:
            st_t = _sf.states_convert(st_t.reshape(1, B, K, 2).contiguous(), sos_rows, to_df2t=True)
            st_np = st_t.reshape(B, K, 2).cpu().numpy()
            st_py = {f: {k: st_np[i, k].copy() for k in range(K)} for i, f in enumerate(cfs)}
            return (y.t(), st_py) if on_device else (_sf._to_host(y.t().contiguous()), st_py)
        st_rows = st_t.reshape(B, 2 * K)
        if on_device:
            return y.t(), {f: st_rows[i] for i, f in enumerate(cfs)}
        y_data = _sf._to_host(y.t().contiguous())                           # (N, B) C-ordered like the reference
        st_np = st_rows.cpu().numpy()
        return y_data, {f: st_np[i].copy() for i, f in enumerate(cfs)}
