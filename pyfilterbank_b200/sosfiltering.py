"""Second-order-section filtering on the GPU with the reference's Python signatures.

Mirrors pyfilterbank/sosfiltering.py:54-248 (`sosfilter_c`, `sosfilter_double_c`,
`sosfilter_double_mimo_c`): same argument meaning, same sos layout `[b0 b1 b2 a0 a1 a2]` per
section, same `(w1, w2)` state layout, same return shapes and orders.  numpy in -> numpy out
(fresh arrays, the input is never mutated); CUDA `torch.Tensor` in -> CUDA tensors out without
leaving the device.  Extra keyword arguments (`mode`, `exact`, `arith64`) select the kernel;
the defaults match the reference within 1e-9 relative L2 in float64.

PyTorch is used for device memory and streams only; the arithmetic runs in libpfb_sos.so.
"""
import collections

import numpy as np
import torch

from . import _lib
from ._lib import Bank

__all__ = ["sosfilter_c", "sosfilter_double_c", "sosfilter_double_mimo_c", "sosfilter_py", "bank_filter", "bank_levels",
           "bank_filter_weighted", "bilinear_sos", "get_bank", "states_convert"]

_bank_cache = collections.OrderedDict()


def get_bank(sos, warm_tol=None):
    """Bank for sos (B, K, 6) or (C, B, K, 6), cached by content, CUDA device and pre-pass tolerance
    (chunk-transition tables live in the bank, on the device it was created on, so block-wise callers
    should not rebuild it every block).  warm_tol: relative-L2 budget of the truncated pre-pass of the
    time-chunked kernels (None: the library default 1e-11; 0: no truncation) -- part of the key, so changing
    it never alters the accuracy other users of the same coefficients get."""
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    key = (dev, sos.shape, sos.tobytes(), None if warm_tol is None else float(warm_tol))
    bank = _bank_cache.get(key)
    if bank is None:
        bank = Bank(sos)
        if warm_tol is not None:
            bank.set_warm_tol(warm_tol)
        _bank_cache[key] = bank
        if len(_bank_cache) > 32:
            _bank_cache.popitem(last=False)
    else:
        _bank_cache.move_to_end(key)
    return bank


def _device():
    if not torch.cuda.is_available():
        raise _lib.PfbError("no CUDA device: pyfilterbank_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _signal_types(x, arith64):
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be a (C, N) CUDA tensor with contiguous rows")
    if x.dtype == torch.float64:
        return _lib.PFB_F64, torch.float64
    if x.dtype == torch.float32:
        return _lib.PFB_F32, (torch.float64 if arith64 else torch.float32)
    raise ValueError("x must be float32 or float64")


def _check_states(states, sdt, C, B, K, device):
    if states is None:
        return torch.zeros((C, B, K, 2), dtype=sdt, device=device)
    if (not isinstance(states, torch.Tensor) or not states.is_cuda or states.dtype != sdt or not states.is_contiguous()
            or states.numel() != C * B * K * 2):
        raise ValueError("states must be a contiguous CUDA %s tensor with C*B*K*2 elements" % sdt)
    return states


def bank_filter_weighted(bank, x, b, a, states=None, wz=None, *, mode="auto", exact=False, arith64=False, out=None):
    """y = bank(lfilter(b, a, x)): the SPL-weighting prefilter (splweighting.py:61-80) and the bank as ONE call
    (pfb_bank_filter_weighted).  x (C, N) CUDA tensor, float64 or float32 (the prefilter always runs in float64
    arithmetic); wz (C, ntaps - 1) float64 delay values in scipy's zi convention, in/out (None: zeros).
    Returns (y (C, B, N), states, wz)."""
    dtype, sdt = _signal_types(x, arith64)
    C, N = x.shape
    B, K = bank.B, bank.K
    states = _check_states(states, sdt, C, B, K, x.device)
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    nz = max(max(len(b), len(a)) - 1, 1)
    if wz is None:
        wz = torch.zeros((C, nz), dtype=torch.float64, device=x.device)
    elif wz.dtype != torch.float64 or not wz.is_contiguous() or wz.numel() != C * nz or not wz.is_cuda:
        raise ValueError("wz must be a contiguous CUDA float64 tensor with C*(ntaps-1) elements")
    if out is None:
        pitch = (N + 31) // 32 * 32              # 16-byte aligned rows keep the TMA path available
        out = torch.empty((C, B, pitch), dtype=x.dtype, device=x.device)[:, :, :N]
    ldx = x.stride(0) if C > 1 else max(N, 1)
    if N > 0:
        with torch.cuda.device(x.device):
            bank.filter_weighted_ptr(dtype, x.data_ptr(), ldx, out.data_ptr(), max(out.stride(1), N, 1), states.data_ptr(),
                                     b, a, wz.data_ptr(), N, C, _lib.flags_of(mode, exact, arith64),
                                     torch.cuda.current_stream().cuda_stream)
    return out, states, wz


def bank_filter(bank, x, states=None, *, mode="auto", exact=False, arith64=False, out=None, decimate=None):
    """Filter x (C, N) CUDA tensor (float32/float64, rows contiguous) through `bank`.

    Returns (y (C, B, N), states (C, B, K, 2)) CUDA tensors; `states` (if given) is updated in
    place and returned.  This is the device-resident entry the wrappers below are built on.
    decimate=0/1: keep only the filtered samples 2 i + decimate (y is (C, B, (N - decimate + 1) // 2))."""
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be a (C, N) CUDA tensor with contiguous rows")
    C, N = x.shape
    if x.dtype == torch.float64:
        dtype, sdt = _lib.PFB_F64, torch.float64
    elif x.dtype == torch.float32:
        dtype, sdt = _lib.PFB_F32, (torch.float64 if arith64 else torch.float32)
    else:
        raise ValueError("x must be float32 or float64")
    B, K = bank.B, bank.K
    if states is None:
        states = torch.zeros((C, B, K, 2), dtype=sdt, device=x.device)
    elif states.dtype != sdt or not states.is_contiguous() or states.numel() != C * B * K * 2:
        raise ValueError("states must be a contiguous %s tensor with C*B*K*2 elements" % sdt)
    n_out = N if decimate is None else max(0, (N - int(decimate) + 1) // 2)
    if out is None:
        pitch = (n_out + 31) // 32 * 32          # 16-byte aligned rows keep the TMA kernels available for any N
        out = torch.empty((C, B, pitch), dtype=x.dtype, device=x.device)[:, :, :n_out]
    ldx = x.stride(0) if C > 1 else max(N, 1)
    # a decimated block can be empty (one input sample on the odd phase): states still advance
    dst = out if out.numel() else torch.empty(1, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        bank.filter_ptr(dtype, x.data_ptr(), ldx, dst.data_ptr(), max(out.stride(1), n_out, 1),
                        states.data_ptr(), N, C, _lib.flags_of(mode, exact, arith64, decimate),
                        torch.cuda.current_stream().cuda_stream)
    return out, states


def bank_levels(bank, x, block_len=None, states=None, *, arith64=False, weighting=None, wz=None):
    """Band energies of x (C, N) CUDA tensor through `bank` without materialising the band signals:
    returns (energy (C, B, N // block_len) float64 = sum of y^2 per block, states).  block_len=None: one
    block = the whole signal.  N must be a multiple of block_len, block_len of 32 (float64) / 64 (float32).
    weighting=(b, a): transfer-function taps of an SPL-weighting prefilter (5 to 7 taps) run inside the same kernel in
    front of the bank (pfb_bank_levels_weighted); wz (C, ntaps - 1) float64 delay values (scipy's zi, None: zeros);
    the call then returns (energy, states, wz)."""
    if not x.is_cuda or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be a (C, N) CUDA tensor with contiguous rows")
    C, N = x.shape
    if x.dtype == torch.float64:
        dtype, sdt = _lib.PFB_F64, torch.float64
    elif x.dtype == torch.float32:
        dtype, sdt = _lib.PFB_F32, (torch.float64 if arith64 else torch.float32)
    else:
        raise ValueError("x must be float32 or float64")
    block_len = N if block_len is None else int(block_len)
    B, K = bank.B, bank.K
    states = _check_states(states, sdt, C, B, K, x.device)
    nblocks = N // block_len if block_len > 0 else 0
    out = torch.empty((C, B, max(nblocks, 1)), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        if weighting is not None:
            wb, wa = (np.ascontiguousarray(t, dtype=np.float64) for t in weighting)
            nz = max(len(wb), len(wa)) - 1
            if wz is None:
                wz = torch.zeros((C, nz), dtype=torch.float64, device=x.device)
            elif not (wz.is_cuda and wz.dtype == torch.float64 and wz.is_contiguous() and tuple(wz.shape) == (C, nz)):
                raise ValueError("wz must be a contiguous float64 CUDA tensor (C, %d)" % nz)
            bank.levels_weighted_ptr(dtype, x.data_ptr(), x.stride(0) if C > 1 else max(N, 1), out.data_ptr(), out.stride(1),
                                     states.data_ptr(), wb, wa, wz.data_ptr(), N, C, block_len,
                                     _lib.flags_of("auto", False, arith64), torch.cuda.current_stream().cuda_stream)
            return out[:, :, :nblocks], states, wz
        bank.levels_ptr(dtype, x.data_ptr(), x.stride(0) if C > 1 else max(N, 1), out.data_ptr(), out.stride(1),
                        states.data_ptr(), N, C, block_len, _lib.flags_of("auto", False, arith64),
                        torch.cuda.current_stream().cuda_stream)
    return out[:, :, :nblocks], states


def channel_rows(sig, dev, dtype=torch.float64):
    """(N, C) or (N,) signal (numpy array or tensor, host or device) -> (C, N) CUDA tensor view of a buffer whose
    row pitch is a multiple of 32 elements: every row starts 16-byte aligned whatever N is, which keeps the
    TMA kernels available (unaligned rows fall back to the cp.async kernels)."""
    if not isinstance(sig, torch.Tensor):
        sig = torch.from_numpy(np.asarray(sig))
    sig = sig.detach()
    n = sig.shape[0]
    src = sig.reshape(n, -1)
    C = src.shape[1]
    pitch = max(32, (n + 31) // 32 * 32)
    rows = torch.empty((C, pitch), dtype=dtype, device=dev)[:, :n]
    if src.is_cuda:
        rows.copy_(src.t())
    else:
        rows.copy_(src.t().contiguous().to(dtype))      # one contiguous host-to-device copy
    return rows


def _to_host(t):
    """Device tensor -> numpy array.  Large results go through page-locked memory (PCIe DMA at full rate into the
    array the caller receives; torch's host allocator recycles the pinned blocks of results that were dropped)."""
    if t.numel() * t.element_size() < (8 << 20):
        return t.cpu().numpy()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t)
    return host.numpy()


def _sos_rows(sos):
    sos = np.asarray(sos)
    return np.array(sos, dtype=np.float64).reshape(-1, 6)


def _single(signal, sos, states, np_dtype, t_dtype, mode, exact, arith64):
    on_device = isinstance(signal, torch.Tensor) and signal.is_cuda
    dev = signal.device if on_device else _device()
    if isinstance(sos, torch.Tensor):
        sos = sos.detach().cpu().numpy()
    rows = _sos_rows(sos if np_dtype == np.float64 else np.asarray(sos, dtype=np.float32))
    K = rows.shape[0]
    bank = get_bank(rows.reshape(1, K, 6))
    nsamp = len(signal)                                   # sosfiltering.py:83,140
    if on_device:
        x = signal.detach().to(t_dtype).reshape(-1)[:nsamp].reshape(1, nsamp)
    else:
        x = torch.from_numpy(np.array(signal, dtype=np_dtype).flatten()[:nsamp].reshape(1, nsamp)).to(dev)
    sdt = torch.float64 if (np_dtype == np.float64 or arith64) else torch.float32
    if states is None:
        st = torch.zeros(2 * K, dtype=sdt, device=dev)
    elif isinstance(states, torch.Tensor):
        st = states.detach().to(device=dev, dtype=sdt).reshape(-1).clone()
    else:
        st = torch.from_numpy(np.array(states, dtype=np_dtype).flatten()).to(device=dev, dtype=sdt)
    y, st = bank_filter(bank, x.contiguous(), st.reshape(1, 1, K, 2), mode=mode, exact=exact, arith64=arith64)
    y, st = y.reshape(nsamp), st.reshape(2 * K).to(t_dtype)
    if on_device:
        return y, st
    return y.cpu().numpy(), st.cpu().numpy()


def sosfilter_c(signal, sos, states=None, *, mode="auto", exact=False, arith64=False):
    """float32 cascade, the reference's `sosfilter_c` (sosfiltering.py:54-108): signal, sos and
    states are cast to float32.  `exact=True` (sequential, separately rounded operations) is
    bit-identical to sosfilt.c:5-28."""
    return _single(signal, sos, states, np.float32, torch.float32, mode, exact, arith64)


def sosfilter_double_c(signal, sos, states=None, *, mode="auto", exact=False):
    """float64 cascade, the reference's `sosfilter_double_c` (sosfiltering.py:111-163)."""
    return _single(signal, sos, states, np.float64, torch.float64, mode, exact, False)


def states_convert(states, sos, to_df2t):
    """Convert filter memory in place between the two conventions the reference uses: the (w1, w2) Direct-Form-II
    delay values of sosfilt.c (and of every kernel here) and the Direct-Form-II-Transposed `zi` of scipy.signal.lfilter
    that the reference's `sosfilter_py` carries (sosfiltering.py:368-379).  states: float64 CUDA tensor (..., K, 2),
    contiguous; sos (B, K, 6) shared by rows (row q uses bank row q % B) or one row set per cascade (Q, K, 6).
    Runs as a CUDA kernel (pfb_states_convert).  Returns `states`."""
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    K = sos.shape[-2]
    if not (states.is_cuda and states.dtype == torch.float64 and states.is_contiguous() and states.shape[-2:] == (K, 2)):
        raise ValueError("states must be a contiguous float64 CUDA tensor (..., %d, 2)" % K)
    Q = states.numel() // (2 * K)
    rows = sos.reshape(-1, K, 6)
    per_row = int(rows.shape[0] == Q and Q > 1 and sos.ndim > 3)
    if not per_row and Q % rows.shape[0]:
        raise ValueError("%d cascades do not divide into banks of %d rows" % (Q, rows.shape[0]))
    _lib.check(_lib.lib().pfb_states_convert(states.data_ptr(), rows.ctypes.data, Q, rows.shape[0], K, per_row, int(bool(to_df2t)),
                                             torch.cuda.current_stream().cuda_stream))
    return states


def sosfilter_py(x, sos, states=None):
    """The reference's `sosfilter_py` (sosfiltering.py:330-379) on the GPU: sos (K, 6) rows [b0 b1 b2 a0 a1 a2] (a0
    normalises the section, like scipy.signal.lfilter), `states` None or a dict {section index: zi (2,)} of
    Direct-Form-II-Transposed delay values -- what lfilter returns and what the reference's 'py' filter function carries
    between blocks.  Returns (y (N,), states dict).  The cascade runs in the bank kernels (Direct Form II); the state
    adapter kernel (pfb_states_convert) maps the carried values between the two forms on the device, so blocks filtered
    here continue blocks filtered by the reference's sosfilter_py and vice versa (to ~1e-10: the two forms round
    differently)."""
    on_device = isinstance(x, torch.Tensor) and x.is_cuda
    dev = x.device if on_device else _device()
    rows = _sos_rows(sos.detach().cpu().numpy() if isinstance(sos, torch.Tensor) else sos)
    K = rows.shape[0]
    rows = rows / rows[:, 3:4]                                # lfilter divides by a0; sosfilt.c would ignore it
    zi = np.zeros((K, 2))
    if states is not None:
        for i in range(K):
            zi[i] = np.asarray(states[i].detach().cpu().numpy() if isinstance(states[i], torch.Tensor) else states[i],
                               dtype=np.float64).reshape(2)
    st = torch.from_numpy(zi.reshape(1, 1, K, 2)).to(dev)
    states_convert(st, rows.reshape(1, K, 6), to_df2t=False)
    n = len(x)
    if on_device:
        sig = x.detach().to(torch.float64).reshape(1, n)
    else:
        sig = torch.from_numpy(np.array(x, dtype=np.float64).reshape(1, n)).to(dev)
    bank = get_bank(rows.reshape(1, K, 6))
    y, st = bank_filter(bank, sig.contiguous(), st)
    states_convert(st, rows.reshape(1, K, 6), to_df2t=True)
    y = y.reshape(n)
    out = states if states is not None else dict()       # like the reference, a dict passed in is updated and returned
    if on_device:
        out.update({i: st[0, 0, i] for i in range(K)})
        return y, out
    st = st.cpu().numpy()
    out.update({i: st[0, 0, i].copy() for i in range(K)})
    return y.cpu().numpy(), out


def sosfilter_double_mimo_c(signal, sos, states=None, *, mode="auto", exact=False):
    """Multi-channel, multi-band float64 filtering, the reference's `sosfilter_double_mimo_c`
    (sosfiltering.py:166-248).

    signal (N,) or (N, C); sos (6K,), (6K, B) or (6K, B, C); states None, flat (C*B*K*2,) or
    (2K, B, C).  Returns (out (N, B, C) Fortran-ordered, states flat (C*B*K*2,)), like the
    reference.  With a CUDA tensor input the outputs are CUDA tensors: `out` is a (N, B, C) view
    of a (C, B, N) buffer (the same memory order as the reference's Fortran array)."""
    on_device = isinstance(signal, torch.Tensor) and signal.is_cuda
    dev = signal.device if on_device else _device()
    shape = tuple(signal.shape)
    nframes = int(shape[0])
    nchan = int(shape[1]) if len(shape) > 1 else 1
    if isinstance(sos, torch.Tensor):
        sos = sos.detach().cpu().numpy()
    sos = np.asarray(sos, dtype=np.float64)
    ksos = sos.shape[0] // 6
    if sos.ndim == 1:
        kbands = 1
        bank_sos = sos.reshape(1, ksos, 6)
    elif sos.ndim == 2:
        kbands = sos.shape[1]
        bank_sos = np.ascontiguousarray(sos.T).reshape(kbands, ksos, 6)          # column b = band b
    else:
        kbands = sos.shape[1]
        if sos.shape[2] != nchan:
            raise ValueError("sos has %d channels, signal has %d" % (sos.shape[2], nchan))
        bank_sos = np.ascontiguousarray(sos.transpose(2, 1, 0)).reshape(nchan, kbands, ksos, 6)
    bank = get_bank(bank_sos)
    x = channel_rows(signal, dev)
    nst = nchan * kbands * ksos * 2
    if states is None:
        st = torch.zeros(nst, dtype=torch.float64, device=dev)
    elif isinstance(states, torch.Tensor):
        st = states.detach().to(device=dev, dtype=torch.float64)
        st = (st.permute(*reversed(range(st.dim()))).contiguous() if st.dim() > 1 else st.clone()).reshape(-1)
    else:
        st = torch.from_numpy(np.array(states, dtype=np.float64).flatten('F')).to(dev)   # sosfiltering.py:214
    if st.numel() != nst:
        raise ValueError("states must hold C*B*K*2 = %d values" % nst)
    y, st = bank_filter(bank, x, st.reshape(nchan, kbands, ksos, 2), mode=mode, exact=exact)
    st = st.reshape(-1)
    if on_device:
        return y.permute(2, 1, 0), st
    return _to_host(y).transpose(2, 1, 0), st.cpu().numpy()


def bilinear_sos(d, c):
    """Bilinear transform of analog 1-pole-pair section weights to digital biquads
    (the reference's `bilinear_sos`, sosfiltering.py:382-432): for each section with numerator
    weights (d0, d1) and denominator weights (c0, c1) returns rows b, a (L/2 x 3) with a[:, 0] == 1."""
    d = np.asarray(d)
    c = np.asarray(c)
    if d.ndim != 2 or c.ndim != 2 or d.shape[1] != 2 or c.shape != d.shape:
        raise Exception('Inputs d and c must both be L/2 x 2 arrays.')

    def quad(w):
        s, t = w[:, 0] + w[:, 1], w[:, 1] - w[:, 0]
        return np.stack([np.abs(s) ** 2, 2 * np.real(s * np.conj(t)), np.abs(t) ** 2], axis=1).astype(np.double)

    a, b = quad(c), quad(d)
    if np.min(a[:, 0]) == 0:
        raise Exception('"c" should not have a row of zeros.')
    norm = a[:, :1].copy()
    return b / norm, a / norm
