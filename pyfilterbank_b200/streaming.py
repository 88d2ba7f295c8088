"""Block-wise filtering of host audio with carried states: a persistent streaming session.

The reference's filter functions take and return `states` so that a long recording can be processed
block by block (pyfilterbank/blockprocessing.py:3-72 cuts the blocks, sosfiltering.py:86-90,143-147,
210-214 carry the states).  `BlockSession` is the device-side counterpart (pfb_stream_* in
include/pfb_sos.h): the states stay on the GPU, streams, events and staging buffers are created once,
`push` enqueues the input copy, the kernels and the output copy of one block and returns, and blocks
of one length allocate nothing after the first.
"""
import numpy as np
import torch

from . import _lib
from .sosfiltering import get_bank, _device


class BlockSession:
    """Stream blocks of a (N, C) signal through a shared bank.

    sos: (B, K, 6) coefficient rows [b0 b1 b2 a0 a1 a2], or a FractionalOctaveFilterbank (its `sosmat`).
    `push(block)` takes a (n, C) (or (n,) for C == 1) array with n <= block_len and returns a ticket;
    `result(ticket)` waits for that block and returns its band signals as a (n, B, C) Fortran-ordered
    view (the reference's filter_mimo_c layout) of a page-locked ring slot that stays valid until `depth`
    further blocks have been pushed.  exact=True runs the bit-identical sequential kernel."""

    def __init__(self, sos, channels, block_len, *, dtype=np.float64, exact=False, arith64=False, depth=4):
        if hasattr(sos, "sosmat"):
            m = sos.sosmat
            sos = np.ascontiguousarray(m.T).reshape(m.shape[1], m.shape[0] // 6, 6)
        self.dev = _device()
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("dtype must be float32 or float64")
        code = _lib.PFB_F64 if self.np_dtype == np.float64 else _lib.PFB_F32
        self.C, self.L, self.depth = int(channels), int(block_len), max(2, int(depth))
        with torch.cuda.device(self.dev):
            self.bank = get_bank(sos)
            self.B, self.K = self.bank.B, self.bank.K
            flags = _lib.flags_of("seq" if exact else "auto", exact, arith64)
            self._s = _lib.Stream(self.bank, code, self.C, self.L, flags)
        tdt = torch.float64 if code == _lib.PFB_F64 else torch.float32
        self._state_dtype = np.float64 if (code == _lib.PFB_F64 or arith64) else np.float32
        self._x = [torch.empty((self.C, self.L), dtype=tdt, pin_memory=True) for _ in range(self.depth)]
        self._y = [torch.empty((self.C, self.B, self.L), dtype=tdt, pin_memory=True) for _ in range(self.depth)]
        self._len = [0] * self.depth
        self._pushed = 0

    def push(self, block):
        a = np.asarray(block)
        n = a.shape[0]
        if n > self.L or a.reshape(n, -1).shape[1] != self.C:
            raise ValueError("block must be (n <= %d, %d)" % (self.L, self.C))
        slot = self._pushed % self.depth
        if self._pushed >= self.depth:
            self._s.wait(self._pushed - self.depth)     # the slot's previous block may still be in flight
        self._x[slot].numpy()[:, :n] = a.reshape(n, self.C).T
        with torch.cuda.device(self.dev):
            self._s.push(self._x[slot].data_ptr(), self.L, self._y[slot].data_ptr(), self.L, n)
        self._len[slot] = n
        self._pushed += 1
        return self._pushed - 1

    def result(self, ticket):
        if not (self._pushed - self.depth <= ticket < self._pushed):
            raise ValueError("ticket %d is no longer (or not yet) in the ring" % ticket)
        with torch.cuda.device(self.dev):
            self._s.wait(ticket)
        slot = ticket % self.depth
        return self._y[slot].numpy()[:, :, :self._len[slot]].transpose(2, 1, 0)

    def filter(self, block):
        """push + result: one block in, its (n, B, C) band signals out."""
        return self.result(self.push(block))

    @property
    def states(self):
        """Carried states, flat (C*B*K*2,) like filter_mimo_c returns them."""
        out = np.empty(self.C * self.B * self.K * 2, dtype=self._state_dtype)
        with torch.cuda.device(self.dev):
            return self._s.get_states(out)

    @states.setter
    def states(self, value):
        v = np.ascontiguousarray(np.asarray(value, dtype=self._state_dtype).reshape(-1))
        if v.size != self.C * self.B * self.K * 2:
            raise ValueError("states must hold C*B*K*2 values")
        with torch.cuda.device(self.dev):
            self._s.set_states(v)

    def stats(self):
        return self._s.stats()

    def close(self):
        self._s.close()
