"""Frame pipeline: streams of CPI frames through  LS_Filter -> fast_xambg  on one B200.

The reference processes one dask chunk (= one CPI frame) per task on a thread pool
(``main.py:169-194``).  On the GPU the same independence is used differently: a
``FramePipeline`` owns ``nslots`` CUDA streams, each with its own device buffers and
libprcore workspace, and issues frame ``i`` on slot ``i % nslots``.  Host->device copies,
the kernels of neighbouring frames and device->host copies of the maps overlap; the
single-CTA Toeplitz solve of one frame hides behind the lag-correlation kernels of the
others.  torch is used for device memory, pinned host memory, streams and events only.

    pipe = FramePipeline(n=2**20, range_bins=300, freq_bins=256)      # LS filterLen = range_bins
    maps = pipe.run_host(ref_frames, srv_frames)      # (nframes, F, R+1, 1) complex64, host
    pipe.run_device(ref_d, srv_d, maps_d)             # everything resident in HBM
"""
from __future__ import annotations

import numpy as np
import scipy.signal as signal

from . import _lib


def _torch():
    import torch
    return torch


def pinned_empty(shape, dtype=np.complex64):
    """Host ndarray backed by page-locked memory (so copies to the GPU are truly async)."""
    torch = _torch()
    tdt = {np.dtype(np.complex64): torch.complex64, np.dtype(np.float32): torch.float32,
           np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
    t = torch.empty(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=tdt, pin_memory=True)
    a = t.numpy()
    _PINNED[a.ctypes.data] = t          # keep the owner alive as long as the module
    return a


_PINNED = {}


class FramePipeline:
    def __init__(self, n, range_bins, freq_bins, filter_len=None, reg=1.0, peek=10,
                 window=("kaiser", 5.0), device=None, nslots=4):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.PrcoreError(_lib.PRC_E_CUDA, "FramePipeline needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.n = int(n)
        self.R = int(range_bins)
        self.F = int(freq_bins)
        self.filter_len = self.R if filter_len is None else int(filter_len)
        self.peek = int(peek)
        self.reg = float(reg)
        self.device = _lib.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        self.nslots = int(nslots)
        if isinstance(window, (tuple, str)):
            window = signal.get_window(window, self.n)
        self.window = None
        if window is not None:
            w = np.ascontiguousarray(window, dtype=np.float64)
            if w.shape != (self.n,):
                raise ValueError(f"window must have shape ({self.n},)")
            self.window = torch.from_numpy(w.astype(np.float32)).to(self.tdev)
        with torch.cuda.device(self.tdev):
            self.streams = [torch.cuda.Stream(device=self.tdev) for _ in range(self.nslots)]
            self.ref_d = [torch.empty(self.n, dtype=torch.complex64, device=self.tdev) for _ in range(self.nslots)]
            self.srv_d = [torch.empty(self.n, dtype=torch.complex64, device=self.tdev) for _ in range(self.nslots)]
            self.map_d = [torch.empty((self.F, self.R + 1), dtype=torch.complex64, device=self.tdev)
                          for _ in range(self.nslots)]
            self.done = [torch.cuda.Event() for _ in range(self.nslots)]
        self.map_bytes = self.F * (self.R + 1) * 8
        self.frame_bytes = 2 * self.n * 8

    # ------------------------------------------------------------------ one frame, device pointers
    def _enqueue(self, ref_ptr, srv_ptr, map_ptr, stream, taps_ptr=None, cleaned_ptr=None):
        flags = _lib.FLAG_ASYNC | (_lib.FLAG_WINDOW_F32 if self.window is not None else 0)
        st = self.lib.prc_frame_c64(ref_ptr, srv_ptr, self.n, self.filter_len, self.peek, self.reg,
                                    self.R, self.F,
                                    None if self.window is None else self.window.data_ptr(),
                                    map_ptr, taps_ptr, cleaned_ptr, _lib.MEM_DEVICE, self.device,
                                    stream.cuda_stream, flags)
        _lib.check(st)

    def run_device(self, ref_d, srv_d, maps_d):
        """ref_d, srv_d: (nframes, n) complex64 CUDA tensors; maps_d: (nframes, F, R+1) complex64.
        Work is forked from, and joined back into, the current torch stream."""
        torch = _torch()
        nf = ref_d.shape[0]
        cur = torch.cuda.current_stream(self.tdev)
        fork = torch.cuda.Event()
        fork.record(cur)
        used = min(self.nslots, nf)
        for s in self.streams[:used]:
            s.wait_event(fork)
        for i in range(nf):
            s = self.streams[i % self.nslots]
            self._enqueue(ref_d[i].data_ptr(), srv_d[i].data_ptr(), maps_d[i].data_ptr(), s)
        for k in range(used):
            self.done[k].record(self.streams[k])
            cur.wait_event(self.done[k])
        return maps_d

    def run_host(self, ref_frames, srv_frames, out=None):
        """ref_frames, srv_frames: (nframes, n) complex64 host arrays (pinned for full overlap).
        Returns (nframes, F, R+1, 1) complex64 host array; synchronises before returning."""
        torch = _torch()
        ref_frames = np.asarray(ref_frames)
        srv_frames = np.asarray(srv_frames)
        if ref_frames.shape != srv_frames.shape:
            raise ValueError('Input vectors must have the same length')
        if ref_frames.ndim != 2 or ref_frames.shape[1] != self.n:
            raise ValueError(f"expected frames of shape (nframes, {self.n})")
        if ref_frames.dtype != np.complex64 or srv_frames.dtype != np.complex64:
            raise TypeError("frames must be complex64")
        nf = ref_frames.shape[0]
        if out is None:
            out = pinned_empty((nf, self.F, self.R + 1, 1))
        tref = torch.from_numpy(ref_frames)
        tsrv = torch.from_numpy(srv_frames)
        tout = torch.from_numpy(out.reshape(nf, self.F, self.R + 1))
        for i in range(nf):
            k = i % self.nslots
            s = self.streams[k]
            with torch.cuda.stream(s):
                # slot buffers are reused in stream order: the copy below waits for frame i - nslots
                self.ref_d[k].copy_(tref[i], non_blocking=True)
                self.srv_d[k].copy_(tsrv[i], non_blocking=True)
                self._enqueue(self.ref_d[k].data_ptr(), self.srv_d[k].data_ptr(), self.map_d[k].data_ptr(), s)
                tout[i].copy_(self.map_d[k], non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return out

    def process(self, ref, srv):
        """One frame, host arrays in, (F, R+1, 1) map out (convenience; synchronous)."""
        ref = _lib.as_c64(ref, "refChannel")
        srv = _lib.as_c64(srv, "srvChannel")
        return self.run_host(ref[None, :], srv[None, :])[0]
