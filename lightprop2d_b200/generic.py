"""Grids outside the power-of-two kernel sets: any ``npoints`` (Bluestein) and grids above
``lp2d_max_npoints`` (long-line kernels + transposes), still entirely on the device.

The reference accepts every ``npoints`` (beam2d.py:299-308), its ``expand`` produces e.g. 82
points (beam2d.py:385-389) and its own timing script runs 8192^2 complex128
(tests/gaussbeam_performance.py:16-30).  Such a ``Beam2D`` runs here as

* pointwise kernels with the reference's FP64 operation order (``lp2d_pointwise``: transfer
  function / lens phase, ``lp2d_mul_mask``, ``lp2d_intensity_n``), and
* a 2-D DFT engine: ``BluesteinFFT2`` (the n x n DFT as a cyclic convolution of a power-of-two
  size m >= 2n - 1, run by the fused / per-pass ``lp2d_fft2``), or ``LineFFT2`` (one long line per
  thread-block cluster, ``lp2d_pass_lines``, with a tiled transpose between the axes).

There is no CPU fallback here either: everything is CUDA through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _dt(bits):
    return _lib.C64 if bits == 64 else _lib.C128


class _Engine:
    """Shared part: device tables and the pointwise kernels."""

    def __init__(self, n, bits, device, a_tab, b_tab, x2_tab):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.n, self.bits = int(n), bits
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self.cdtype = torch.complex64 if bits == 64 else torch.complex128
        self.rdtype = torch.float32 if bits == 64 else torch.float64
        tabs = np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in (a_tab, b_tab, x2_tab)])
        self.tabs = torch.from_numpy(tabs).to(self.tdev)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    def _call(self, name, *args):
        with self.torch.cuda.device(self.tdev):
            _lib.check(getattr(self.lib, name)(*args), name)

    def _batch(self, t):
        return t.numel() // (self.n * self.n)

    def transfer(self, spectrum, z):
        """spectrum *= exp(i 2 pi (trunc(Kz z) - Kz z)) in place (beam2d.py:485-489)."""
        self._call("lp2d_pointwise", _dt(self.bits), spectrum.data_ptr(), self._batch(spectrum), self.n, self.n, 0,
                   _lib.PW_H, self.tabs.data_ptr(), float(z), 0.0, self._stream())

    def lens(self, field, k0, f):
        """field *= exp(i (X^2 + Y^2) k0 / (2 f)) in place (beam2d.py:508-510)."""
        self._call("lp2d_pointwise", _dt(self.bits), field.data_ptr(), self._batch(field), self.n, self.n, 0,
                   _lib.PW_LENS, self.tabs.data_ptr(), float(k0), 2.0 * float(f), self._stream())

    def mul_mask(self, data, mask, kind, stride):
        self._call("lp2d_mul_mask", _dt(self.bits), data.data_ptr(), mask.data_ptr(), int(kind), self.n * self.n,
                   int(stride), self._batch(data), self._stream())

    def intensity(self, field):
        out = self.torch.empty(field.shape, dtype=self.rdtype, device=self.tdev)
        self._call("lp2d_intensity_n", _dt(self.bits), field.data_ptr(), out.data_ptr(), field.numel(), self._stream())
        return out


class BluesteinFFT2(_Engine):
    """fft2 / ifft2 of (..., n, n) device tensors for ANY n (beam2d.py:207-229 semantics)."""
    kind = "bluestein"

    def __init__(self, n, bits, device, a_tab, b_tab, x2_tab, chunk_bytes=1 << 30):
        super().__init__(n, bits, device, a_tab, b_tab, x2_tab)
        from .beam2d import _Plan
        self.m = int(self.lib.lp2d_bluestein_size(self.n))
        if self.m > int(self.lib.lp2d_max_npoints(_dt(bits))):
            raise NotImplementedError(
                f"npoints={n}: the Bluestein convolution size {self.m} is above lp2d_max_npoints "
                f"({int(self.lib.lp2d_max_npoints(_dt(bits)))}) for complex{bits}")
        z = np.zeros(self.m)
        with self.torch.cuda.device(self.tdev):
            self.plan = _Plan(self.m, bits, z, z, z)          # transforms only: the tables are not used
        self._filters = {}
        self.chunk = max(1, int(chunk_bytes // (self.m * self.m * (8 if bits == 64 else 16))))

    def _filter(self, inverse):
        fb = self._filters.get(inverse)
        if fb is None:
            fb = self.torch.empty(self.m, self.m, dtype=self.cdtype, device=self.tdev)
            self._call("lp2d_bluestein_filter", _dt(self.bits), fb.data_ptr(), self.n, self.m, int(inverse), self._stream())
            self._call("lp2d_fft2", self.plan.handle, fb.data_ptr(), fb.data_ptr(), 1, 0, self._stream())
            self._filters[inverse] = fb
        return fb

    def fft2(self, x, inverse=False):
        x = x.contiguous()
        out = self.torch.empty_like(x)
        B, n, m = self._batch(x), self.n, self.m
        fb = self._filter(bool(inverse))
        xs, os_ = x.reshape(B, n, n), out.reshape(B, n, n)
        nb = min(B, self.chunk)
        y = self.torch.empty(nb, m, m, dtype=self.cdtype, device=self.tdev)
        w = self.torch.empty(nb, m, m, dtype=self.cdtype, device=self.tdev)
        for b0 in range(0, B, nb):
            k = min(nb, B - b0)
            self._call("lp2d_bluestein_pre", _dt(self.bits), xs[b0:b0 + k].data_ptr(), y.data_ptr(), n, m, k, int(inverse), self._stream())
            self._call("lp2d_fft2", self.plan.handle, y.data_ptr(), w.data_ptr(), k, 0, self._stream())
            # w *= FFT(filter);  y = ifft2_m(w): the cyclic convolution
            self._call("lp2d_mul_fft2", self.plan.handle, w.data_ptr(), y.data_ptr(), fb.data_ptr(), _lib.MASK_COMPLEX, 0, k, 1,
                       self._stream())
            self._call("lp2d_bluestein_post", _dt(self.bits), y.data_ptr(), os_[b0:b0 + k].data_ptr(), n, m, k, int(inverse), self._stream())
        return out


class LineFFT2(_Engine):
    """fft2 / ifft2 for power-of-two grids above ``lp2d_max_npoints`` (up to ``lp2d_max_line_points``): row pass of
    the long-line (thread-block cluster) kernel, tiled transpose, row pass, transpose back."""
    kind = "lines"

    def __init__(self, n, bits, device, a_tab, b_tab, x2_tab):
        super().__init__(n, bits, device, a_tab, b_tab, x2_tab)
        from .beam2d import _Plan
        with self.torch.cuda.device(self.tdev):
            self.plan = _Plan(n, bits, a_tab, b_tab, x2_tab)
        self._tmp = None

    def _lines(self, src, dst, inverse):
        self._call("lp2d_pass_lines", self.plan.handle, _lib.OP_FFT, src.data_ptr(), dst.data_ptr(), None, None, 0, self.n,
                   1 if inverse else 0, 0.0, self._stream())

    def _transpose(self, src, dst):
        # one tile of n columns: dst = src^T
        self._call("lp2d_pack_transpose", _dt(self.bits), src.data_ptr(), dst.data_ptr(), self.n, self.n, self.n, self._stream())

    def fft2(self, x, inverse=False):
        x = x.contiguous()
        out = self.torch.empty_like(x)
        B, n = self._batch(x), self.n
        xs, os_ = x.reshape(B, n, n), out.reshape(B, n, n)
        if self._tmp is None:
            self._tmp = self.torch.empty(n, n, dtype=self.cdtype, device=self.tdev)
        t = self._tmp
        for b in range(B):
            self._lines(xs[b], os_[b], inverse)       # along x
            self._transpose(os_[b], t)
            self._lines(t, t, inverse)                # along y (lines of the transposed plane)
            self._transpose(t, os_[b])
        return out


def make_engine(n, bits, device, a_tab, b_tab, x2_tab):
    """The engine for a grid the 2-D plan does not serve, or raise NotImplementedError."""
    lib = _lib.load()
    n = int(n)
    pow2 = n >= 1 and (n & (n - 1)) == 0
    if pow2 and n > int(lib.lp2d_max_npoints(_dt(bits))):
        if n > int(lib.lp2d_max_line_points(_dt(bits))):
            raise NotImplementedError(f"npoints={n} is above lp2d_max_line_points; use SlabBeam2D over several GPUs")
        return LineFFT2(n, bits, device, a_tab, b_tab, x2_tab)
    return BluesteinFFT2(n, bits, device, a_tab, b_tab, x2_tab)
