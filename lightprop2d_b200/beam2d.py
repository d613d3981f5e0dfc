"""Beam2D with the reference's public surface, driven by the sm_100a CUDA backend.

Host-side mirror of ``lightprop2d/beam2d.py:44-757`` (same constructor arguments in
the same positional order, same method names, argument meaning and error
behaviour).  All array maths of the hot path runs in ``liblp2d.so`` through the C
ABI of ``include/lp2d.h``; PyTorch is only the owner of device memory and of the
CUDA stream.  There is no CPU fallback: without a CUDA device or without the
library every constructor raises.

State, as in the reference: ``field`` and ``spectrum`` are both observable after
every call.  ``spectrum`` is materialised lazily (first access, or as a by-product
of the first ``propagate``); its value is the one the reference would hold.

Differences a caller can see (all documented in DESIGN.md):
  * powers of two in [16, 8192] (complex64) / [16, 4096] (complex128) run on the fused / per-pass
    kernel sets; every other ``npoints`` (the reference accepts any, ``expand`` produces e.g. 82) runs
    through Bluestein's algorithm on top of them, larger powers of two (up to 32768) through the
    long-line kernels (``generic.py``) -- all on the device, several times slower per point;
  * with ``use_gpu=True`` the ``field``/``spectrum`` attributes are torch CUDA tensors
    (the analogue of cupy arrays); with ``use_gpu=False`` they are numpy copies, so
    scripts written for the numpy backend run unchanged;
  * ``crop`` and ``lens_image`` raise ``NotImplementedError`` (the reference's own
    implementation raises on SciPy >= 1.14, there is no oracle to match).
"""
from __future__ import annotations

import ctypes as C
import os
from collections.abc import Iterable
from dataclasses import dataclass
from typing import Tuple

import numpy as np

from . import _lib

nm = 1e-7
um = 1e-4
mm = 1e-1
cm = 1
m = 100

_PLAN_CACHE: dict = {}
# numpy dtypes that are uploaded as they are and cast to the complex type on the device
_DEVICE_CAST = ("bool", "uint8", "int8", "int16", "int32", "int64", "float32", "float64")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("lightprop2d_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _Plan:
    """Owner of one lp2d_plan (twiddles + FP64 grid tables on one device)."""

    def __init__(self, n, bits, a_tab, b_tab, x2_tab, spectrum_t=False):
        lib = _lib.load()
        self.handle = C.c_void_p()
        a_tab, b_tab, x2_tab = (np.ascontiguousarray(v, dtype=np.float64) for v in (a_tab, b_tab, x2_tab))
        _lib.check(lib.lp2d_plan_create(C.byref(self.handle), int(n), _lib.C64 if bits == 64 else _lib.C128,
                                        _dptr(a_tab), _dptr(b_tab), _dptr(x2_tab)), "plan_create")
        # transposed spectrum layout (include/lp2d.h): taken where the library offers it, natural layout otherwise
        self.spectrum_t = bool(spectrum_t) and lib.lp2d_plan_set_spectrum_layout(self.handle, 1) == 0

    def __del__(self):
        try:
            if self.handle:
                _lib.load().lp2d_plan_destroy(self.handle)
        except Exception:
            pass


def _get_plan(device_index, n, bits, area_size, wl, a_tab, b_tab, x2_tab, spectrum_t=False):
    # X may have been shifted in place by a generator (reference quirk), so x2 is part of the key
    key = (device_index, n, bits, float(area_size), float(wl), x2_tab.tobytes(), bool(spectrum_t))
    p = _PLAN_CACHE.get(key)
    if p is None:
        if len(_PLAN_CACHE) > 16:
            _PLAN_CACHE.clear()
        torch = _torch()
        with torch.cuda.device(device_index):
            p = _Plan(n, bits, a_tab, b_tab, x2_tab, spectrum_t)
        _PLAN_CACHE[key] = p
    return p


@dataclass
class Beam2D:
    """Electromagnetic field propagation by the angular-spectrum method (see module docstring).

    Parameters are those of the reference (beam2d.py:52-91); lengths in centimetres.
    """

    area_size: float
    npoints: int
    wl: float
    z: float = 0.
    xp: object = np
    init_field: object = None
    init_spectrum: object = None
    init_field_gen: object = None
    init_gen_args: tuple = ()
    complex_bits: int = 128

    use_gpu: bool = False
    numpy_output: bool = True
    device: object = None          # extension: CUDA device index (default: current device)

    # ------------------------------------------------------------------ construction
    def __post_init__(self):
        torch = _torch()
        self._lib = _lib.load()
        if self.device is None:
            self.device = torch.cuda.current_device()
        elif isinstance(self.device, torch.device):
            self.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._tdev = torch.device("cuda", int(self.device))
        self.xp = torch if self.use_gpu else np
        if self.complex_bits == 64:
            self.complex = np.complex64
            self._cdtype, self._rdtype = torch.complex64, torch.float32
        elif self.complex_bits == 128:
            self.complex = np.complex128
            self._cdtype, self._rdtype = torch.complex128, torch.float64
        # (other values leave self.complex unset and fail below with AttributeError, as the reference)
        self._construct_grids()
        if self.init_field_gen is not None:
            field = self.init_field_gen(self.X, self.Y, *self.init_gen_args)
        elif self.init_field is not None:
            field = self.init_field
        else:
            raise ValueError(
                "Init field data is None : " +
                "'init_field_gen' must be a function or 'init_field' must be an array.")
        self._f = self._asdev(field, self.complex)
        self._check_shape(self._f)
        self._new_plan()
        self._s = None if self.init_spectrum is None else self._ingest_spectrum(self.init_spectrum)

    def _construct_grids(self):
        """Grids with the reference's expressions (beam2d.py:299-308); host numpy float64."""
        self.dL = self.area_size / self.npoints
        self.X = np.arange(-self.npoints // 2, self.npoints // 2, 1) * self.dL
        self.Y = self.X.reshape((-1, 1))
        self.k0 = 1 / self.wl
        self.Kx = self._k_grid(self.dL, self.npoints)
        self.Ky = self.Kx.reshape((-1, 1))
        self._Kz = None

    def _k_grid(self, dL: float, npoints: int):
        return np.fft.fftfreq(npoints, d=dL)

    @property
    def Kz(self):
        """Materialised (N, N) float64 Kz of the reference (beam2d.py:307-308); the CUDA
        path never reads it, it is built on first access for callers that index it."""
        if self._Kz is None:
            with np.errstate(invalid="ignore"):
                self._Kz = np.nan_to_num(np.sqrt(self.k0 ** 2 - self.Kx ** 2 - self.Ky ** 2))
        return self._Kz

    def _new_plan(self):
        a_tab = self.k0 ** 2 - self.Kx ** 2            # left-to-right as beam2d.py:308
        b_tab = self.Kx ** 2
        x2_tab = self.X ** 2
        n = int(self.npoints)
        dt = _lib.C64 if self.complex_bits == 64 else _lib.C128
        pow2 = n >= 1 and (n & (n - 1)) == 0
        if pow2 and int(self._lib.lp2d_min_npoints()) <= n <= int(self._lib.lp2d_max_npoints(dt)):
            # LP2D_SPECTRUM_T=1: the spectrum state is kept transposed on the device (every HBM access of a fused step
            # is then a contiguous row); `spectrum` is converted on access, so callers see the reference's orientation
            self._plan = _get_plan(int(self.device), n, self.complex_bits, self.area_size, self.wl, a_tab, b_tab, x2_tab,
                                   os.environ.get("LP2D_SPECTRUM_T", "0") == "1")
            self._gen = None
        else:
            # any other npoints / grids above lp2d_max_npoints: pointwise kernels + a generic 2-D DFT engine
            from . import generic
            key = ("gen", int(self.device), n, self.complex_bits, float(self.area_size), float(self.wl), x2_tab.tobytes())
            eng = _PLAN_CACHE.get(key)
            if eng is None:
                if len(_PLAN_CACHE) > 16:
                    _PLAN_CACHE.clear()
                eng = generic.make_engine(n, self.complex_bits, int(self.device), a_tab, b_tab, x2_tab)
                _PLAN_CACHE[key] = eng
            self._plan = None
            self._gen = eng

    # ------------------------------------------------------------------ host <-> device
    def _np(self, data):
        """Device tensor (or numpy array) -> numpy array (beam2d.py:135-152)."""
        if isinstance(data, np.ndarray):
            return data
        return data.detach().cpu().numpy()

    def _asxp(self, data):
        """numpy array / torch tensor -> array of the exposed backend (beam2d.py:154-187)."""
        torch = _torch()
        if type(data) == np.ndarray:
            return torch.from_numpy(np.ascontiguousarray(data)).to(self._tdev) if self.use_gpu else data
        if isinstance(data, torch.Tensor):
            return data.to(self._tdev) if self.use_gpu else data.detach().cpu().numpy()
        raise ValueError(
            "Unknown type of the given array, not numpy or cupy array")

    def _asdev(self, data, np_complex=None):
        """To a contiguous device tensor, cast like ``astype(self.complex)`` when asked."""
        torch = _torch()
        if isinstance(data, torch.Tensor):
            t = data.to(self._tdev)
        elif isinstance(data, np.ndarray):
            # (a bool / real array is uploaded as it is and cast on the device -- same rounding as ``astype``, but not
            # a complex copy 16x the size of an 8192^2 mask built on the host first)
            if np_complex is not None and data.dtype != np_complex and data.dtype.name not in _DEVICE_CAST:
                data = data.astype(np_complex)
            t = torch.from_numpy(np.ascontiguousarray(data)).to(self._tdev)
        else:
            raise ValueError(
                "Unknown type of the given array, not numpy or cupy array")
        if np_complex is not None and t.dtype != self._cdtype:
            t = t.to(self._cdtype)
        return t.contiguous()

    def _check_shape(self, t):
        if t.dim() < 2 or t.shape[-1] != self.npoints or t.shape[-2] != self.npoints:
            raise ValueError(f"field shape {tuple(t.shape)} does not end in ({self.npoints}, {self.npoints})")

    def _expose(self, t):
        return t if self.use_gpu else self._np(t)

    def _stream(self):
        torch = _torch()
        return C.c_void_p(torch.cuda.current_stream(self._tdev).cuda_stream)

    def _batch(self):
        b = 1
        for d in self._f.shape[:-2]:
            b *= int(d)
        return b

    def _call(self, name, *args):
        torch = _torch()
        if torch.cuda.current_device() == self._tdev.index:       # (the device switch costs a few microseconds per call)
            _lib.check(getattr(self._lib, name)(*args), name)
            return
        with torch.cuda.device(self._tdev):
            _lib.check(getattr(self._lib, name)(*args), name)

    # ---------------------------------------------------------------------- state
    @property
    def field(self):
        return self._expose(self._f)

    @field.setter
    def field(self, value):
        self._f = self._asdev(value, self.complex)

    @property
    def _st(self):
        """True when the spectrum state ``_s`` is held transposed (plan with the transposed spectrum layout)."""
        return self._plan is not None and self._plan.spectrum_t

    def _transposed(self, t):
        out = _torch().empty_like(t)
        n = int(self.npoints)
        self._call("lp2d_transpose", _lib.C64 if self.complex_bits == 64 else _lib.C128, t.data_ptr(), out.data_ptr(),
                   n, t.numel() // (n * n), self._stream())
        return out

    def _ingest_spectrum(self, value):
        """A spectrum in the reference's orientation -> the state ``_s`` in the plan's layout."""
        t = self._asdev(value, self.complex)
        self._check_shape(t)
        return self._transposed(t) if self._st else t

    @property
    def spectrum(self):
        self._materialise_spectrum()
        return self._expose(self._transposed(self._s) if self._st else self._s)

    @spectrum.setter
    def spectrum(self, value):
        self._s = self._ingest_spectrum(value)

    def _materialise_spectrum(self):
        if self._s is None:
            torch = _torch()
            if self._gen is not None:
                self._s = self._gen.fft2(self._f)
                return
            self._s = torch.empty_like(self._f)
            self._call("lp2d_fft2", self._plan.handle, self._f.data_ptr(), self._s.data_ptr(),
                       self._batch(), 0, self._stream())

    def _update_obj(self, field, spectrum=None):
        """Fast re-seed (beam2d.py:283-288): z = 0, new field, spectrum = fft2(field) or the given one."""
        self.z = 0
        self._f = self._asdev(field, None)
        if self._f.dtype != self._cdtype:
            self._f = self._f.to(self._cdtype)
        self._check_shape(self._f)
        self._s = None if spectrum is None else self._ingest_spectrum(spectrum)

    # ------------------------------------------------------------------ hot path
    def propagate(self, z: float):
        """beam2d.py:481-490."""
        if z == 0:
            return
        torch = _torch()
        self.z += z
        if self._gen is not None:
            self._materialise_spectrum()
            if self._s.data_ptr() == self._f.data_ptr():
                self._s = self._s.clone()
            self._gen.transfer(self._s, z)                     # spectrum *= H, in place (beam2d.py:489)
            self._f = self._gen.fft2(self._s, inverse=True)    # :490
            return
        if self._s is None:
            # first step after a coordinate-space update: fused fft2 -> H -> ifft2, the
            # spectrum state (fft2(field)*H) is written as a by-product
            self._s = torch.empty_like(self._f)
            out = torch.empty_like(self._f)
            self._call("lp2d_propagate_field", self._plan.handle, self._f.data_ptr(), out.data_ptr(),
                       self._s.data_ptr(), None, None, 0, self._batch(), float(z), self._stream())
            self._f = out
        else:
            out = torch.empty_like(self._s)
            self._call("lp2d_propagate", self._plan.handle, self._s.data_ptr(), out.data_ptr(), None,
                       self._batch(), float(z), self._stream())
            self._f = out

    def propagate_sweep(self, dz: float, nsteps: int, probe="centre", moments: bool = False):
        """``nsteps`` x ``propagate(dz)`` enqueued by ONE library call (extension; the reference's sweeps are Python
        loops that read ``iprofile[ny, nx]`` or ``centroid`` / ``D4sigma`` every step: tests/test_analytics.py:43-50,
        tests/gaussbeam_diffraction.py:43-49).  Returns ``(I, M)``: ``I[k]`` = ``iprofile[probe]`` after step k
        (``probe`` = "centre" -> ``[N/2, N/2]``, an ``(iy, ix)`` pair, or None), and with ``moments=True`` a dict with
        ``centroid`` (nsteps, 2) and ``D4sigma`` (nsteps, 2) per step, else None.  One synchronisation at the end
        instead of one per step.  The state after the call is that of the loop (``z`` is advanced ``nsteps`` times)."""
        torch = _torch()
        nsteps = int(nsteps)
        if dz == 0 or nsteps <= 0:
            return None, None
        if self._gen is not None or probe is None and not moments:
            for _ in range(nsteps):
                self.propagate(dz)
            return None, None
        n, B = self.npoints, self._batch()
        if self._s is None:
            self._materialise_spectrum()
        if probe == "centre":
            probe = (n // 2, n // 2)
        pidx = -1 if probe is None else int(probe[0]) * n + int(probe[1])
        pout = None if probe is None else torch.empty(nsteps, B, dtype=self._rdtype, device=self._tdev)
        mout = torch.empty(nsteps, B, 5, dtype=torch.float64, device=self._tdev) if moments else None
        out = torch.empty_like(self._s)
        self._call("lp2d_propagate_sweep", self._plan.handle, self._s.data_ptr(), out.data_ptr(), B, nsteps, float(dz), pidx,
                   None if pout is None else pout.data_ptr(), float(self.dL), None if mout is None else mout.data_ptr(),
                   self._stream())
        self._f = out
        for _ in range(nsteps):
            self.z += dz
        I = None if pout is None else self._output(pout.reshape(nsteps, *self._f.shape[:-2]))
        M = None
        if moments:
            m = mout.cpu().numpy().reshape(nsteps, *self._f.shape[:-2], 5)
            xc, yc = m[..., 1] / m[..., 0], m[..., 2] / m[..., 0]
            with np.errstate(invalid="ignore"):
                sx = np.sqrt(np.maximum(m[..., 3] / m[..., 0] - xc ** 2, 0.0))
                sy = np.sqrt(np.maximum(m[..., 4] / m[..., 0] - yc ** 2, 0.0))
            M = {"centroid": np.stack([xc, yc], axis=-1), "D4sigma": 4 * np.stack([sx, sy], axis=-1)}
        return I, M

    def lens(self, f: float):
        """beam2d.py:508-511."""
        torch = _torch()
        if self._gen is not None:
            self._gen.lens(self._f, self.k0, f)
            self._s = self._gen.fft2(self._f)
            return
        self._s = torch.empty_like(self._f)
        self._call("lp2d_lens", self._plan.handle, self._f.data_ptr(), self._s.data_ptr(), self._batch(),
                   float(self.k0), float(f), self._stream())

    def _mask_to_device(self, mask):
        """Mask -> (tensor, kind, beam stride).  bool -> bytes, real -> field's real type,
        complex -> field's complex type (numpy's in-place multiply casts the same way)."""
        torch = _torch()
        if isinstance(mask, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(mask)).to(self._tdev)
        elif isinstance(mask, torch.Tensor):
            t = mask.to(self._tdev)
        else:
            t = torch.as_tensor(np.asarray(mask)).to(self._tdev)
        if t.dtype == torch.bool:
            t, kind = t.to(torch.uint8), _lib.MASK_U8
        elif t.is_complex():
            t, kind = t.to(self._cdtype), _lib.MASK_COMPLEX
        else:
            t, kind = t.to(self._rdtype), _lib.MASK_REAL
        n = self.npoints
        if t.shape[-2:] != (n, n):
            t = t.expand(*t.shape[:-2], n, n) if t.dim() >= 2 else t.expand(n, n)
        if t.dim() == 2:
            stride = 0
        else:
            t = t.expand(*self._f.shape)
            stride = n * n
        return t.contiguous(), kind, stride

    def _apply_masks(self, masks, spectral):
        torch = _torch()
        if spectral:
            self._materialise_spectrum()
            data = self._s
        else:
            data = self._f
        if self._gen is not None:
            for mask in masks:
                mt, kind, stride = self._mask_to_device(mask)
                self._gen.mul_mask(data, mt, kind, stride)
            if spectral:
                self._f = self._gen.fft2(data, inverse=True)
            else:
                self._s = self._gen.fft2(data)
            return
        other = torch.empty_like(data)
        if not masks:
            masks = [None]
        for mask in masks:
            if mask is None:
                mt, kind, stride, ptr = None, 0, 0, None
            else:
                mt, kind, stride = self._mask_to_device(mask)
                if spectral and self._st:                  # the spectrum is held transposed: so must its mask be
                    mt = mt.transpose(-1, -2).contiguous()
                ptr = mt.data_ptr()
            self._call("lp2d_mul_fft2", self._plan.handle, data.data_ptr(), other.data_ptr(), ptr, kind,
                       stride, self._batch(), 1 if spectral else 0, self._stream())
        if spectral:
            self._f = other
        else:
            self._s = other

    def coordinate_filter(self, f_init=None, f_gen=None, fargs=()) -> None:
        """beam2d.py:335-342: field *= f_gen(X, Y, *fargs) and/or f_init; spectrum = fft2(field)."""
        masks = []
        if f_gen is not None:
            assert callable(f_gen)
            assert isinstance(fargs, Iterable)
            masks.append(f_gen(self.X, self.Y, *fargs))
        if f_init is not None:
            assert isinstance(f_init, (np.ndarray, _torch().Tensor))
            masks.append(f_init)
        self._apply_masks(masks, spectral=False)

    def spectral_filter(self, f_init=None, f_gen=None, fargs=()) -> None:
        """beam2d.py:366-373: spectrum *= f_gen(Kx, Ky, *fargs) and/or f_init; field = ifft2(spectrum)."""
        masks = []
        if f_gen is not None:
            assert callable(f_gen)
            assert isinstance(fargs, Iterable)
            masks.append(f_gen(self.Kx, self.Ky, *fargs))
        if f_init is not None:
            assert isinstance(f_init, (np.ndarray, _torch().Tensor))
            masks.append(f_init)
        self._apply_masks(masks, spectral=True)

    # -------------------------------------------------------------- expand / coarse
    def expand(self, area_size: float):
        """beam2d.py:385-404: grow the area at constant dL with edge-replicated samples."""
        torch = _torch()
        old_npoints = self.npoints
        npoints = int(self.npoints * area_size / self.area_size)
        npoints -= npoints % 2
        d_npoints = (npoints - old_npoints) // 2
        # validate before any attribute changes: a caught exception must leave the beam as it was
        if d_npoints < 0:
            raise ValueError("count < 0")              # numpy.repeat with a negative count
        n_arr = old_npoints + 2 * d_npoints            # size of the array the reference builds
        if n_arr != npoints:
            # latent reference bug (odd difference): the array and npoints disagree; the
            # reference's fft2 then runs on the array size.  Not reproducible with one plan.
            raise NotImplementedError("expand: npoints difference is odd; array and grid sizes disagree")
        saved = (self.npoints, self.area_size)
        self.npoints, self.area_size = npoints, area_size
        try:
            self._construct_grids()
            self._new_plan()
        except Exception:
            self.npoints, self.area_size = saved
            self._construct_grids()
            raise
        out = torch.empty(*self._f.shape[:-2], n_arr, n_arr, dtype=self._cdtype, device=self._tdev)
        self._call("lp2d_expand", _lib.C64 if self.complex_bits == 64 else _lib.C128, self._f.data_ptr(),
                   out.data_ptr(), old_npoints, n_arr, self._batch(), self._stream())
        self._f = out
        self._s = None

    @staticmethod
    def _linear_knots(xs, xq):
        """Interval index and the two weights of the degree-1 spline FITPACK evaluates for the reference's ``interp2d``
        (kind='linear'): arguments outside [xs[0], xs[-1]] are clamped (fpbisp), interval l with xs[l] <= x < xs[l+1]
        (the last one closed), weights (xs[l+1] - x) / h and (x - xs[l]) / h.  2 n numbers, built on the host."""
        xs = np.asarray(xs, dtype=np.float64)
        x = np.clip(np.asarray(xq, dtype=np.float64), xs[0], xs[-1])
        l = np.clip(np.searchsorted(xs, x, side="right") - 1, 0, len(xs) - 2)
        h = xs[l + 1] - xs[l]
        w = np.stack([(xs[l + 1] - x) / h, (x - xs[l]) / h])
        return l.astype(np.int32), np.ascontiguousarray(w)

    def _resample(self, src, src_plane, ld_in, n_out, xs, xq):
        """Bilinear resampling of (a window of) ``src`` onto the mesh xq x xq, on the device (``lp2d_resample``)."""
        torch = _torch()
        idx, w = self._linear_knots(xs, xq)
        d_idx = torch.from_numpy(idx).to(self._tdev)
        d_w = torch.from_numpy(w).to(self._tdev)
        out = torch.empty(*self._f.shape[:-2], n_out, n_out, dtype=self._cdtype, device=self._tdev)
        self._call("lp2d_resample", _lib.C64 if self.complex_bits == 64 else _lib.C128, src.data_ptr(), int(src_plane), int(ld_in),
                   out.data_ptr(), int(n_out), d_idx.data_ptr(), d_w.data_ptr(), d_idx.data_ptr(), d_w.data_ptr(), self._batch(),
                   self._stream())
        return out

    def crop(self, area_size: float, npoints: int = 0):
        """beam2d.py:419-435: keep the window ``|X| < area_size / 2`` of the field and resample it (linear
        ``interp2d`` in the reference: a tensor-product linear spline, evaluated here on the device) onto the new
        ``npoints`` grid of the new area.  The reference itself cannot run this on SciPy >= 1.14 (``interp2d`` was
        removed); parity is defined against ``RectBivariateSpline(kx=1, ky=1)``, SciPy's documented replacement."""
        old_X = self.X[np.abs(self.X) < area_size / 2]
        Nc = self.npoints // 2
        n = len(old_X) // 2
        if n < 1:
            raise ValueError("crop: the new area holds fewer than two grid points")
        old_X = self.X[Nc - n:Nc + n].copy()
        old_n = self.npoints
        saved = (self.npoints, self.area_size)
        if npoints:
            self.npoints = npoints
        self.area_size = area_size
        try:
            self._construct_grids()
            self._new_plan()
        except Exception:
            self.npoints, self.area_size = saved
            self._construct_grids()
            raise
        win = self._f.reshape(-1, old_n, old_n)[:, Nc - n:, Nc - n:]          # view: window origin, full row stride
        self._f = self._resample(win, old_n * old_n, old_n, self.npoints, old_X, self.X)
        self._s = None

    def coarse(self, mean_order: int = 1):
        """beam2d.py:454-460: npoints //= m, block mean of the field, fft2."""
        torch = _torch()
        old = self.npoints
        npoints = old // mean_order
        if npoints * mean_order != old:
            raise ValueError(f"cannot reshape array of size {old * old} into shape "
                             f"({npoints},{mean_order},{npoints},{mean_order})")
        self.npoints = npoints
        try:
            self._construct_grids()
            self._new_plan()
        except Exception:
            self.npoints = old
            self._construct_grids()
            raise
        out = torch.empty(*self._f.shape[:-2], self.npoints, self.npoints, dtype=self._cdtype, device=self._tdev)
        self._call("lp2d_coarse", _lib.C64 if self.complex_bits == 64 else _lib.C128, self._f.data_ptr(),
                   out.data_ptr(), old, int(mean_order), self._batch(), self._stream())
        self._f = out
        self._s = None

    def lens_image(self, f: float, l1: float, l2: float):
        """beam2d.py:525-536: image through a lens between conjugated planes.  Reference quirks kept: ``X`` (and,
        through the view, ``Y``) is scaled IN PLACE by ``-l1 / l2`` and stays scaled; ``interp2d.__call__`` sorts its
        arguments, so the mesh is ascending whatever the sign of the scale; the phase uses the grid before the
        scaling.  Resampling on the device, parity against ``RectBivariateSpline(kx=1, ky=1)`` as for ``crop``."""
        torch = _torch()
        scale_factor = - l1 / l2
        X = self.X
        xs = X.copy()
        phase = - (self.X ** 2 + self.Y ** 2) * self.k0 * (l1 + l2) / 2 / l2 ** 2
        X *= scale_factor                                     # in place: self.X and self.Y change (beam2d.py:531)
        n = self.npoints
        res = self._resample(self._f, n * n, n, n, xs, np.sort(X))
        mask = scale_factor * (np.cos(phase) + 1j * np.sin(phase))
        mt, kind, stride = self._mask_to_device(mask)
        self._f = res
        self._new_plan()                                      # the grid changed: lens() after this uses the scaled X
        self._call("lp2d_mul_mask", _lib.C64 if self.complex_bits == 64 else _lib.C128, self._f.data_ptr(), mt.data_ptr(), kind,
                   n * n, stride, self._batch(), self._stream())
        self._s = None

    # -------------------------------------------------------------------- modes
    def _modes_to_device(self, modes_list):
        torch = _torch()
        if isinstance(modes_list, torch.Tensor):
            mm_ = modes_list.to(self._tdev)
        else:
            mm_ = torch.from_numpy(np.ascontiguousarray(np.asarray(modes_list))).to(self._tdev)
        if mm_.dim() != 2:
            mm_ = mm_.reshape(mm_.shape[0], -1)
        return mm_.to(self._cdtype).contiguous()

    def deconstruct_by_modes(self, modes_list, mode: str = 'scalar'):
        """beam2d.py:592-604, 'scalar' branch on the GPU: coeffs = M @ vec(centre crop), no conjugate."""
        torch = _torch()
        if mode == 'lstsq':
            return self._deconstruct_lstsq(modes_list)
        if mode != 'scalar':
            return getattr(self, "modes_coeffs", None)        # the reference falls through both branches
        self.modes_coeffs = self._expose(self._project_dev(modes_list))
        return self.modes_coeffs

    def _project_dev(self, modes_list):
        torch = _torch()
        M = self._modes_to_device(modes_list)
        nb = int(np.sqrt(M.shape[1]))
        out = torch.empty(*self._f.shape[:-2], M.shape[0], dtype=self._cdtype, device=self._tdev)
        self._call("lp2d_modes_project", _lib.C64 if self.complex_bits == 64 else _lib.C128, self._f.data_ptr(),
                   M.data_ptr(), out.data_ptr(), int(self.npoints), nb, int(M.shape[0]), self._batch(),
                   self._stream())
        return out

    def _project128(self, fields128, M128, n, nb):
        torch = _torch()
        batch = fields128.numel() // (n * n)
        out = torch.empty(batch, M128.shape[0], dtype=torch.complex128, device=self._tdev)
        self._call("lp2d_modes_project", _lib.C128, fields128.data_ptr(), M128.data_ptr(), out.data_ptr(), int(n), int(nb),
                   int(M128.shape[0]), batch, self._stream())
        return out

    def _solve(self, A, B):
        """A X = B on the device (``lp2d_solve``: LU with partial pivoting, complex128); returns X (M, nrhs)."""
        torch = _torch()
        A = A.to(torch.complex128).contiguous().clone()
        B = B.to(torch.complex128).contiguous().clone()
        info = torch.zeros(1, dtype=torch.int32, device=self._tdev)
        self._call("lp2d_solve", A.data_ptr(), B.data_ptr(), int(A.shape[0]), int(B.shape[1]), info.data_ptr(), self._stream())
        return B

    def _deconstruct_lstsq(self, modes_list):
        """beam2d.py:601-603: ``lstsq(modes_matrix.T, vec(crop), rcond=-1)[0]``, here through the normal equations on
        the device in complex128 whatever the field type: G = conj(M) M^T and conj(M) f come from the mode projection
        kernel (the rows of conj(M) are projected like a batch of fields), the M x M solve is ``lp2d_solve``.  Equal to
        the reference's SVD solution when the modes are linearly independent (cond(M)^2 eps accuracy); a rank-deficient
        basis, where numpy returns the minimum-norm solution, is reported by a ValueError."""
        torch = _torch()
        M = self._modes_to_device(modes_list).to(torch.complex128)
        nm, nb = int(M.shape[0]), int(np.sqrt(M.shape[1]))
        G = self._project128(M.conj().resolve_conj().contiguous(), M, nb, nb)             # G[l][k] = sum conj(M_l) M_k
        rhs = self._project128(self._f.to(torch.complex128).conj().resolve_conj().contiguous(), M, self.npoints, nb).conj()
        X = self._solve(G, rhs.reshape(-1, nm).transpose(0, 1))
        sol = X.transpose(0, 1).reshape(*self._f.shape[:-2], nm).to(self._cdtype)
        if not bool(torch.isfinite(torch.view_as_real(sol)).all()):
            raise ValueError("deconstruct_by_modes(mode='lstsq'): the mode basis is rank deficient")
        self.modes_coeffs = self._expose(sol.contiguous())
        return self.modes_coeffs

    def fast_deconstruct_by_modes(self, modes_matrix_t, modes_matrix_dot_t):
        """beam2d.py:640-650: ``lstsq(MtM, Mt . vec(centre crop), rcond=None)[0]``.  ``Mt`` is the (M, K) stack of
        flattened modes, ``MtM`` the (M, M) system matrix.  The projection ``Mt . f`` (M x K, the expensive
        part) runs on the GPU (tensor cores for complex64); the M x M least-squares solve is the reference's own
        ``numpy.linalg.lstsq`` call on the host (tiny, LAPACK)."""
        torch = _torch()
        keep = self.numpy_output, self.use_gpu
        proj = self.deconstruct_by_modes(modes_matrix_t, mode='scalar')
        proj = proj if isinstance(proj, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(proj)).to(self._tdev)
        mtm = modes_matrix_dot_t
        mtm = (mtm if isinstance(mtm, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(mtm)))).to(self._tdev)
        nm = int(proj.shape[-1])
        X = self._solve(mtm, proj.reshape(-1, nm).transpose(0, 1))           # one right-hand side per beam
        sol = X.transpose(0, 1).reshape(proj.shape).to(self._cdtype).contiguous()
        self.modes_coeffs = self._expose(sol)
        return self.modes_coeffs

    def construct_by_modes(self, modes_list, modes_coeffs):
        """beam2d.py:664-667: field = sum_i c_i * zero-padded mode_i; spectrum = fft2(field)."""
        torch = _torch()
        M = self._modes_to_device(modes_list)
        nb = int(np.sqrt(M.shape[1]))
        if isinstance(modes_coeffs, torch.Tensor):
            c = modes_coeffs.to(self._tdev).to(self._cdtype).contiguous()
        else:
            c = torch.from_numpy(np.ascontiguousarray(np.asarray(modes_coeffs))).to(self._tdev).to(self._cdtype)
        batch = 1
        for d in c.shape[:-1]:
            batch *= int(d)
        out = torch.empty(*c.shape[:-1], self.npoints, self.npoints, dtype=self._cdtype, device=self._tdev)
        self._call("lp2d_modes_synthesize", _lib.C64 if self.complex_bits == 64 else _lib.C128, c.data_ptr(),
                   M.data_ptr(), out.data_ptr(), int(self.npoints), nb, int(M.shape[0]), batch, self._stream())
        self._f = out
        self._s = None

    # -------------------------------------------------------------- observables
    def _output(self, t):
        return self._np(t) if self.numpy_output else t

    @property
    def iprofile(self):
        """|A|^2 (beam2d.py:717): float32 for complex64, float64 for complex128."""
        torch = _torch()
        if self._gen is not None:
            return self._output(self._gen.intensity(self._f))
        out = torch.empty(self._f.shape, dtype=self._rdtype, device=self._tdev)
        self._call("lp2d_intensity", self._plan.handle, self._f.data_ptr(), out.data_ptr(), self._batch(),
                   self._stream())
        return self._output(out)

    @property
    def phiprofile(self):
        """arg(A) (beam2d.py:734)."""
        torch = _torch()
        out = torch.empty(self._f.shape, dtype=self._rdtype, device=self._tdev)
        self._call("lp2d_phase", _lib.C64 if self.complex_bits == 64 else _lib.C128, self._f.data_ptr(),
                   out.data_ptr(), self._f.numel(), self._stream())
        return self._output(out)

    def _moments(self, xc=0.0, yc=0.0):
        """[sum I, sum x I, sum y I, sum (x-xc)^2 I, sum (y-yc)^2 I] in FP64, x = (i - n//2) dL."""
        torch = _torch()
        if self._batch() != 1:
            raise ValueError("centroid / D4sigma are defined for a single field, as in the reference")
        if self.npoints % 2:
            # beam2d.py:680-683: mgrid[-n//2 : n//2] has npoints - 1 points per axis, numpy.average refuses the weights
            raise TypeError("Axis must be specified when shapes of a and weights differ.")
        out = torch.empty(5, dtype=torch.float64, device=self._tdev)
        self._call("lp2d_moments", _lib.C64 if self.complex_bits == 64 else _lib.C128, self._f.data_ptr(),
                   int(self.npoints), float(self.dL), float(xc), float(yc), out.data_ptr(), self._stream())
        return out.cpu().numpy()

    @property
    def centroid(self) -> Tuple[float, float, int, int]:
        """beam2d.py:680-685."""
        n = self.npoints // 2
        s = self._moments()
        Xc = s[1] / s[0]
        Yc = s[2] / s[0]
        return Xc, Yc, int(Xc / self.dL) + n, int(Yc / self.dL) + n

    @property
    def D4sigma(self) -> Tuple[float, float]:
        """beam2d.py:696-704."""
        Xc, Yc, _, _ = self.centroid
        s = self._moments(Xc, Yc)
        return 4 * np.sqrt(s[3] / s[0]).tolist(), 4 * np.sqrt(s[4] / s[0]).tolist()

    @property
    def centroid_intensity(self) -> float:
        """beam2d.py:751-752: the intensity at the centroid sample."""
        _, _, nxc, nyc = self.centroid
        v = self._f.reshape(self.npoints, self.npoints)[nyc, nxc].cpu().numpy()
        return (np.abs(v) ** 2).tolist()

    def __repr__(self) -> str:
        return (f"Beam {self.npoints:d}x{self.npoints:d} points" +
                f"\n\t{self.area_size:.3g}x{self.area_size:.3g} cm " +
                f"\n\t<wl={self.wl / nm:.3g} nm, z={self.z:.3g} cm>")
