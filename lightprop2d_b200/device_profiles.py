"""The generators of ``profiles.py`` evaluated on the GPU (``lp2d_generate``, csrc/lp2d_gen.cu).

Same names and positional arguments as ``lightprop2d/profiles.py:35-136``: every generator is called as
``f(X, Y, *args)`` with the 1-D grid ``X`` and the column ``Y`` (numpy, as ``Beam2D`` passes them) and returns a
**CUDA tensor** -- ``Beam2D`` takes it as ``init_field`` / mask without a host round trip.  Keyword-only extras:

``batch``         number of planes (an ensemble); ``None`` -> one ``(N, N)`` plane
``complex_bits``  64 / 128: complex dtype of fields (masks are ``torch.bool``)
``device``        CUDA device index (default: current)
``seed``, ``first_beam``  random profiles: beam ``first_beam + b`` depends on ``(seed, first_beam + b)`` only, not on
                  the batch or the rank that generates it (counter-based Philox4x32-10).  numpy's global stream
                  (``np.random.seed``) is NOT reproduced: seeded parity with the reference uses ``profiles.py``.

Deterministic profiles are the reference's FP64 expressions in its operation order.  ``gaussian_beam`` shifts its
``x`` / ``y`` arguments IN PLACE like the reference (profiles.py:69-70).  Usage::

    from functools import partial
    from lightprop2d_b200 import Beam2D, device_profiles as dp, um, nm
    gen = partial(dp.random_round_hole_phase, batch=4096, complex_bits=64, seed=7)
    beam = Beam2D(200 * um, 512, 632 * nm, init_field_gen=gen, init_gen_args=(25 * um,), complex_bits=64, use_gpu=True)
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ('plane_wave', 'random_wave', 'random_wave_bin', 'gaussian_beam',
           'round_hole', 'random_round_hole', 'random_round_hole_bin',
           'random_round_hole_phase', 'rectangle_hole', 'square_hole',
           'square_slits')

(PLANE_WAVE, RANDOM_WAVE, RANDOM_WAVE_BIN, GAUSSIAN_BEAM, ROUND_HOLE, RANDOM_ROUND_HOLE, RANDOM_ROUND_HOLE_BIN,
 RANDOM_ROUND_HOLE_PHASE, RECTANGLE_HOLE, SQUARE_SLITS) = range(10)
OUT_MASK = 16


def _run(kind, x, y, params=(), *, mask=False, batch=None, complex_bits=128, device=None, seed=0, first_beam=0,
         binning_order=1, out=None):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("lightprop2d_b200.device_profiles needs a CUDA device; the host generators are in profiles.py")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    xs = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    ys = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
    n = xs.size
    if ys.size != n:
        raise ValueError(f"square grids only: len(x) = {n}, len(y) = {ys.size}")
    xd, yd = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
    nb = 1 if batch is None else int(batch)
    shape = (n, n) if batch is None else (nb, n, n)
    if mask:
        out_kind, dtype = OUT_MASK, torch.uint8
    else:
        out_kind, dtype = (_lib.C64, torch.complex64) if complex_bits == 64 else (_lib.C128, torch.complex128)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=dev)
    elif out.dtype != dtype or out.numel() != nb * n * n or not out.is_cuda or not out.is_contiguous():
        raise ValueError("out: wrong dtype, size or device")
    p = (C.c_double * 6)(*([float(v) for v in params] + [0.0] * (6 - len(params))))
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.lp2d_generate(out_kind, out.data_ptr(), n, nb, xd.data_ptr(), yd.data_ptr(), kind, p,
                                     int(binning_order), int(seed), int(first_beam), st), "generate")
    return out.view(torch.bool) if mask else out


def _hole_params(R, x0, y0):
    # the threshold as the host evaluates it (profiles.py:80)
    return (R, x0, y0, 0.0, 0.0, 1 / np.exp(0.5))


def plane_wave(x, y, **kw):
    return _run(PLANE_WAVE, x, y, **kw)


def random_wave_bin(x, y, binning_order=1, **kw):
    """Binary {0, 1} speckle, one value per ``binning_order`` x ``binning_order`` cell (profiles.py:41-50)."""
    return _run(RANDOM_WAVE_BIN, x, y, binning_order=binning_order, **kw)


def random_wave(x, y, binning_order=1, **kw):
    """Uniform [0, 1) speckle (profiles.py:53-61)."""
    return _run(RANDOM_WAVE, x, y, binning_order=binning_order, **kw)


def gaussian_beam(x, y, A0, rho0, x0=0, y0=0, **kw):
    """``A0 * exp(-((x-x0)^2 + (y-y0)^2) / rho0^2)`` (profiles.py:64-71); shifts x, y in place."""
    x -= x0
    y -= y0
    return _run(GAUSSIAN_BEAM, x, y, (A0, rho0), **kw)


def round_hole(x, y, R, x0=0, y0=0, **kw):
    """Boolean disc of radius R/sqrt(2) (profiles.py:74-81)."""
    kw.setdefault("mask", True)
    return _run(ROUND_HOLE, x, y, _hole_params(R, x0, y0), **kw)


def random_round_hole_bin(x, y, R, x0=0, y0=0, binning_order=1, **kw):
    return _run(RANDOM_ROUND_HOLE_BIN, x, y, _hole_params(R, x0, y0), binning_order=binning_order, **kw)


def random_round_hole(x, y, R, x0=0, y0=0, binning_order=1, **kw):
    return _run(RANDOM_ROUND_HOLE, x, y, _hole_params(R, x0, y0), binning_order=binning_order, **kw)


def random_round_hole_phase(x, y, R, x0=0, y0=0, binning_order=1, **kw):
    """Unit-amplitude random phase inside the disc (profiles.py:105-113)."""
    return _run(RANDOM_ROUND_HOLE_PHASE, x, y, _hole_params(R, x0, y0), binning_order=binning_order, **kw)


def rectangle_hole(x, y, dx, dy, x0=0, y0=0, **kw):
    kw.setdefault("mask", True)
    return _run(RECTANGLE_HOLE, x, y, (dx, dy, x0, y0), **kw)


def square_hole(x, y, d, x0=0, y0=0, **kw):
    return rectangle_hole(x, y, d, d, x0, y0, **kw)


def square_slits(x, y, d, slits_distance, x0=0, y0=0, **kw):
    kw.setdefault("mask", True)
    return _run(SQUARE_SLITS, x, y, (d, slits_distance, x0, y0), **kw)
