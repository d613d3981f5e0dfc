"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference hot path.

This is a CPU restatement (numpy, float64 grids, pocketfft through ``numpy.fft``)
of the algorithm the reference implements in ``lightprop2d/beam2d.py``.  Each
function cites the reference lines it follows.  It is written as stateless
functions over plain arrays plus one small stateful wrapper (``OracleBeam``) so the
parity tests can read like reference user scripts.

Parity pin: the reference holds no golden vectors for this path.
The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, generated
in the build container by ``oracle/make_golden.py`` (which imports the unmodified
reference from /root/reference through ``oracle/ref_shim.py``) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every fixture bit-for-bit
(grids) / to <= 4 ulp-level tolerance (FFT results on the same numpy build are
bit-identical; the tolerance only guards against a different numpy on the GPU box).

PARITY UNPINNED for ``crop`` and ``lens_image`` only: the reference resamples with
``scipy.interpolate.interp2d`` (beam2d.py:427-434, 527-535), which SciPy removed in 1.14 (the reference
raises NotImplementedError here, 1.18).  Those two functions are restated with SciPy's documented
replacement, ``RectBivariateSpline(x, y, z.T, kx=1, ky=1)`` (same FITPACK routines: regrid + bispev), plus
the two behaviours of the removed class that matter: ``__call__`` sorted its arguments, and arguments
outside the data range are clamped.

The third-party arithmetic on the path: numpy (unpinned by the
reference; 2.3.5 here) ``numpy.fft.fft2/ifft2`` (pocketfft, forward unscaled,
inverse scaled by 1/N^2), ``numpy.fft.fftfreq`` (``k * (1.0/(n*d))`` with integer
k = 0..(n-1)//2, -(n//2)..-1), ``sqrt``, ``nan_to_num``, ``trunc``, ``cos``, ``sin``.
"""
from __future__ import annotations

import numpy as np

nm, um, mm, cm, m = 1e-7, 1e-4, 1e-1, 1, 100  # beam2d.py:37-41 (base unit: cm)


# --------------------------------------------------------------------------- grids
def make_grids(area_size: float, npoints: int, wl: float) -> dict:
    """Coordinate / wave-vector grids.  Follows beam2d.py:299-308 and :205.

    X is centred with floor division of the NEGATIVE count (odd n => one more
    negative sample), Kx is numpy's fftfreq in cycles/cm (no 2*pi), Kz is the
    materialised (n, n) float64 array with evanescent NaNs mapped to 0.
    """
    dL = area_size / npoints                                    # :299
    X = np.arange(-npoints // 2, npoints // 2, 1) * dL          # :300-301
    Y = X.reshape((-1, 1))                                      # :302
    k0 = 1 / wl                                                 # :304
    Kx = np.fft.fftfreq(npoints, d=dL)                          # :305 -> :205
    Ky = Kx.reshape((-1, 1))                                    # :306
    with np.errstate(invalid="ignore"):
        Kz = np.nan_to_num(np.sqrt(k0 ** 2 - Kx ** 2 - Ky ** 2))  # :307-308
    return dict(dL=dL, X=X, Y=Y, k0=k0, Kx=Kx, Ky=Ky, Kz=Kz)


def kz2_tables(area_size: float, npoints: int, wl: float):
    """The two 1-D tables the CUDA path uploads instead of the (n, n) Kz array.

    ``a[i] = k0**2 - Kx[i]**2`` and ``b[j] = Ky[j]**2`` so that
    ``Kz[j, i] = nan_to_num(sqrt(a[i] - b[j]))`` reproduces the broadcasting and
    left-to-right evaluation order of beam2d.py:307-308 bit for bit.
    """
    g = make_grids(area_size, npoints, wl)
    a = g["k0"] ** 2 - g["Kx"] ** 2
    b = g["Kx"] ** 2
    return a, b


# ----------------------------------------------------------------------------- FFTs
def fft2(a: np.ndarray) -> np.ndarray:
    """beam2d.py:207-217 (numpy branch): last two axes, forward, unscaled."""
    return np.fft.fft2(a)


def ifft2(a: np.ndarray) -> np.ndarray:
    """beam2d.py:219-229 (numpy branch): last two axes, inverse, 1/N^2."""
    return np.fft.ifft2(a)


# ------------------------------------------------------------------------ propagate
def transfer_phase(Kz: np.ndarray, z: float) -> np.ndarray:
    """Range-reduced propagation phase.  beam2d.py:485-487.

    ``2*pi*(trunc(Kz*z) - Kz*z)`` keeps only the fractional number of cycles, in
    float64 regardless of the field dtype.
    """
    deltak = Kz * z
    return 2 * np.pi * (np.trunc(deltak) - deltak)


def propagate(spectrum: np.ndarray, Kz: np.ndarray, z: float):
    """One stateful step.  beam2d.py:481-490.

    Returns ``(spectrum', field')``; ``spectrum`` is updated IN PLACE (the in-place
    multiply down-casts the complex128 factor into a complex64 spectrum, as the
    reference does).  ``z == 0`` is a no-op (returns ``(spectrum, None)``).
    """
    if z == 0:                                                  # :481-482
        return spectrum, None
    ph = transfer_phase(Kz, z)
    spectrum *= np.cos(ph) + 1j * np.sin(ph)                    # :489
    return spectrum, ifft2(spectrum)                            # :490


# ----------------------------------------------------------------------------- lens
def lens(field: np.ndarray, X: np.ndarray, Y: np.ndarray, k0: float, f: float):
    """Thin-lens phase.  beam2d.py:508-511 (k0 = 1/wl, no 2*pi: reproduced as is).

    ``field`` is multiplied in place; returns ``(field, spectrum)``.
    """
    ph = (X ** 2 + Y ** 2) * k0 / (2 * f)                       # :508
    field *= np.cos(ph) + 1j * np.sin(ph)                       # :510
    return field, fft2(field)                                   # :511


# -------------------------------------------------------------------------- filters
def coordinate_filter(field, X, Y, f_init=None, f_gen=None, fargs=()):
    """beam2d.py:335-342: multiply the field by a generated and/or given mask, fft2."""
    if f_gen is not None:
        field *= f_gen(X, Y, *fargs)
    if f_init is not None:
        field *= f_init
    return field, fft2(field)


def spectral_filter(spectrum, Kx, Ky, f_init=None, f_gen=None, fargs=()):
    """beam2d.py:366-373: multiply the spectrum by a mask, ifft2."""
    if f_gen is not None:
        spectrum *= f_gen(Kx, Ky, *fargs)
    if f_init is not None:
        spectrum *= f_init
    return spectrum, ifft2(spectrum)


# ------------------------------------------------------------------- expand / coarse
def expand_shape(npoints: int, area_size: float, new_area: float):
    """New point count and per-side pad of beam2d.py:385-389."""
    n_new = int(npoints * new_area / area_size)
    n_new -= n_new % 2
    return n_new, (n_new - npoints) // 2


def expand(field: np.ndarray, pad: int) -> np.ndarray:
    """Edge-replicating pad of beam2d.py:392-401 (first/last column, then first/last row)."""
    if pad <= 0:
        # np.repeat(..., 0) yields empty blocks: array unchanged (reference behaviour)
        return field.copy() if pad == 0 else field
    return np.pad(field, [(0, 0)] * (field.ndim - 2) + [(pad, pad), (pad, pad)], mode="edge")


def coarse(field: np.ndarray, mean_order: int) -> np.ndarray:
    """Block mean of beam2d.py:454-459 (npoints //= m, reshape, mean over the m-axes)."""
    n = field.shape[-1] // mean_order
    return field.reshape((n, mean_order, n, mean_order)).mean((1, 3))


# ---------------------------------------------------------------------- observables
def iprofile(field: np.ndarray) -> np.ndarray:
    """|A|^2, beam2d.py:717 (float32 for complex64, float64 for complex128)."""
    return np.abs(field) ** 2


def phiprofile(field: np.ndarray) -> np.ndarray:
    """arg(A), beam2d.py:734."""
    return np.angle(field)


def centroid(field: np.ndarray, dL: float):
    """Intensity-weighted mean position.  beam2d.py:680-685."""
    n = field.shape[-1] // 2
    Y, X = np.mgrid[-n:n, -n:n] * dL
    I = iprofile(field)
    Xc = np.average(X, weights=I)
    Yc = np.average(Y, weights=I)
    return Xc, Yc, int(Xc / dL) + n, int(Yc / dL) + n


def d4sigma(field: np.ndarray, dL: float):
    """4-sigma widths.  beam2d.py:696-704."""
    n = field.shape[-1] // 2
    I = iprofile(field)
    Y, X = np.mgrid[-n:n, -n:n] * dL
    Xc, Yc, _, _ = centroid(field, dL)
    sx = np.sqrt(np.average((X - Xc) ** 2, weights=I)).tolist()
    sy = np.sqrt(np.average((Y - Yc) ** 2, weights=I)).tolist()
    return 4 * sx, 4 * sy


def centroid_intensity(field: np.ndarray, dL: float) -> float:
    """beam2d.py:751-752."""
    _, _, nxc, nyc = centroid(field, dL)
    return iprofile(field)[nyc, nxc].tolist()


# ---------------------------------------------------------------------------- modes
def modes_project(field: np.ndarray, modes: np.ndarray) -> np.ndarray:
    """Scalar mode projection, beam2d.py:592-600: centre crop, flatten, M @ f (NO conjugate)."""
    modes = np.asarray(modes)
    npoints = field.shape[-1]
    Nb = int(np.sqrt(modes.shape[1]))
    dN = (npoints - Nb) // 2
    flat = field[dN:npoints - dN, dN:npoints - dN].flatten()
    return np.vstack(modes) @ flat


def modes_fast_lstsq(field: np.ndarray, modes_matrix_t: np.ndarray, modes_matrix_dot_t: np.ndarray) -> np.ndarray:
    """fast_deconstruct_by_modes (beam2d.py:640-650): ``lstsq(MtM, Mt . vec(centre crop), rcond=None)[0]`` with
    ``Mt`` of shape (M, K) (the stack of flattened modes) and ``MtM`` the (M, M) system matrix."""
    modes_matrix_t = np.asarray(modes_matrix_t)
    Nb = int(np.sqrt(modes_matrix_t.shape[1]))
    n = field.shape[-1]
    dN = (n - Nb) // 2
    flat = field[dN:n - dN, dN:n - dN].flatten()
    return np.linalg.lstsq(np.asarray(modes_matrix_dot_t), modes_matrix_t.dot(flat), rcond=None)[0]


def modes_lstsq(field: np.ndarray, modes: np.ndarray) -> np.ndarray:
    """deconstruct_by_modes(mode='lstsq'), beam2d.py:601-603: lstsq(modes_matrix.T, vec(centre crop), rcond=-1)[0]."""
    modes = np.vstack(np.asarray(modes))
    n = field.shape[-1]
    Nb = int(np.sqrt(modes.shape[1]))
    dN = (n - Nb) // 2
    flat = field[dN:n - dN, dN:n - dN].flatten()
    return np.linalg.lstsq(modes.T, flat, rcond=-1)[0]


def _interp2d_linear(xs, z, xq):
    """What ``interp2d(xs, xs, z)(xq, xq)`` returned (kind='linear'): SciPy's replacement recipe."""
    from scipy.interpolate import RectBivariateSpline
    xq = np.sort(np.asarray(xq, dtype=np.float64))                  # interp2d.__call__ sorted x and y
    spl = RectBivariateSpline(xs, xs, np.ascontiguousarray(z.T), kx=1, ky=1)
    return spl(xq, xq).T


def crop(field: np.ndarray, X: np.ndarray, npoints_old: int, area_size: float, X_new: np.ndarray) -> np.ndarray:
    """beam2d.py:419-434 (field on the new grid; the caller rebuilds the grids in between)."""
    old_X = X[np.abs(X) < area_size / 2]
    Nc = npoints_old // 2
    n = len(old_X) // 2
    old_X = X[Nc - n:Nc + n]
    win = field[Nc - n:Nc + n, Nc - n:Nc + n]
    return _interp2d_linear(old_X, np.real(win), X_new) + 1j * _interp2d_linear(old_X, np.imag(win), X_new)


def lens_image(field: np.ndarray, X: np.ndarray, Y: np.ndarray, k0: float, f: float, l1: float, l2: float):
    """beam2d.py:525-535.  ``X`` (and the view ``Y``) is scaled in place, as the reference does."""
    scale_factor = - l1 / l2
    xs = X.copy()
    re, im = np.real(field), np.imag(field)
    phase = - (X ** 2 + Y ** 2) * k0 * (l1 + l2) / 2 / l2 ** 2
    X *= scale_factor
    return scale_factor * (_interp2d_linear(xs, re, X) + 1j * _interp2d_linear(xs, im, X)) * (np.cos(phase) + 1j * np.sin(phase))


def modes_synthesize(modes: np.ndarray, coeffs: np.ndarray, npoints: int) -> np.ndarray:
    """beam2d.py:552-562 + :664-666: zero-pad every mode to the grid, sum_i c_i M_i."""
    modes = np.asarray(modes)
    Nb = int(np.sqrt(modes.shape[1]))
    if Nb != npoints:
        dN = (npoints - Nb) // 2
        Nt = npoints - dN
        full = np.zeros((len(modes), npoints, npoints), dtype=np.complex128)
        full[:, dN:Nt, dN:Nt] = modes.reshape((-1, Nb, Nb))
    else:
        full = modes.reshape((-1, Nb, Nb))
    return np.einsum("ijk,i->jk", full, coeffs)


# -------------------------------------------------------------------------- wrapper
class OracleBeam:
    """Minimal stateful wrapper with the reference's attribute names (beam2d.py:79-133)."""

    def __init__(self, area_size, npoints, wl, init_field=None, init_field_gen=None,
                 init_gen_args=(), complex_bits=128, init_spectrum=None):
        self.area_size, self.npoints, self.wl, self.z = area_size, npoints, wl, 0.
        self.complex = {64: np.complex64, 128: np.complex128}[complex_bits]
        self._grids()
        if init_field_gen is not None:
            field = init_field_gen(self.X, self.Y, *init_gen_args)
        elif init_field is not None:
            field = init_field
        else:
            raise ValueError("Init field data is None : 'init_field_gen' must be a "
                             "function or 'init_field' must be an array.")
        if field.dtype != self.complex:
            field = field.astype(self.complex)
        self.field = field
        self.spectrum = fft2(field) if init_spectrum is None else init_spectrum

    def _grids(self):
        g = make_grids(self.area_size, self.npoints, self.wl)
        for k, v in g.items():
            setattr(self, k, v)

    def update(self, field, spectrum=None):                     # beam2d.py:283-288
        self.z = 0
        self.field = field
        self.spectrum = fft2(field) if spectrum is None else spectrum

    def propagate(self, z):
        if z == 0:
            return
        self.z += z
        self.spectrum, self.field = propagate(self.spectrum, self.Kz, z)

    def lens(self, f):
        self.field, self.spectrum = lens(self.field, self.X, self.Y, self.k0, f)

    def coordinate_filter(self, f_init=None, f_gen=None, fargs=()):
        self.field, self.spectrum = coordinate_filter(self.field, self.X, self.Y, f_init, f_gen, fargs)

    def spectral_filter(self, f_init=None, f_gen=None, fargs=()):
        self.spectrum, self.field = spectral_filter(self.spectrum, self.Kx, self.Ky, f_init, f_gen, fargs)

    def expand(self, area_size):
        n_new, pad = expand_shape(self.npoints, self.area_size, area_size)
        self.npoints, self.area_size = n_new, area_size
        self._grids()
        self.field = expand(self.field, pad)
        self.spectrum = fft2(self.field)

    def coarse(self, mean_order=1):
        self.npoints //= mean_order
        self._grids()
        self.field = coarse(self.field, mean_order)
        self.spectrum = fft2(self.field)

    def crop(self, area_size, npoints=0):
        X_old, n_old = self.X, self.npoints
        if npoints:
            self.npoints = npoints
        self.area_size = area_size
        self._grids()
        self.field = crop(self.field, X_old, n_old, area_size, self.X).astype(self.complex)
        self.spectrum = fft2(self.field)

    def lens_image(self, f, l1, l2):
        self.field = lens_image(self.field, self.X, self.Y, self.k0, f, l1, l2).astype(self.complex)
        self.spectrum = fft2(self.field)

    @property
    def iprofile(self):
        return iprofile(self.field)

    @property
    def phiprofile(self):
        return phiprofile(self.field)

    @property
    def centroid(self):
        return centroid(self.field, self.dL)

    @property
    def D4sigma(self):
        return d4sigma(self.field, self.dL)

    @property
    def centroid_intensity(self):
        return centroid_intensity(self.field, self.dL)

    def deconstruct_by_modes(self, modes, mode='scalar'):
        return modes_lstsq(self.field, modes) if mode == 'lstsq' else modes_project(self.field, modes)

    def fast_deconstruct_by_modes(self, modes_matrix_t, modes_matrix_dot_t):
        return modes_fast_lstsq(self.field, modes_matrix_t, modes_matrix_dot_t)

    def construct_by_modes(self, modes, coeffs):
        self.field = modes_synthesize(modes, coeffs, self.npoints)
        self.spectrum = fft2(self.field)
