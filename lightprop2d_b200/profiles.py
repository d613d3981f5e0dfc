"""Initial-field and mask generators with the reference's call contract.

Every generator is called as ``f(X, Y, *args)`` where ``X`` is the 1-D coordinate
grid and ``Y`` the same grid as a column (``Beam2D.X`` / ``Beam2D.Y``), and returns
an ``(N, N)`` bool / real / complex array by broadcasting.  They mirror
``lightprop2d/profiles.py:35-136`` (same names, argument order, RNG call sequence,
so a seeded ``numpy.random`` gives the same field) but are host-side input
producers: they run once per beam and are not part of the CUDA hot path
(SURVEY.md §2 row 11).  Quirks kept on purpose, because callers see them:

* ``round_hole(R)`` thresholds a unit Gaussian at ``exp(-1/2)``, so the true
  aperture radius is ``R/sqrt(2)`` (profiles.py:79-80);
* ``gaussian_beam`` shifts its ``x``/``y`` arguments IN PLACE (profiles.py:69-70).
"""
import numpy as np

__all__ = ('plane_wave', 'random_wave', 'random_wave_bin', 'gaussian_beam',
           'round_hole', 'random_round_hole', 'random_round_hole_bin',
           'random_round_hole_phase', 'rectangle_hole', 'square_hole',
           'square_slits')


def binning(field, binning_order):
    """Upsample by repeating every sample ``binning_order`` times along both axes
    (Kronecker product with a block of ones, profiles.py:29-32)."""
    return np.kron(field, np.ones((binning_order, binning_order)))


def _coarse_shape(x, y, binning_order):
    return (int(len(y) // binning_order), int(len(x) // binning_order))


def plane_wave(x, y):
    return np.ones((len(y), len(x)))


def random_wave_bin(x, y, binning_order=1):
    """Binary {0,1} speckle (profiles.py:41-50)."""
    return binning(np.random.randint(0, 2, size=_coarse_shape(x, y, binning_order)), binning_order)


def random_wave(x, y, binning_order=1):
    """Uniform [0,1) speckle (profiles.py:53-61)."""
    return binning(np.random.uniform(size=_coarse_shape(x, y, binning_order)), binning_order)


def gaussian_beam(x, y, A0, rho0, x0=0, y0=0):
    """``A0 * exp(-((x-x0)^2 + (y-y0)^2) / rho0^2)`` (profiles.py:64-71); shifts x, y in place."""
    x -= x0
    y -= y0
    return A0 * np.exp(- (x ** 2 + y ** 2) / rho0 ** 2)


def round_hole(x, y, R, x0=0, y0=0):
    """Boolean disc of radius R/sqrt(2) (profiles.py:74-81)."""
    return gaussian_beam(x - x0, y - y0, 1, R) >= 1 / np.exp(0.5)


def _apertured(field, x, y, R, x0, y0):
    field[round_hole(x, y, R, x0, y0) == 0] = 0
    return field


def random_round_hole_bin(x, y, R, x0=0, y0=0, binning_order=1):
    return _apertured(random_wave_bin(x, y, binning_order), x, y, R, x0, y0)


def random_round_hole(x, y, R, x0=0, y0=0, binning_order=1):
    return _apertured(random_wave(x, y, binning_order), x, y, R, x0, y0)


def random_round_hole_phase(x, y, R, x0=0, y0=0, binning_order=1):
    """Unit-amplitude random phase inside the disc (profiles.py:105-113)."""
    angle = random_round_hole(x, y, R, x0, y0, binning_order) * 2 * np.pi
    return round_hole(x, y, R, x0, y0) * (np.cos(angle) + 1j * np.sin(angle))


def rectangle_hole(x, y, dx, dy, x0=0, y0=0):
    return (np.abs(x - x0) < (dx / 2)) & (np.abs(y - y0) < (dy / 2))


def square_hole(x, y, d, x0=0, y0=0):
    return rectangle_hole(x, y, d, d, x0, y0)


def square_slits(x, y, d, slits_distance, x0=0, y0=0):
    dl = slits_distance / 2
    return square_hole(x, y, d, x0 - dl, y0) | square_hole(x, y, d, x0 + dl, y0)
