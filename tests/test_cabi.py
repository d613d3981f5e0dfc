"""CPU: the C-ABI library loads, exports every symbol include/lp2d.h declares, and its
host-only helpers agree with numpy bit for bit.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden
from lightprop2d_b200 import _lib
from oracle import beam_oracle as orc


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lp2d.h")).read()
    return sorted(set(re.findall(r"LP2D_API\s+[\w\s\*]+?\b(lp2d_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in lp2d.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_limits():
    lib = _lib.load()
    assert lib.lp2d_version() == 100
    assert lib.lp2d_min_npoints() == 16
    assert lib.lp2d_max_npoints(_lib.C64) >= 4096 and lib.lp2d_max_npoints(_lib.C128) >= 4096
    # BASELINE.json config 4: lines of a 32768^2 grid (cluster line kernel, slab path)
    assert lib.lp2d_max_line_points(_lib.C64) >= 32768 and lib.lp2d_max_line_points(_lib.C128) >= 32768
    assert lib.lp2d_pass_lines_tiles(None, 0, None, 4, None, None, None, 0, 0, 0, 0.0, None) == 1     # NULL plan


@pytest.mark.parametrize("case", range(6))
def test_grid_tables_bit_exact_vs_reference(case):
    """lp2d_grid_tables (C) == Beam2D._construct_grids of the reference (golden X, Kx) and the
    oracle's Kz^2 tables, bit for bit, including an odd and a non power-of-two npoints."""
    g = golden("grids")
    area, n, wl = g[f"c{case}_params"]
    n = int(n)
    lib = _lib.load()
    X, Kx, a, b, x2 = (np.empty(n) for _ in range(5))
    p = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    assert lib.lp2d_grid_tables(n, area, wl, p(X), p(Kx), p(a), p(b), p(x2)) == 0
    assert np.array_equal(X, g[f"c{case}_X"])
    assert np.array_equal(Kx, g[f"c{case}_Kx"])
    oa, ob = orc.kz2_tables(area, n, wl)
    assert np.array_equal(b, ob)
    # k0**2 in Python is pow(k0, 2); the C helper uses k0*k0: equal for these inputs
    assert np.array_equal(a, oa)
    assert np.array_equal(x2, g[f"c{case}_X"] ** 2)
    rows = g[f"c{case}_Kz_rows"]
    with np.errstate(invalid="ignore"):
        kz = np.nan_to_num(np.sqrt(a[None, :] - b[rows][:, None]))
    assert np.array_equal(kz, g[f"c{case}_Kz"])


def test_error_paths_without_gpu():
    lib = _lib.load()
    assert lib.lp2d_grid_tables(0, 1.0, 1.0, None, None, None, None, None) == 1
    assert b"bad n" in lib.lp2d_last_error()
    h = C.c_void_p()
    t = np.zeros(48)
    p = t.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.lp2d_plan_create(C.byref(h), 48, _lib.C64, p, p, p) == 2      # not a power of two
    assert b"not supported" in lib.lp2d_last_error()
    assert lib.lp2d_plan_create(C.byref(h), 64, 7, p, p, p) == 1
    assert lib.lp2d_plan_create(C.byref(h), 64, _lib.C64, None, p, p) == 1
    assert lib.lp2d_fft2(None, None, None, 1, 0, None) == 1
    assert lib.lp2d_expand(_lib.C64, None, None, 4, 8, 1, None) == 1
    assert lib.lp2d_plan_destroy(None) == 0
    with pytest.raises(NotImplementedError):
        _lib.check(2, "x")
    with pytest.raises(ValueError):
        _lib.check(1, "x")
    with pytest.raises(_lib.Lp2dError):
        _lib.check(3, "x")
