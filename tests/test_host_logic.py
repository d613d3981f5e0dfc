"""CPU: host-side logic — profile generators, sharding, and that the product path fails
loudly (no CPU fallback) when there is no CUDA device."""
import numpy as np
import pytest
import torch

from conftest import golden
from lightprop2d_b200 import profiles as prof
from lightprop2d_b200.ensemble import shard_range, shard_sizes
import lightprop2d_b200 as lp


def test_public_names():
    for name in ("Beam2D", "nm", "um", "mm", "cm", "m", *prof.__all__):
        assert hasattr(lp, name)
    assert (lp.nm, lp.um, lp.mm, lp.cm, lp.m) == (1e-7, 1e-4, 1e-1, 1, 100)


def test_profile_identities():
    """The four asserts of the reference's tests/test_profiles.py:31-38."""
    npoints, dL = 256, 1e-4
    x = np.arange(-npoints // 2, npoints // 2, 1) * dL
    y = x.reshape((-1, 1))
    R = d = rho0 = dL * npoints / 4
    dx = dy = dL * npoints / 4
    assert prof.gaussian_beam(x, y, 1., rho0)[npoints // 2, npoints // 2] == 1.
    assert np.sum(prof.plane_wave(x, y)) == npoints ** 2
    prof.random_wave(x, y); prof.round_hole(x, y, R); prof.random_round_hole(x, y, R)
    assert np.allclose(prof.square_hole(x, y, d), prof.rectangle_hole(x, y, dx, dy))
    assert np.allclose(prof.square_slits(x, y, d, d),
                       prof.square_hole(x, y, d, x0=-d / 2) + prof.square_hole(x, y, d, x0=d / 2))


def test_round_hole_true_radius_and_seeded_rng():
    x = np.arange(-128, 128) * 1e-4
    y = x.reshape((-1, 1))
    hole = prof.round_hole(x, y, 50e-4)
    r_eff = np.sqrt(hole.sum() / np.pi) * 1e-4
    assert abs(r_eff / 50e-4 - 1 / np.sqrt(2)) < 0.01          # radius R/sqrt(2)
    np.random.seed(0)
    a = prof.random_round_hole(x, y, 25e-4)
    np.random.seed(0)
    ref = np.random.uniform(size=(256, 256))
    ref[~prof.round_hole(x, y, 25e-4)] = 0
    assert np.array_equal(a, ref)
    np.random.seed(1)
    ph = prof.random_round_hole_phase(x, y, 25e-4)
    assert np.iscomplexobj(ph) and np.allclose(np.abs(ph[hole & (np.abs(ph) > 0)]), 1)
    b = prof.random_wave_bin(x, y, 4)
    assert b.shape == (256, 256) and set(np.unique(b)) <= {0., 1.}


def test_gaussian_beam_shifts_grid_in_place():
    x = np.arange(-4, 4) * 1.0
    y = x.reshape((-1, 1))            # a view, as Beam2D.Y is of Beam2D.X
    prof.gaussian_beam(x, y, 1., 2., 1., 0.5)
    assert np.array_equal(x, np.arange(-4, 4) - 1.5)


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 8), (10, 4), (3, 8), (0, 2), (7, 7)])
def test_shard_range_tiles_in_order(n, world):
    blocks = [shard_range(n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a1 >= a0
    sizes = shard_sizes(n, world)
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, world, world)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lp.Beam2D(200 * lp.um, 256, 632 * lp.nm, init_field_gen=lp.plane_wave)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        from lightprop2d_b200.ensemble import EnsemblePropagator
        EnsemblePropagator(200 * lp.um, 256, 632 * lp.nm)


def test_product_never_imports_oracle():
    import os
    import re
    root = os.path.dirname(os.path.abspath(lp.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "numpy.fft.fft2" not in text and "np.fft.fft2" not in text and "np.fft.ifft2" not in text, f


def test_device_profiles_mirror_the_host_generators_and_need_a_gpu():
    """device_profiles has the reference's generator names with the same positional arguments (profiles.py:35-136) and,
    like the rest of the product path, no CPU fallback."""
    import inspect
    from lightprop2d_b200 import device_profiles as dp
    assert tuple(dp.__all__) == tuple(prof.__all__)
    for name in prof.__all__:
        host = [p.name for p in inspect.signature(getattr(prof, name)).parameters.values()]
        dev = [p.name for p in inspect.signature(getattr(dp, name)).parameters.values() if p.kind != p.VAR_KEYWORD]
        assert dev == host, (name, dev, host)
    if not torch.cuda.is_available():
        x = np.arange(-8, 8) * 1e-4
        with pytest.raises(RuntimeError, match="CUDA"):
            dp.round_hole(x, x.reshape(-1, 1), 5e-4)
