"""CPU: the oracle restatement against the UNMODIFIED reference imported live (the read-only mount in
the build container, else the pip install under oracle/_ref made by oracle/build_ref.py).  Skipped
when neither is present.  Same numpy build on both sides, so the results must agree to rounding."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import beam_oracle as orc
from oracle import ref_shim
from lightprop2d_b200 import profiles as prof

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not present")

EPS = {64: 1e-6, 128: 1e-15}
CT = {64: np.complex64, 128: np.complex128}


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


@pytest.mark.parametrize("bits", [64, 128])
@pytest.mark.parametrize("n", [64, 130, 256])
def test_chain_matches_reference(ref, bits, n):
    rng = np.random.default_rng(n + bits)
    f0 = (rng.uniform(size=(n, n)) * np.exp(2j * np.pi * rng.uniform(size=(n, n)))).astype(CT[bits])
    area, wl = 4e-1 * n / 1024, 532e-7
    r = ref.Beam2D(area, n, wl, init_field=f0.copy(), complex_bits=bits)
    o = orc.OracleBeam(area, n, wl, init_field=f0.copy(), complex_bits=bits)
    assert np.array_equal(r.X, o.X) and np.array_equal(r.Kx, o.Kx) and np.array_equal(r.Kz, o.Kz)
    for f, z in ((0.1, 0.02), (-0.3, 10.0)):
        r.lens(f); o.lens(f)
        r.propagate(z); o.propagate(z)
        assert rel_l2(o.field, r.field) <= EPS[bits]
        assert rel_l2(o.spectrum, r.spectrum) <= EPS[bits]
        assert rel_l2(o.iprofile, r.iprofile) <= EPS[bits]
    assert o.z == r.z


@pytest.mark.parametrize("bits", [64, 128])
def test_shifted_generator_matches_reference(ref, bits):
    args = (1., 0.03, 0.011, -0.007)
    r = ref.Beam2D(0.3, 128, 632e-7, init_field_gen=ref.gaussian_beam, init_gen_args=args, complex_bits=bits)
    o = orc.OracleBeam(0.3, 128, 632e-7, init_field_gen=prof.gaussian_beam, init_gen_args=args, complex_bits=bits)
    assert np.array_equal(r.X, o.X)
    r.lens(2.); o.lens(2.)
    r.propagate(0.7); o.propagate(0.7)
    assert rel_l2(o.field, r.field) <= EPS[bits]
    assert np.allclose(o.centroid[:2], r.centroid[:2], rtol=1e-12 if bits == 128 else 1e-6)
    assert np.allclose(o.D4sigma, r.D4sigma, rtol=1e-12 if bits == 128 else 1e-6)


def test_generators_match_reference(ref):
    n = 64
    x = np.arange(-n // 2, n // 2) * 1e-3
    y = x.reshape(-1, 1)
    for name, args in (("plane_wave", ()), ("round_hole", (0.01,)), ("rectangle_hole", (0.02, 0.01)),
                       ("square_hole", (0.02,)), ("square_slits", (0.005, 0.02)), ("gaussian_beam", (1., 0.01))):
        a = getattr(ref, name)(x.copy(), y.copy(), *args)
        b = getattr(prof, name)(x.copy(), y.copy(), *args)
        assert np.array_equal(np.asarray(a), np.asarray(b)), name
    for name, args in (("random_wave", ()), ("random_wave_bin", ()), ("random_round_hole", (0.01,)),
                       ("random_round_hole_phase", (0.01,)), ("random_round_hole_bin", (0.01,))):
        np.random.seed(4)
        a = getattr(ref, name)(x.copy(), y.copy(), *args)
        np.random.seed(4)
        b = getattr(prof, name)(x.copy(), y.copy(), *args)
        assert np.array_equal(np.asarray(a), np.asarray(b)), name
