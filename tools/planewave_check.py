"""Size-independent parity check for grids too large for the numpy oracle (BASELINE.json config 4).

A sum of M plane waves  f[j, i] = sum_m c_m exp(2 pi i (kx_m i + ky_m j) / N)  has the spectrum
N^2 c_m at the grid frequencies (ky_m, kx_m) and nothing elsewhere, so after ``propagate(z)`` the
reference (beam2d.py:485-490) leaves  sum_m c_m H[ky_m, kx_m] exp(...)  with
H = exp(2 pi i (trunc(Kz z) - Kz z)),  Kz = nan_to_num(sqrt(k0^2 - Kx^2 - Ky^2))  (:307-308).  The
expectation costs M scalars in FP64 from the oracle's own tables; it exercises every index map
of the decomposed transform (row pass, exchange, transposed pass, H orientation) at full size.
Both sides are built slab by slab as (rows x M) @ (M x N) products, so the check also runs on a
row slab of a multi-GPU grid.
"""
import numpy as np
import torch


def waves(n, seed=5, m=6):
    rng = np.random.default_rng(seed)
    kx = rng.integers(0, n, m)
    ky = rng.integers(0, n, m)
    # one wave beyond the evanescent limit when the grid reaches it is fine: H = 1 there (nan_to_num)
    c = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    return kx, ky, c


def slab_sum(n, r0, r1, kx, ky, coef, device, dtype):
    """rows r0..r1-1 of sum_m coef_m exp(2 pi i (kx_m i + ky_m j) / n), phases formed in FP64 from exact integers."""
    i = torch.arange(n, device=device, dtype=torch.int64)
    j = torch.arange(r0, r1, device=device, dtype=torch.int64)
    kxt = torch.as_tensor(kx, device=device, dtype=torch.int64)
    kyt = torch.as_tensor(ky, device=device, dtype=torch.int64)
    ex = torch.polar(torch.ones((), device=device, dtype=torch.float64),
                     (2 * np.pi / n) * ((kxt[:, None] * i[None, :]) % n).to(torch.float64))       # (M, n)
    ey = torch.polar(torch.ones((), device=device, dtype=torch.float64),
                     (2 * np.pi / n) * ((j[:, None] * kyt[None, :]) % n).to(torch.float64))       # (rows, M)
    ey = ey * torch.as_tensor(coef, device=device, dtype=torch.complex128)[None, :]
    out = torch.empty(r1 - r0, n, device=device, dtype=dtype)
    step = max(1, (1 << 24) // n)
    for a in range(0, r1 - r0, step):            # FP64 product in row blocks, rounded once to the field type
        out[a:a + step] = (ey[a:a + step] @ ex).to(dtype)
    return out


def transfer(area_size, n, wl, kx, ky, z):
    """H[ky_m, kx_m] with the reference's operation order (oracle.beam_oracle.kz2_tables / transfer_phase)."""
    from oracle import beam_oracle as orc
    a, b = orc.kz2_tables(area_size, n, wl)
    with np.errstate(invalid="ignore"):
        kz = np.nan_to_num(np.sqrt(a[kx] - b[ky]))
    ph = orc.transfer_phase(kz, z)
    return np.cos(ph) + 1j * np.sin(ph)


def rel_l2_slab(x, ref, group=None):
    """relative L2 over the whole grid (sums reduced over the ranks of `group` when distributed)."""
    import torch.distributed as dist
    num = torch.zeros(2, device=x.device, dtype=torch.float64)
    step = max(1, (1 << 24) // x.shape[1])
    for a in range(0, x.shape[0], step):
        d = (x[a:a + step].to(torch.complex128) - ref[a:a + step].to(torch.complex128))
        num[0] += (d.real ** 2 + d.imag ** 2).sum()
        r = ref[a:a + step].to(torch.complex128)
        num[1] += (r.real ** 2 + r.imag ** 2).sum()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(num, group=group)
    return float(torch.sqrt(num[0] / num[1]).item())


def check(n, bits, area_size, wl, z, rank=0, world=1, group=None, m=6):
    """Build the plane-wave slab, run SlabBeam2D.propagate(z) and propagate(-z); returns the relative L2
    errors of the spectrum support, the propagated field and the round trip."""
    from lightprop2d_b200.slab import SlabBeam2D, slab_rows
    dev = torch.device("cuda", torch.cuda.current_device())
    dt = torch.complex64 if bits == 64 else torch.complex128
    kx, ky, c = waves(n, m=m)
    r0, r1 = slab_rows(n, rank, world)
    f0 = slab_sum(n, r0, r1, kx, ky, c, dev, dt)
    b = SlabBeam2D(area_size, n, wl, init_field=f0.clone(), complex_bits=bits, group=group)   # the beam owns (and overwrites) its buffer
    # spectrum (transposed slab: spec_t[i - r0, j] = spectrum[j, i]): N^2 c_m at (ky_m, kx_m), zero elsewhere
    want = torch.zeros_like(b.spectrum_t)
    for k, (x, y) in enumerate(zip(kx, ky)):
        if r0 <= x < r1:
            want[x - r0, y] += complex(c[k]) * n * n
    e_spec = rel_l2_slab(b.spectrum_t, want, group)
    del want
    h = transfer(area_size, n, wl, kx, ky, z)
    b.propagate(z)
    ref = slab_sum(n, r0, r1, kx, ky, c * h, dev, dt)
    e_field = rel_l2_slab(b.field, ref, group)
    del ref
    b.propagate(-z)
    e_round = rel_l2_slab(b.field, f0, group)
    return {"spectrum": e_spec, "field": e_field, "round_trip": e_round}
