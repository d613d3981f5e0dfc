"""CPU, world_size 2, gloo: the N>1 host logic of the ensemble path — block sharding and the
optional gather of per-beam results.  The data path itself has no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightprop2d_b200.ensemble import gather_to_rank0, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        # stand-in for "propagate my block": a per-beam map that depends only on the global beam index
        beams = torch.arange(lo, hi, dtype=torch.float64)
        local = beams[:, None, None] * torch.ones(1, 4, 4, dtype=torch.float64) + 0.5
        full = gather_to_rank0(local, n_total)
        # per-beam scalars too
        sc = gather_to_rank0(beams * 2, n_total)
        if rank == 0:
            q.put((full.numpy(), sc.numpy()))
        else:
            assert full is None and sc is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    world, n_total = 2, 7        # ragged: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, sc = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert full.shape == (n_total, 4, 4)
    assert np.array_equal(full[:, 0, 0], np.arange(n_total) + 0.5)
    assert np.array_equal(sc, np.arange(n_total) * 2.0)


def test_single_process_gather_is_identity():
    t = torch.arange(6.).reshape(3, 2)
    assert gather_to_rank0(t, 3) is t


# --------------------------------------------------------------------------- slab-decomposed grid
class _NumpyLineBackend:
    """TEST stand-in for lightprop2d_b200.slab.CudaLineBackend: the same interface, line transforms and
    the transfer function from the oracle's expressions, CPU torch tensors (so that gloo can move them).
    Exercises the partitioning / exchange / orientation logic of SlabBeam2D without a GPU."""

    def __init__(self, area_size, npoints, wl, complex_bits, reads_tiles=False):
        from oracle import beam_oracle as orc
        # True: lines() is handed the (G, rows, tile) receive buffer of the exchange, as the CUDA backend is
        # (lp2d_pass_lines_tiles); False: SlabBeam2D unpacks it with a strided copy first
        self.reads_tiles = reads_tiles
        self.n = npoints
        self.g = orc.make_grids(area_size, npoints, wl)
        self.npc = np.complex64 if complex_bits == 64 else np.complex128
        self.cdtype = torch.complex64 if complex_bits == 64 else torch.complex128
        self.rdtype = torch.float32 if complex_bits == 64 else torch.float64

    def empty(self, *shape, real=False):
        return torch.empty(*shape, dtype=self.rdtype if real else self.cdtype)

    def lines(self, x, line0, inverse=False, transposed=False, z=None, state_out=None, intensity=None, out=None):
        if x.dim() == 3:                            # tiles [c][i][k] = slab[i][c*tile + k] (include/lp2d.h, lp2d_pass_lines_tiles)
            assert self.reads_tiles
            x = x.permute(1, 0, 2).reshape(x.shape[1], -1)
        a = x.numpy().copy()
        rows = a.shape[0]
        if z is not None:
            Kz = self.g["Kz"]                       # [j, i]
            # slab orientation: natural = rows j of the plane; transposed = line i, position j
            kz = Kz[:, line0:line0 + rows].T if transposed else Kz[line0:line0 + rows, :]
            d = kz * z
            ph = 2 * np.pi * (np.trunc(d) - d)
            a = (a * (np.cos(ph) + 1j * np.sin(ph))).astype(self.npc)
            if state_out is not None:
                state_out.copy_(torch.from_numpy(a))
        r = (np.fft.ifft(a, axis=-1) if inverse else np.fft.fft(a, axis=-1)).astype(self.npc)
        res = torch.from_numpy(r)
        if intensity is not None:
            intensity.copy_(torch.from_numpy((np.abs(r) ** 2).astype(intensity.numpy().dtype)))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def pack_transpose(self, x, tile):
        rows, ld = x.shape
        return x.reshape(rows, ld // tile, tile).permute(1, 2, 0).contiguous()

    def lens(self, x, line0, k0, f):
        X = self.g["X"]
        ph = (X[None, :] ** 2 + X[line0:line0 + x.shape[0], None] ** 2) * k0 / (2 * f)      # beam2d.py:508
        x.copy_(torch.from_numpy((x.numpy() * (np.cos(ph) + 1j * np.sin(ph))).astype(self.npc)))

    def mul_mask(self, x, mask):
        m = mask.numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
        x.copy_(torch.from_numpy((x.numpy() * m).astype(self.npc)))

    def intensity(self, x):
        return torch.from_numpy(np.abs(x.numpy()) ** 2)

    def project(self, flat_field, modes_cols):
        return torch.from_numpy((modes_cols.numpy() @ flat_field.numpy()).astype(self.npc))


def _slab_worker(rank, world, port, n, bits, q, reads_tiles=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightprop2d_b200.slab import SlabBeam2D, slab_rows
        area, wl, z1, z2 = 200e-4 * n / 256, 632e-7, 100e-4, -37e-4
        rng = np.random.default_rng(5)
        full = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64 if bits == 64 else np.complex128)
        r0, r1 = slab_rows(n, rank, world)
        b = SlabBeam2D(area, n, wl, init_field=full[r0:r1].copy(), complex_bits=bits,
                       backend=_NumpyLineBackend(area, n, wl, bits, reads_tiles))
        spec0 = b.gather_spectrum().numpy()
        inten = b.propagate(z1, intensity=True)
        b.propagate(z2)
        fld = b.gather_field().numpy()
        spec = b.gather_spectrum().numpy()
        # lens, filters and the K-split mode projection (partial sums + one all-reduce) on the decomposed grid
        from lightprop2d_b200 import profiles as prof
        b.lens(0.01)
        mask = np.random.default_rng(6).standard_normal((n, n)) + 0.5
        b.coordinate_filter(f_gen=prof.round_hole, fargs=(area / 3,), f_init=mask[r0:r1])
        b.spectral_filter(f_gen=prof.rectangle_hole, fargs=(4000., 2500.), f_init=mask[:, r0:r1])
        fld2 = b.gather_field().numpy()
        nb = n // 2
        modes = np.random.default_rng(7).standard_normal((3, nb * nb)) + 1j * np.random.default_rng(8).standard_normal((3, nb * nb))
        coeffs = b.deconstruct_by_modes(modes.astype(full.dtype)).numpy()
        if rank == 0:
            q.put((full, spec0, fld, spec, inten.numpy(), b.z, fld2, mask, modes, coeffs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("reads_tiles", [False, True])
def test_two_rank_slab_matches_oracle(reads_tiles):
    """Both forms of the exchange: unpacked slab, and the receive buffer handed to the line pass as tiles."""
    from oracle import beam_oracle as orc
    world, n, bits = 2, 64, 128
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, n, bits, q, reads_tiles)) for r in range(world)]
    for p in procs:
        p.start()
    full, spec0, fld, spec, inten0, z, fld2, mask, modes, coeffs = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    area, wl, z1, z2 = 200e-4 * n / 256, 632e-7, 100e-4, -37e-4
    o = orc.OracleBeam(area, n, wl, init_field=full.copy(), complex_bits=bits)
    assert np.linalg.norm(spec0 - o.spectrum) <= 1e-12 * np.linalg.norm(o.spectrum)
    o.propagate(z1)
    i1 = o.iprofile[: n // world]
    assert np.linalg.norm(inten0 - i1) <= 1e-12 * np.linalg.norm(i1)
    o.propagate(z2)
    assert abs(z - o.z) < 1e-15
    assert np.linalg.norm(fld - o.field) <= 1e-12 * np.linalg.norm(o.field)
    assert np.linalg.norm(spec - o.spectrum) <= 1e-12 * np.linalg.norm(o.spectrum)
    from lightprop2d_b200 import profiles as prof
    o.lens(0.01)
    o.coordinate_filter(f_gen=prof.round_hole, fargs=(area / 3,), f_init=mask)
    o.spectral_filter(f_gen=prof.rectangle_hole, fargs=(4000., 2500.), f_init=mask)
    assert np.linalg.norm(fld2 - o.field) <= 1e-12 * np.linalg.norm(o.field)
    want = o.deconstruct_by_modes(modes.astype(full.dtype))
    assert np.linalg.norm(coeffs - want) <= 1e-11 * np.linalg.norm(want)


def test_slab_rows_validation():
    import pytest
    from lightprop2d_b200.slab import slab_rows
    assert slab_rows(1024, 3, 8) == (384, 512)
    with pytest.raises(ValueError):
        slab_rows(1000, 0, 3)
    with pytest.raises(ValueError):
        slab_rows(1024, 8, 8)
