"""Ensembles of independent beams: the fused fresh chain and its sharding across GPUs.

The reference has one field per ``Beam2D`` and no cross-beam coupling anywhere
(SURVEY.md §8e), so a speckle / ghost-imaging ensemble or a z-sweep over many
beams is embarrassingly parallel: beams are split in contiguous blocks over the
ranks (one process per GPU), grids and twiddles are replicated, and there is NO
data-path collective.  The only optional communication is a gather of per-beam
results (scalars or intensity maps) to rank 0.

Per beam the work is what the reference does in
``Beam2D(init_field=f); beam.propagate(z); beam.iprofile`` (beam2d.py:118-133,
481-490, 717): fft2 -> H(z) -> ifft2 -> |E|^2, here as three fused passes that never
materialise the spectrum.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .beam2d import _get_plan, _torch


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block [start, stop) of ``n_items`` owned by ``rank`` of ``world``.

    The first ``n_items % world`` ranks get one extra item; the blocks tile
    [0, n_items) in rank order, so concatenating per-rank results in rank order
    restores the original beam order.
    """
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_items < 0:
        raise ValueError("n_items must be >= 0")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_to_rank0(local, n_total: int, group=None):
    """Gather per-beam results (first axis = this rank's beams) on rank 0, in beam order.

    ``local`` is a torch tensor (CPU for gloo, CUDA for nccl).  Returns the full
    tensor on rank 0 and ``None`` elsewhere.  With an uninitialised process group
    (single process) it returns ``local``.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} beams, expected {sizes[rank]}")
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


class EnsemblePropagator:
    """fft2 -> H(z) -> ifft2 -> |E|^2 over a batch of independent beams on one GPU.

    ``area_size``, ``npoints``, ``wl`` and ``complex_bits`` have the meaning of the
    ``Beam2D`` constructor arguments; the grids are built with the same expressions.
    """

    def __init__(self, area_size, npoints, wl, complex_bits=64, device=None, work_mb=32):
        torch = _torch()
        self.area_size, self.npoints, self.wl, self.complex_bits = area_size, int(npoints), wl, complex_bits
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._tdev = torch.device("cuda", self.device)
        self._cdtype = torch.complex64 if complex_bits == 64 else torch.complex128
        self._rdtype = torch.float32 if complex_bits == 64 else torch.float64
        dL = area_size / npoints
        X = np.arange(-npoints // 2, npoints // 2, 1) * dL
        k0 = 1 / wl
        Kx = np.fft.fftfreq(npoints, d=dL)
        self._plan = _get_plan(self.device, self.npoints, complex_bits, area_size, wl,
                               k0 ** 2 - Kx ** 2, Kx ** 2, X ** 2)
        self._lib = _lib.load()
        plane = self.npoints * self.npoints * (8 if complex_bits == 64 else 16)
        self._work = torch.empty(max(plane, work_mb << 20), dtype=torch.uint8, device=self._tdev)

    def _stream(self):
        torch = _torch()
        return C.c_void_p(torch.cuda.current_stream(self._tdev).cuda_stream)

    def random_fields(self, batch, R, x0=0., y0=0., seed=0, first_beam=0, phase=False, out=None):
        """``batch`` input fields generated on the device: ``profiles.random_round_hole`` (``phase=False``:
        amplitude uniform in [0, 1) inside the aperture) or ``profiles.random_round_hole_phase``
        (``phase=True``: exp(2 pi i u) inside), zero outside (profiles.py:94-113).  Aperture and numbers are evaluated on the device
        (``device_profiles``, ``lp2d_generate``: the reference's FP64 ``round_hole`` expression per sample); the
        uniform numbers come from a counter-based generator keyed by (seed, first_beam + b, sample), so
        beam ``first_beam + b`` is the same whatever batch or rank generates it (shard with
        ``shard_range`` and pass the block start as ``first_beam``)."""
        from . import device_profiles as dp
        dL = self.area_size / self.npoints
        X = np.arange(-self.npoints // 2, self.npoints // 2, 1) * dL
        gen = dp.random_round_hole_phase if phase else dp.random_round_hole
        return gen(X, X.reshape((-1, 1)), R, x0, y0, batch=int(batch), complex_bits=self.complex_bits, device=self.device,
                   seed=int(seed), first_beam=int(first_beam), out=out)

    def intensity(self, fields, z, out=None, work=None):
        """|ifft2(fft2(fields) * H(z))|^2 for device ``fields`` of shape (B, N, N); returns a device tensor.
        ``work``: scratch for the intermediates (default: the propagator's own buffer; callers that run
        several calls concurrently on different streams pass one buffer per stream)."""
        torch = _torch()
        if fields.dtype != self._cdtype or not fields.is_cuda:
            raise ValueError(f"fields must be a CUDA tensor of dtype {self._cdtype}")
        fields = fields.contiguous()
        n = self.npoints
        if fields.shape[-2:] != (n, n):
            raise ValueError(f"fields must end in ({n}, {n})")
        batch = fields.numel() // (n * n)
        if out is None:
            out = torch.empty(fields.shape, dtype=self._rdtype, device=self._tdev)
        with torch.cuda.device(self._tdev):
            _lib.check(self._lib.lp2d_propagate_field(self._plan.handle, fields.data_ptr(), None, None,
                                                      out.data_ptr(), (self._work if work is None else work).data_ptr(),
                                                      (self._work if work is None else work).numel(),
                                                      batch, float(z), self._stream()), "propagate_field")
        return out

    def field(self, fields, z, out=None):
        """ifft2(fft2(fields) * H(z)) for device ``fields`` (B, N, N)."""
        torch = _torch()
        fields = fields.contiguous()
        n = self.npoints
        batch = fields.numel() // (n * n)
        if out is None:
            out = torch.empty_like(fields)
        with torch.cuda.device(self._tdev):
            _lib.check(self._lib.lp2d_propagate_field(self._plan.handle, fields.data_ptr(), out.data_ptr(), None,
                                                      None, None, 0, batch, float(z), self._stream()),
                       "propagate_field")
        return out

    def intensity_from_host(self, host_fields, z, chunk=16, out=None, sync=True):
        """End-to-end form: ``host_fields`` is a pinned (or pageable) CPU tensor / numpy array of
        shape (B, N, N); every chunk is copied H2D, propagated and its intensity copied D2H, with the
        copies of neighbouring chunks overlapped on three streams.  Returns a pinned CPU tensor that is
        complete when the call returns (``sync=False``: the caller synchronises the current stream before
        reading it, e.g. to queue several ensembles back to back)."""
        torch = _torch()
        if isinstance(host_fields, np.ndarray):
            host_fields = torch.from_numpy(host_fields)
        B = host_fields.shape[0]
        if out is None:
            out = torch.empty(host_fields.shape, dtype=self._rdtype, pin_memory=True)
        # streams and staging buffers are kept between calls (per chunk shape)
        key = (chunk,) + tuple(host_fields.shape[1:])
        cache = getattr(self, "_e2e_cache", None)
        if cache is None or cache[0] != key:
            nbuf = 3
            cache = (key,
                     [torch.cuda.Stream(self._tdev) for _ in range(nbuf)],
                     [torch.empty((chunk,) + tuple(host_fields.shape[1:]), dtype=self._cdtype, device=self._tdev) for _ in range(nbuf)],
                     [torch.empty((chunk,) + tuple(host_fields.shape[1:]), dtype=self._rdtype, device=self._tdev) for _ in range(nbuf)],
                     [torch.empty_like(self._work) for _ in range(nbuf)])      # one scratch buffer per stream
            self._e2e_cache = cache
        _, streams, dev_in, dev_out, works = cache
        nbuf = len(streams)
        cur = torch.cuda.current_stream(self._tdev)
        for s in streams:
            s.wait_stream(cur)
        for k, b0 in enumerate(range(0, B, chunk)):
            nb = min(chunk, B - b0)
            s = streams[k % nbuf]
            with torch.cuda.stream(s):
                dev_in[k % nbuf][:nb].copy_(host_fields[b0:b0 + nb], non_blocking=True)
                self.intensity(dev_in[k % nbuf][:nb], z, out=dev_out[k % nbuf][:nb], work=works[k % nbuf])
                out[b0:b0 + nb].copy_(dev_out[k % nbuf][:nb], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
        if sync:
            # the D2H copies are asynchronous (pinned destination): the host may read `out` only once they landed
            for s in streams:
                s.synchronize()
        return out


def propagate_ensemble(fields, area_size, wl, z, complex_bits=64, rank=None, world=None, gather=False,
                       group=None):
    """Shard ``fields`` (host array/tensor (B, N, N), identical on every rank) across the ranks,
    propagate this rank's block and return its intensities (CPU tensor), or with ``gather=True`` the
    full (B, N, N) intensity stack on rank 0 (None elsewhere).  No data-path collective is used."""
    torch = _torch()
    import torch.distributed as dist
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    B, n = fields.shape[0], fields.shape[-1]
    lo, hi = shard_range(B, rank, world)
    prop = EnsemblePropagator(area_size, n, wl, complex_bits=complex_bits)
    local = fields[lo:hi]
    if isinstance(local, np.ndarray):
        local = torch.from_numpy(np.ascontiguousarray(local))
    res = prop.intensity_from_host(local.to(prop._cdtype), z)
    torch.cuda.synchronize()
    if gather:
        return gather_to_rank0(res if dist.get_backend(group) == "gloo" else res.to(prop._tdev), B, group)
    return res
