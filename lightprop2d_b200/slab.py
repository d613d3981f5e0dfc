"""A single grid decomposed in slabs over several GPUs (BASELINE.json config 4).

The reference has nothing like this -- a grid that does not fit its device simply runs out of
memory (tests/gaussbeam_performance.py:33-34).  Here rank r of G owns

* the ROW slab of the field: rows ``r*N/G .. (r+1)*N/G - 1`` of the N x N grid, row-major
  ``(N/G, N)``;
* the COLUMN slab of the spectrum, stored TRANSPOSED: ``spec_t[i, j] = spectrum[j, r*N/G + i]``,
  also ``(N/G, N)``, so that a transform along grid columns is again a transform along the
  contiguous axis (a row pass of the CUDA library).

``fft2``: local row pass -> all-to-all transpose -> local row pass on the transposed slab.
``propagate(z)`` (beam2d.py:481-490) with the spectrum resident in the transposed layout needs ONE
exchange: ``spec_t *= H`` and the inverse transform along grid columns are one fused local pass
(H in the transposed orientation comes from a second plan whose two FP64 tables are swapped and
negated: ``(-b[pos]) - (-a[line])`` is bit for bit ``a[line] - b[pos]``), then the all-to-all
transpose back, then the local inverse row pass (optionally with the fused ``|E|^2``).

The exchange is the only collective: ``torch.distributed.all_to_all_single`` (NCCL over
NVLink / NVSwitch on GPUs) of G tiles of ``(N/G)^2`` samples per rank; tiles are transposed while
they are packed (``lp2d_pack_transpose``), the unpack is one strided copy.

The line passes come from a small backend object so that the partitioning / exchange logic can be
exercised on CPU with the gloo backend (tests/test_dist_gloo.py substitutes a numpy backend; the
product backend below is CUDA only and raises without a device).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def slab_rows(npoints: int, rank: int, world: int):
    """[start, stop) of the grid rows (or columns) owned by ``rank``; ``world`` must divide ``npoints``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if npoints % world:
        raise ValueError(f"npoints={npoints} is not divisible by the number of slabs {world}")
    nb = npoints // world
    return rank * nb, (rank + 1) * nb


def kz2_tables(area_size: float, npoints: int, wl: float):
    """``a[i] = k0^2 - Kx[i]^2``, ``b[j] = Ky[j]^2``, ``x2[i] = X[i]^2`` with the reference's expressions
    (beam2d.py:299-308) -- the same tables ``Beam2D`` uploads."""
    dL = area_size / npoints
    X = np.arange(-npoints // 2, npoints // 2, 1) * dL
    k0 = 1 / wl
    Kx = np.fft.fftfreq(npoints, d=dL)
    return k0 ** 2 - Kx ** 2, Kx ** 2, X ** 2


class CudaLineBackend:
    """Row passes and the transposing pack of liblp2d.so on the current CUDA device."""
    reads_tiles = True      # lines() accepts the (G, rows, tile) receive buffer of the exchange

    def __init__(self, area_size, npoints, wl, complex_bits, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("the slab-decomposed grid needs a CUDA device (there is no CPU fallback)")
        from .beam2d import _Plan
        self.torch = torch
        self.lib = _lib.load()
        self.n, self.bits = int(npoints), complex_bits
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.cdtype = torch.complex64 if complex_bits == 64 else torch.complex128
        self.rdtype = torch.float32 if complex_bits == 64 else torch.float64
        a, b, x2 = kz2_tables(area_size, npoints, wl)
        self._host_tabs, self._tabs = (a, b, x2), None
        with torch.cuda.device(self.device):
            self.plan = _Plan(npoints, complex_bits, a, b, x2)
            # transposed orientation: line = grid column i, position = grid row j
            self.plan_t = _Plan(npoints, complex_bits, -b, -a, x2)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def empty(self, *shape, real=False):
        return self.torch.empty(*shape, dtype=self.rdtype if real else self.cdtype, device=f"cuda:{self.device}")

    def lines(self, x, line0, inverse=False, transposed=False, z=None, state_out=None, intensity=None, out=None):
        """(i)FFT along the last axis of the slab ``x`` (lines line0 .. line0 + rows - 1 of the plane).
        With ``z``: x *= H(z) first (H in the slab's orientation), the product goes to ``state_out``.
        ``x`` is either the ``(rows, n)`` slab or the ``(n // tile, rows, tile)`` receive buffer of the tile
        exchange (``exchange_transpose(..., tiles=True)``), which the kernel reads in place of an unpacked copy."""
        plan = self.plan_t if transposed else self.plan
        tiled = x.dim() == 3
        rows = int(x.shape[1] if tiled else x.shape[0])
        out = self.empty(rows, self.n) if out is None else out
        op = _lib.OP_FFT if z is None else _lib.OP_H_FFT
        tail = (None if state_out is None else state_out.data_ptr(), None if intensity is None else intensity.data_ptr(),
                int(line0), rows, 1 if inverse else 0, 0.0 if z is None else float(z), self._stream())
        if tiled:
            _lib.check(self.lib.lp2d_pass_lines_tiles(plan.handle, op, x.data_ptr(), int(x.shape[2]), out.data_ptr(), *tail),
                       "pass_lines_tiles")
        else:
            _lib.check(self.lib.lp2d_pass_lines(plan.handle, op, x.data_ptr(), out.data_ptr(), *tail), "pass_lines")
        return out

    def lens(self, x, line0, k0, f):
        """x *= exp(i (X^2 + Y^2) k0 / (2 f)) in place on the row slab starting at grid row line0 (beam2d.py:508-510)."""
        if self._tabs is None:
            a, b, x2 = self._host_tabs
            self._tabs = self.torch.from_numpy(np.concatenate([a, b, x2])).to(x.device)
        _lib.check(self.lib.lp2d_pointwise(_lib.C64 if self.bits == 64 else _lib.C128, x.data_ptr(), 1, int(x.shape[0]), self.n,
                                           int(line0), _lib.PW_LENS, self._tabs.data_ptr(), float(k0), 2.0 * float(f),
                                           self._stream()), "pointwise")

    def mul_mask(self, x, mask):
        """x *= mask in place; mask: bool / real / complex array or tensor of x's shape (beam2d.py:338, 369)."""
        torch = self.torch
        t = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(mask)))
        t = t.to(x.device)
        if t.dtype == torch.bool:
            t, kind = t.to(torch.uint8), _lib.MASK_U8
        elif t.is_complex():
            t, kind = t.to(self.cdtype), _lib.MASK_COMPLEX
        else:
            t, kind = t.to(self.rdtype), _lib.MASK_REAL
        t = t.expand(x.shape).contiguous()
        _lib.check(self.lib.lp2d_mul_mask(_lib.C64 if self.bits == 64 else _lib.C128, x.data_ptr(), t.data_ptr(), kind, x.numel(), 0, 1,
                                          self._stream()), "mul_mask")

    def intensity(self, x):
        out = self.empty(*x.shape, real=True)
        _lib.check(self.lib.lp2d_intensity_n(_lib.C64 if self.bits == 64 else _lib.C128, x.data_ptr(), out.data_ptr(), x.numel(),
                                             self._stream()), "intensity_n")
        return out

    def project(self, flat_field, modes_cols):
        """partial coeffs[m] = sum_p modes_cols[m][p] * flat_field[p] over this rank's part of the crop: the mode
        projection kernel on a 1 x K "grid" (tensor cores for complex64 when the shape allows, SIMT otherwise)."""
        K = int(flat_field.numel())
        M = int(modes_cols.shape[0])
        out = self.empty(M)
        if K == 0:
            return out.zero_()
        _lib.check(self.lib.lp2d_modes_project_flat(_lib.C64 if self.bits == 64 else _lib.C128, flat_field.data_ptr(),
                                                    modes_cols.data_ptr(), out.data_ptr(), K, int(modes_cols.stride(0)), M,
                                                    self._stream()), "modes_project_flat")
        return out

    def pack_transpose(self, x, tile):
        """(rows, ld) -> (ld // tile, tile, rows): tile c transposed."""
        rows, ld = x.shape
        out = self.empty(ld // tile, tile, rows)
        _lib.check(self.lib.lp2d_pack_transpose(_lib.C64 if self.bits == 64 else _lib.C128, x.data_ptr(), out.data_ptr(),
                                                rows, ld, tile, self._stream()), "pack_transpose")
        return out


def exchange_transpose(x, backend, group=None, tiles=False):
    """All-to-all transpose of a slab-decomposed N x N array.

    ``x`` is this rank's ``(N/G, N)`` slab of A (rows ``r*N/G ..``).  Returns this rank's
    ``(N/G, N)`` slab of A^T.  The same call maps the transposed layout back.  With ``tiles=True`` (and a
    backend whose line pass reads tiles, ``backend.reads_tiles``) the receive buffer ``(G, N/G, N/G)`` is
    returned as it is -- ``out[r][i][k] = A^T[i][r*N/G + k]`` -- and the unpack copy is skipped.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nb, n = x.shape
    if nb * world != n:
        raise ValueError(f"slab shape {tuple(x.shape)} does not match {world} slabs of an {n} x {n} grid")
    send = backend.pack_transpose(x, nb)                  # [c][i][k] = x[k][c*nb + i]   (tile c, transposed)
    if world == 1:
        recv = send
    else:
        recv = send.new_empty(send.shape)
        dist.all_to_all_single(recv, send, group=group)   # recv[r] = tile (this rank's columns) of rank r's slab, transposed
    # A^T slab: out[i][r*nb + k] = recv[r][i][k]
    if world == 1:
        return recv.reshape(nb, n)
    if tiles and getattr(backend, "reads_tiles", False):
        return recv
    return recv.permute(1, 0, 2).reshape(nb, n).contiguous()


def exchange_transpose_chunked(first_pass, rows, n, backend, group=None, chunks=4):
    """The first line pass and the all-to-all transpose, OVERLAPPED: the slab's ``rows`` lines are transformed in
    ``chunks`` groups; as soon as a group is transformed and packed its tiles go out (``dist.all_to_all`` of G
    tiles, asynchronous on the communication stream) while the next group is being transformed.  The receive buffer
    is laid out so that the second pass reads it in place (``lp2d_pass_lines_tiles`` with tile = rows / chunks):
    tile index ``q * chunks + c`` = lines ``c * L .. (c+1) * L`` of sender ``q``.  ``first_pass(lo, hi)`` returns
    the transformed lines ``lo .. hi`` of this rank's slab, shape ``(hi - lo, n)``."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rows * world != n or rows % chunks:
        raise ValueError(f"slab of {rows} lines: not {n} / {world}, or not divisible into {chunks} chunks")
    L = rows // chunks
    recv = backend.empty(world, chunks, rows, L)
    works = []
    for c in range(chunks):
        part = first_pass(c * L, (c + 1) * L)                       # (L, n)
        send = backend.pack_transpose(part, rows)                   # (world, rows, L): [q][i][r'] = part[r'][q*rows + i]
        if world == 1:
            recv[0, c].copy_(send[0])
        else:
            works.append(dist.all_to_all([recv[q, c] for q in range(world)], [send[q] for q in range(world)],
                                         group=group, async_op=True))
            works.append(send)                                      # keep the send buffer alive until the wait
    for w in works:
        if hasattr(w, "wait"):
            w.wait()
    return recv.reshape(world * chunks, rows, L)


class SlabBeam2D:
    """``Beam2D`` semantics (field, spectrum, ``propagate``, ``iprofile``) for ONE grid spread over the
    ranks of a process group.  ``init_field_gen(X, Y_slab, *args)`` is called with the full 1-D ``X`` and
    this rank's column vector of ``Y`` (the reference calls generators as ``f(X, Y, *args)``,
    beam2d.py:119-120), or ``init_field`` gives this rank's ``(N/G, N)`` rows."""

    def __init__(self, area_size, npoints, wl, init_field=None, init_field_gen=None, init_gen_args=(),
                 complex_bits=64, group=None, backend=None, device=None, chunks=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.area_size, self.npoints, self.wl, self.complex_bits = area_size, int(npoints), wl, complex_bits
        self.z = 0.
        self.r0, self.r1 = slab_rows(self.npoints, self.rank, self.world)
        self.backend = backend if backend is not None else CudaLineBackend(area_size, npoints, wl, complex_bits, device)
        # overlap of the first pass with the exchange: groups of lines per all-to-all (1 = one exchange after the
        # whole pass).  Default: 4 on several GPUs with the CUDA backend (measured: profiles/r2c_slab_*.json)
        if chunks is None:
            import os
            env = os.environ.get("LP2D_SLAB_CHUNKS")
            chunks = int(env) if env else (4 if (self.world > 1 and backend is None) else 1)
        nb_ = self.npoints // self.world
        while chunks > 1 and (nb_ % chunks or (nb_ // chunks) & (nb_ // chunks - 1)):
            chunks //= 2
        self.chunks = max(1, int(chunks))
        self.dL = area_size / npoints                                       # beam2d.py:299
        self.X = np.arange(-npoints // 2, npoints // 2, 1) * self.dL        # :300-301
        self.Y = self.X.reshape((-1, 1))[self.r0:self.r1]                   # this rank's rows of :302
        self.k0 = 1 / wl
        npc = np.complex64 if complex_bits == 64 else np.complex128
        if init_field_gen is not None:
            f = np.asarray(init_field_gen(self.X.copy(), self.Y.copy(), *init_gen_args)).astype(npc)
        elif init_field is not None:
            f = init_field
        else:
            raise ValueError("Init field data is None : 'init_field_gen' must be a function or 'init_field' must be an array.")
        if isinstance(f, np.ndarray):
            f = torch.from_numpy(np.ascontiguousarray(f.astype(npc)))
        f = f.to(self.backend.empty(1).device)
        if tuple(f.shape) != (self.r1 - self.r0, self.npoints):
            raise ValueError(f"this rank's slab must have shape {(self.r1 - self.r0, self.npoints)}, got {tuple(f.shape)}")
        self.field = f.contiguous()
        self.spectrum_t = self._fft2(self.field)

    # forward 2-D transform: field rows -> transposed spectrum columns
    def _fft2(self, field):
        if self.chunks > 1:
            cols = exchange_transpose_chunked(lambda lo, hi: self.backend.lines(field[lo:hi], self.r0 + lo),
                                              int(field.shape[0]), self.npoints, self.backend, self.group, self.chunks)
            return self.backend.lines(cols, self.r0, transposed=True)
        rows = self.backend.lines(field, self.r0)                            # FFT along x
        cols = exchange_transpose(rows, self.backend, self.group, tiles=True)  # this rank's grid columns, as lines
        return self.backend.lines(cols, self.r0, transposed=True)            # FFT along y

    def propagate(self, z: float, intensity: bool = False):
        """beam2d.py:481-490 on the decomposed grid; returns this rank's rows of ``|field|^2`` when asked."""
        if z == 0:
            return None
        self.z += z
        # spectrum *= H (state kept, transposed layout) and inverse transform along y, one local pass
        if self.chunks > 1:
            st = self.spectrum_t
            rows = exchange_transpose_chunked(
                lambda lo, hi: self.backend.lines(st[lo:hi], self.r0 + lo, inverse=True, transposed=True, z=z, state_out=st[lo:hi]),
                int(st.shape[0]), self.npoints, self.backend, self.group, self.chunks)
        else:
            cols = self.backend.lines(self.spectrum_t, self.r0, inverse=True, transposed=True, z=z, state_out=self.spectrum_t)
            rows = exchange_transpose(cols, self.backend, self.group, tiles=True)
        inten = self.backend.empty(*self.field.shape, real=True) if intensity else None
        self.field = self.backend.lines(rows, self.r0, inverse=True, intensity=inten, out=self.field)
        return inten

    def _ifft2(self):
        """field = ifft2(spectrum) from the transposed spectrum slab (no H): the second half of ``propagate``."""
        if self.chunks > 1:
            st = self.spectrum_t
            rows = exchange_transpose_chunked(lambda lo, hi: self.backend.lines(st[lo:hi], self.r0 + lo, inverse=True, transposed=True),
                                              int(st.shape[0]), self.npoints, self.backend, self.group, self.chunks)
        else:
            cols = self.backend.lines(self.spectrum_t, self.r0, inverse=True, transposed=True)
            rows = exchange_transpose(cols, self.backend, self.group, tiles=True)
        self.field = self.backend.lines(rows, self.r0, inverse=True, out=self.field)

    def lens(self, f: float):
        """beam2d.py:508-511 on this rank's rows: field *= exp(i (X^2 + Y^2) k0 / (2 f)); spectrum = fft2(field)."""
        self.backend.lens(self.field, self.r0, self.k0, f)
        self.spectrum_t = self._fft2(self.field)

    def coordinate_filter(self, f_init=None, f_gen=None, fargs=()):
        """beam2d.py:335-342.  ``f_gen(X, Y_slab, *fargs)`` is evaluated on this rank's rows; ``f_init`` is this rank's
        ``(N/G, N)`` rows of the mask."""
        if f_gen is not None:
            assert callable(f_gen)
            self.backend.mul_mask(self.field, np.asarray(f_gen(self.X.copy(), self.Y.copy(), *fargs)))
        if f_init is not None:
            self.backend.mul_mask(self.field, f_init)
        self.spectrum_t = self._fft2(self.field)

    def spectral_filter(self, f_init=None, f_gen=None, fargs=()):
        """beam2d.py:366-373.  This rank holds the spectrum COLUMNS ``r0 .. r1`` transposed: ``f_gen`` is called as
        ``f_gen(Kx[r0:r1], Ky, *fargs)`` (an ``(N, N/G)`` block of the reference's mask) and transposed; ``f_init``
        is that block, ``mask[:, r0:r1]``, of the full ``(N, N)`` mask."""
        Kx = np.fft.fftfreq(self.npoints, d=self.dL)                        # beam2d.py:305
        if f_gen is not None:
            assert callable(f_gen)
            blk = np.asarray(f_gen(Kx[self.r0:self.r1].copy(), Kx.reshape((-1, 1)).copy(), *fargs))
            self.backend.mul_mask(self.spectrum_t, np.ascontiguousarray(blk.T))
        if f_init is not None:
            blk = f_init.T if isinstance(f_init, np.ndarray) else f_init.transpose(0, 1)
            self.backend.mul_mask(self.spectrum_t, blk)
        self._ifft2()

    @property
    def iprofile(self):
        """This rank's rows of ``abs(field)**2`` (beam2d.py:717)."""
        return self.backend.intensity(self.field)

    def deconstruct_by_modes(self, modes_list):
        """beam2d.py:592-600 ('scalar') on the decomposed grid: every rank projects
        the part of the centre ``nb x nb`` crop that lies in its rows onto the matching columns of the mode
        matrix, the partial coefficient vectors are summed with ONE all-reduce (``nmodes`` complex numbers)."""
        import torch
        import torch.distributed as dist
        dev = self.field.device
        M = modes_list if isinstance(modes_list, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(modes_list)))
        if M.dim() != 2:
            M = M.reshape(M.shape[0], -1)
        nb = int(np.sqrt(M.shape[1]))
        dN = (self.npoints - nb) // 2                                        # beam2d.py:593
        lo, hi = max(self.r0, dN), min(self.r1, dN + nb)                     # grid rows of the crop owned here
        if hi > lo:
            strip = self.field[lo - self.r0:hi - self.r0, dN:dN + nb].contiguous().reshape(-1)
            cols = M[:, (lo - dN) * nb:(hi - dN) * nb].to(dev).to(self.field.dtype)
            part = self.backend.project(strip, cols)
        else:
            part = self.backend.empty(int(M.shape[0])).zero_()
        if self.world > 1:
            buf = torch.view_as_real(part)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        self.modes_coeffs = part
        return part

    def gather_field(self):
        """The full N x N field on every rank (debug / small grids)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.field
        parts = [torch.empty_like(self.field) for _ in range(self.world)]
        dist.all_gather(parts, self.field, group=self.group)
        return torch.cat(parts, dim=0)

    def gather_spectrum(self):
        """The full N x N spectrum (natural layout) on every rank (debug / small grids)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.spectrum_t.transpose(0, 1).contiguous()
        parts = [torch.empty_like(self.spectrum_t) for _ in range(self.world)]
        dist.all_gather(parts, self.spectrum_t, group=self.group)
        return torch.cat(parts, dim=0).transpose(0, 1).contiguous()
