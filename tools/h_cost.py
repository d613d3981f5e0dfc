"""Cost of evaluating the transfer function in the row pass of the one-kernel-per-pass path (single field and batch):
lp2d_pass rows with OP_H_FFT against OP_FFT, same buffers."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, _lib  # noqa: E402
from lightprop2d_b200 import beam2d as _b  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


lib = _lib.load()
for n, bits, batches in ((1024, 128, (1, 16)), (2048, 128, (1, 8)), (4096, 128, (1, 4)), (4096, 64, (1, 8)), (8192, 64, (1, 2)), (2048, 64, (1, 16))):
    ct = torch.complex64 if bits == 64 else torch.complex128
    for batch in batches:
        _b._PLAN_CACHE.clear()
        f = torch.randn(batch, n, n, dtype=ct, device="cuda")
        o = torch.empty_like(f); s2 = torch.empty_like(f)
        b = Beam2D(0.4 * n / 1024, n, 532e-7, init_field=f[:1], complex_bits=bits, use_gpu=True)
        plan = b._plan.handle
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        t_fft = timeit(lambda: lib.lp2d_pass(plan, _lib.AXIS_ROWS, _lib.OP_FFT, f.data_ptr(), o.data_ptr(), s2.data_ptr(), None, batch, 1, 0.0, st))
        t_h = timeit(lambda: lib.lp2d_pass(plan, _lib.AXIS_ROWS, _lib.OP_H_FFT, f.data_ptr(), o.data_ptr(), s2.data_ptr(), None, batch, 1, 0.02, st))
        t_col = timeit(lambda: lib.lp2d_pass(plan, _lib.AXIS_COLS, _lib.OP_FFT, o.data_ptr(), o.data_ptr(), None, None, batch, 1, 0.0, st))
        t_step = timeit(lambda: lib.lp2d_propagate(plan, f.data_ptr(), o.data_ptr(), None, batch, 0.02, st))
        print(f"n={n} c{bits} batch={batch}: rows+state {t_fft/batch:8.1f} us/beam, rows+H+state {t_h/batch:8.1f}, columns {t_col/batch:8.1f}, propagate {t_step/batch:8.1f}", flush=True)
        del b, f, o, s2
