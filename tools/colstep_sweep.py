"""Column pass alone (lp2d_pass, in place) for the two-step variants: us per beam and fraction of the HBM peak on 2 P B."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, _lib  # noqa: E402
from lightprop2d_b200 import beam2d as _b  # noqa: E402


def timeit(fn, iters=10, warm=3):
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
n, bits, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ct = torch.complex64 if bits == 64 else torch.complex128
f = torch.randn(batch, n, n, dtype=ct, device="cuda")
for cs, var, grid in [("0", "0", "4")] + [("1", v, g) for v in ("0", "1", "2") for g in ("4", "8", "0")]:
    os.environ["LP2D_COLSTEP"] = cs; os.environ["LP2D_COLSTEP_VARIANT"] = var; os.environ["LP2D_COLSTEP_GRID"] = grid
    _b._PLAN_CACHE.clear()
    b = Beam2D(0.4 * n / 1024, n, 532e-7, init_field=f[:1], complex_bits=bits, use_gpu=True)
    plan = b._plan.handle
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    t = timeit(lambda: lib.lp2d_pass(plan, _lib.AXIS_COLS, _lib.OP_FFT, f.data_ptr(), f.data_ptr(), None, None, batch, 1, 0.0, st))
    by = 2.0 * n * n * (8 if bits == 64 else 16)
    print(f"n={n} c{bits} batch={batch} colstep={cs} variant={var} grid={grid}: {t/batch:8.1f} us/beam  {by/(t/batch*1e-6)/6534.8e9:.3f} of HBM on 2PB", flush=True)
    del b
