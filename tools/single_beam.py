"""Single-field z-sweep step (config 2): Beam2D.propagate vs the bare C-ABI call, per size / dtype."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightprop2d_b200 as lp  # noqa: E402
from lightprop2d_b200 import _lib  # noqa: E402


def ev(fn, iters=50, warm=5):
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
for bits in (64, 128):
    for n in (256, 512, 1024, 2048, 4096):
        b = lp.Beam2D(4e-1, n, 532e-7, init_field_gen=lp.gaussian_beam, init_gen_args=(1, 0.05), complex_bits=bits,
                      use_gpu=True, numpy_output=False)
        b.propagate(0.02)
        t_py = ev(lambda: b.propagate(0.02))
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        plan, s, f = b._plan.handle, b._s, b._f
        t_c = ev(lambda: lib.lp2d_propagate(plan, s.data_ptr(), f.data_ptr(), None, 1, 0.02, st))
        t0 = time.perf_counter()
        for _ in range(200):
            b.propagate(0.02)
        host = (time.perf_counter() - t0) / 200 * 1e6
        torch.cuda.synchronize()
        b.propagate_sweep(0.02, 20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b.propagate_sweep(0.02, 500, probe="centre")
        torch.cuda.synchronize()
        sweep = (time.perf_counter() - t0) / 500 * 1e6
        P = n * n * (8 if bits == 64 else 16)
        print(f"c{bits} n={n:5d}: Beam2D.propagate {t_py:8.1f} us  C ABI {t_c:8.1f} us  host-side {host:6.1f} us/call  sweep(one call, incl. probe) {sweep:6.1f} us/step  roofline(3PB) {3*P/6535e9*1e6:7.1f} us", flush=True)
        del b
