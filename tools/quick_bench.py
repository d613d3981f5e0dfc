"""Quick device-side timing of the main passes (CUDA events), for tuning runs.

    python tools/quick_bench.py [n] [bits] [batch]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, um, nm, _lib  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    ct = torch.complex64 if bits == 64 else torch.complex128
    B = 8 if bits == 64 else 16
    f = torch.randn(batch, n, n, dtype=ct, device="cuda")
    b = Beam2D(200 * um * n / 256, n, 632 * nm, init_field=f, complex_bits=bits, use_gpu=True)
    lib = _lib.load()
    plan = b._plan.handle
    s = torch.empty_like(f); o = torch.empty_like(f)
    inten = torch.empty(batch, n, n, dtype=torch.float32 if bits == 64 else torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = n * n * B * batch
    hbm = 6542.1e9
    def rep(name, ms, alg_bytes):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        print(f"{name:28s} {ms*1e3/batch:9.2f} us/beam  {ms:8.3f} ms  alg {gbs:8.1f} GB/s  frac {gbs*1e9/hbm:5.2f}", flush=True)
    ms = timeit(lambda: lib.lp2d_fft2(plan, f.data_ptr(), s.data_ptr(), batch, 0, st))
    rep("fft2 (2 passes)", ms, 2 * P)
    ms = timeit(lambda: lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), None, batch, 1e-3, st))
    rep("propagate stateful", ms, 3 * P)
    ms = timeit(lambda: lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), inten.data_ptr(), batch, 1e-3, st))
    rep("propagate stateful+I", ms, 3 * P + P // 2)
    ms = timeit(lambda: lib.lp2d_propagate_field(plan, f.data_ptr(), o.data_ptr(), None, None, None, 0, batch, 1e-3, st))
    rep("fresh chain field->field", ms, 2 * P)
    work = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: lib.lp2d_propagate_field(plan, f.data_ptr(), None, None, inten.data_ptr(), work.data_ptr(), work.numel(), batch, 1e-3, st))
    rep("fresh chain field->I", ms, P + P // 2)
    ms = timeit(lambda: lib.lp2d_intensity(plan, f.data_ptr(), inten.data_ptr(), batch, st))
    rep("intensity", ms, P + P // 2)
    ms = timeit(lambda: o.copy_(f))
    rep("torch copy (ref)", ms, 2 * P)


if __name__ == "__main__":
    main()
