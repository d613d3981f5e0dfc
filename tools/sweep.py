"""Sweep the host-side scheduling knobs (chunk size, aux streams, cache hints) on a GPU box."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, um, nm, _lib  # noqa: E402
from lightprop2d_b200 import beam2d as _b  # noqa: E402


def timeit(fn, iters=10, warm=3):
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
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    ct = torch.complex64 if bits == 64 else torch.complex128
    B = 8 if bits == 64 else 16
    f = torch.randn(batch, n, n, dtype=ct, device="cuda")
    s = torch.empty_like(f); o = torch.empty_like(f)
    inten = torch.empty(batch, n, n, dtype=torch.float32 if bits == 64 else torch.float64, device="cuda")
    work = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    P = n * n * B * batch
    hbm = 6542.1e9
    print(f"n={n} bits={bits} batch={batch}")
    print("chunkMB streams hints |  stateful us/beam frac | fresh f->f us/beam frac | fresh f->I us/beam frac")
    configs = [(0, 1, 0), (0, 1, 1)]
    for ch in (8, 16, 32, 64, 128):
        for ns in (1, 2, 4, 8):
            for h in (0, 1):
                configs.append((ch, ns, h))
    if len(sys.argv) > 4 and sys.argv[4] == "short":
        configs = [(0, 1, 1), (16, 4, 1), (32, 4, 1), (64, 4, 1), (32, 1, 1), (32, 4, 0), (16, 8, 1), (8, 8, 1)]
    for ch, ns, h in configs:
        os.environ["LP2D_CHUNK_MB"] = str(ch)
        os.environ["LP2D_STREAMS"] = str(ns)
        os.environ["LP2D_HINTS"] = str(h)
        _b._PLAN_CACHE.clear()
        b = Beam2D(200 * um * n / 256, n, 632 * nm, init_field=f[:1], complex_bits=bits, use_gpu=True)
        plan = b._plan.handle
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        t1 = timeit(lambda: lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), None, batch, 1e-3, st))
        t2 = timeit(lambda: lib.lp2d_propagate_field(plan, f.data_ptr(), o.data_ptr(), None, None, None, 0, batch, 1e-3, st))
        t3 = timeit(lambda: lib.lp2d_propagate_field(plan, f.data_ptr(), None, None, inten.data_ptr(), work.data_ptr(), (ch if ch else 256) << 20, batch, 1e-3, st))
        fr = lambda t, bytes_: bytes_ / (t * 1e-3) / hbm  # noqa: E731
        print(f"{ch:7d} {ns:7d} {h:5d} | {t1*1e3/batch:8.2f} {fr(t1, 3*P):5.2f} | {t2*1e3/batch:8.2f} {fr(t2, 2*P):5.2f} | "
              f"{t3*1e3/batch:8.2f} {fr(t3, 1.5*P):5.2f}", flush=True)
        del b


if __name__ == "__main__":
    main()
