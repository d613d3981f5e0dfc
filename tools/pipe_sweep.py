"""Sweep the knobs of the fused step kernel (beams per item, lag, ring slots, hints, variant, which steps use it)."""
import ctypes as C
import itertools
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
    combos = sys.argv[4] if len(sys.argv) > 4 else "PIPE=2;G=0,1,2;LAG=0;SLOTS=0"
    keys, vals = [], []
    for kv in combos.split(";"):
        k, v = kv.split("=")
        keys.append(k)
        vals.append(v.split(","))
    ct = torch.complex64 if bits == 64 else torch.complex128
    B = 8 if bits == 64 else 16
    f = torch.randn(batch, n, n, dtype=ct, device="cuda")
    s = torch.empty_like(f); o = torch.empty_like(f)
    lib = _lib.load()
    P = n * n * B * batch
    hbm = 6535.1e9
    print(f"n={n} bits={bits} batch={batch}; us/beam (fraction of {hbm/1e9:.0f} GB/s on algorithmic bytes)")
    print(" ".join(f"{k:>6s}" for k in keys) + " | fft2(2PB) | stateful(3PB) | fresh f->f(2PB)")
    envname = {"G": "LP2D_PIPE_G", "LAG": "LP2D_PIPE_LAG", "SLOTS": "LP2D_PIPE_SLOTS", 
               "RING": "LP2D_RING_MB", "PIPE": "LP2D_PIPE", "VAR": "LP2D_PIPE_VARIANT", "HINTS": "LP2D_PIPE_HINTS", "PERSIST": "LP2D_PIPE_PERSIST", "OPS": "LP2D_FUSED_OPS", "TMA": "LP2D_TMA", "FORM": "LP2D_PIPE_FORM", "QLAG": "LP2D_Q4_LAG", "QVAR": "LP2D_Q4_VARIANT", "CP": "LP2D_Q4_COLPRIV"}
    for combo in itertools.product(*vals):
        for k, v in zip(keys, combo):
            os.environ[envname[k]] = v
        _b._PLAN_CACHE.clear()
        b = Beam2D(200 * um * n / 256, n, 632 * nm, init_field=f[:1], complex_bits=bits, use_gpu=True)
        plan = b._plan.handle
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        t0 = timeit(lambda: lib.lp2d_fft2(plan, f.data_ptr(), s.data_ptr(), batch, 0, st))
        t1 = timeit(lambda: lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), None, batch, 1e-3, st))
        t2 = timeit(lambda: lib.lp2d_propagate_field(plan, f.data_ptr(), o.data_ptr(), None, None, None, 0, batch, 1e-3, st))
        fr = lambda t, k: k * P / (t * 1e-3) / hbm  # noqa: E731
        u = lambda t: t * 1e3 / batch  # noqa: E731
        print(" ".join(f"{v:>6s}" for v in combo) + f" | {u(t0):6.2f} {fr(t0,2):.2f} | {u(t1):6.2f} {fr(t1,3):.2f} | {u(t2):6.2f} {fr(t2,2):.2f}", flush=True)
        del b


if __name__ == "__main__":
    main()
