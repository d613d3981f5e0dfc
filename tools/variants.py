"""Time the tiling variants (LP2D_VARIANT) of one grid size and check them against variant 0."""
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
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    variants = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1, 2, 3, 4, 5]
    envs = sys.argv[5].split(";") if len(sys.argv) > 5 else [""]
    ct = torch.complex64 if bits == 64 else torch.complex128
    B = 8 if bits == 64 else 16
    f = torch.randn(batch, n, n, dtype=ct, device="cuda")
    s = torch.empty_like(f); o = torch.empty_like(f); o2 = torch.empty_like(f)
    lib = _lib.load()
    P = n * n * B * batch
    hbm = 6542.1e9
    ref = None
    print(f"n={n} bits={bits} batch={batch}; us/beam (fraction of {hbm/1e9:.0f} GB/s on the pass's own bytes)")
    print("var env | rowF  colF | rowH(+state) colI | fresh-col | stateful-step(alg 3PB) fresh-step(alg 2PB) | maxerr vs var0")
    for env in envs:
        for kv in env.split():
            k, v = kv.split("=")
            os.environ[k] = v
        for var in variants:
            os.environ["LP2D_VARIANT"] = str(var)
            _b._PLAN_CACHE.clear()
            b = Beam2D(200 * um * n / 256, n, 632 * nm, init_field=f[:1], complex_bits=bits, use_gpu=True)
            plan = b._plan.handle
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            R, Cc = _lib.AXIS_ROWS, _lib.AXIS_COLS
            t_rf = timeit(lambda: lib.lp2d_pass(plan, R, _lib.OP_FFT, f.data_ptr(), s.data_ptr(), None, None, batch, 0, 0.0, st))
            t_cf = timeit(lambda: lib.lp2d_pass(plan, Cc, _lib.OP_FFT, s.data_ptr(), s.data_ptr(), None, None, batch, 0, 0.0, st))
            t_rh = timeit(lambda: lib.lp2d_pass(plan, R, _lib.OP_H_FFT, s.data_ptr(), o.data_ptr(), s.data_ptr(), None, batch, 1, 1e-3, st))
            t_ci = timeit(lambda: lib.lp2d_pass(plan, Cc, _lib.OP_FFT, o.data_ptr(), o.data_ptr(), None, None, batch, 1, 0.0, st))
            t_cm = timeit(lambda: lib.lp2d_pass(plan, Cc, _lib.OP_FFT_H_IFFT, o.data_ptr(), o.data_ptr(), None, None, batch, 0, 1e-3, st))
            lib.lp2d_fft2(plan, f.data_ptr(), s.data_ptr(), batch, 0, st)
            t_st = timeit(lambda: lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), None, batch, 1e-3, st))
            t_fr = timeit(lambda: lib.lp2d_propagate_field(plan, f.data_ptr(), o2.data_ptr(), None, None, None, 0, batch, 1e-3, st))
            torch.cuda.synchronize()
            if ref is None:
                ref = o2.clone()
                err = 0.0
            else:
                err = float((o2 - ref).abs().max() / ref.abs().max())
            u = lambda t: t * 1e3 / batch  # noqa: E731
            fr = lambda t, k: k * P / (t * 1e-3) / hbm  # noqa: E731
            print(f"{var:3d} {env:18s} | {u(t_rf):5.2f}({fr(t_rf,2):.2f}) {u(t_cf):5.2f}({fr(t_cf,2):.2f}) | "
                  f"{u(t_rh):5.2f}({fr(t_rh,3):.2f}) {u(t_ci):5.2f}({fr(t_ci,2):.2f}) | {u(t_cm):5.2f}({fr(t_cm,2):.2f}) | "
                  f"{u(t_st):5.2f}({fr(t_st,3):.2f}) {u(t_fr):5.2f}({fr(t_fr,2):.2f}) | {err:.1e}", flush=True)
            del b


if __name__ == "__main__":
    main()
