"""Time the cluster line kernel (xl_line_kernel) variants (LP2D_XL_VARIANT): plain line FFT (read + write) and
the H + state + inverse pass (read + 2 writes) over 1 GiB of lines; checks every variant against variant 0."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200.slab import CudaLineBackend  # noqa: E402


def timeit(fn, iters=5, warm=2):
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
    cases = [(64, 32768), (64, 16384), (128, 16384), (128, 8192)]
    if len(sys.argv) > 1:
        cases = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1].split(",")]
    hbm = 6534.8e9
    for bits, n in cases:
        B = 8 if bits == 64 else 16
        nl = min(n, (1 << 30) // (n * B))
        x = torch.randn(nl, n, dtype=torch.complex64 if bits == 64 else torch.complex128, device="cuda")
        out, st = torch.empty_like(x), torch.empty_like(x)
        ref = None
        for var in [int(v) for v in (sys.argv[2].split(',') if len(sys.argv) > 2 else '0,1,2,3'.split(','))]:
            os.environ["LP2D_XL_VARIANT"] = str(var)
            be = CudaLineBackend(0.4 * n / 1024, n, 532e-7, bits)
            name = be.plan.step_kernel_name() if hasattr(be.plan, "step_kernel_name") else ""
            t_f = timeit(lambda: be.lines(x, 0, out=out))
            t_h = timeit(lambda: be.lines(x, 0, inverse=True, z=0.02, state_out=st, out=out))
            torch.cuda.synchronize()
            if ref is None:
                ref, err = out.clone(), 0.0
            else:
                err = float((out - ref).abs().max() / ref.abs().max())
            by = nl * n * B
            print(f"c{bits} n={n} lines={nl} var={var} {name} | fft {t_f*1e3:8.1f} us ({2*by/(t_f*1e-3)/hbm:.2f} of HBM) | "
                  f"H+state+ifft {t_h*1e3:8.1f} us ({3*by/(t_h*1e-3)/hbm:.2f}) | err vs var0 {err:.1e}", flush=True)
            del be


if __name__ == "__main__":
    main()
