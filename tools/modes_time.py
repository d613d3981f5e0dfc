"""Time of the mode projection / synthesis (config 5 shape) per dtype: ms per call, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightprop2d_b200 as lp  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


n, M = 512, 256
for bits, Bt in ((64, 64), (128, 16), (128, 64), (128, 1)):
    ct = torch.complex64 if bits == 64 else torch.complex128
    modes = torch.randn(M, n * n, dtype=ct, device="cuda")
    fld = torch.randn(Bt, n, n, dtype=ct, device="cuda")
    b = lp.Beam2D(0.04, n, 632e-7, init_field=fld, complex_bits=bits, use_gpu=True, numpy_output=False)
    c = b.deconstruct_by_modes(modes)
    ref = torch.einsum("mk,bk->bm", modes, fld.reshape(Bt, -1)) if Bt > 1 else (modes @ fld.reshape(-1))
    err = float((torch.as_tensor(c).reshape(ref.shape) - ref).abs().max() / ref.abs().max())
    tp = timeit(lambda: b.deconstruct_by_modes(modes))
    ts = timeit(lambda: b.construct_by_modes(modes, c))
    B = 8 if bits == 64 else 16
    alg = (M * n * n + Bt * n * n) * B
    print(f"c{bits} {M} modes x {n}^2 x {Bt} beams: project {tp:.3f} ms ({alg/tp/1e6:.0f} GB/s alg, err {err:.1e}), synthesize(+fft2) {ts:.3f} ms", flush=True)
    del b, modes, fld
