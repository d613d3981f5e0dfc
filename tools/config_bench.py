"""Device-side timings of the BASELINE.json configs 2, 3 and 5 on one GPU (CUDA events); one JSON object."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightprop2d_b200 as lp  # noqa: E402
from lightprop2d_b200 import _lib  # noqa: E402
from lightprop2d_b200.ensemble import EnsemblePropagator  # noqa: E402

HBM = 6535e9


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
    return a.elapsed_time(b) / iters


def main():
    res = {}
    # config 2: single complex128 field, lens + z-sweep of propagate(dz)
    for n in (1024, 2048, 4096):
        b = lp.Beam2D(4e-1, n, 532e-7, init_field_gen=lp.gaussian_beam, init_gen_args=(1, 0.05), complex_bits=128,
                      use_gpu=True, numpy_output=False)
        b.lens(10.0)
        ms = timeit(lambda: b.propagate(0.02), iters=20)
        ml = timeit(lambda: b.lens(1e6), iters=5)
        P = n * n * 16
        res[f"cfg2_c128_{n}"] = {"propagate_ms": ms, "propagate_frac_3PB": 3 * P / (ms * 1e-3) / HBM,
                                 "lens_ms": ml, "lens_frac_3PB": 3 * P / (ml * 1e-3) / HBM}
        del b
    # config 3: ensemble 512^2 complex64, fft2 -> H -> ifft2 -> |E|^2
    n, B = 512, 1024
    f = torch.randn(B, n, n, dtype=torch.complex64, device="cuda")
    prop = EnsemblePropagator(200e-4 * 2, n, 632e-7, complex_bits=64)
    out = torch.empty(B, n, n, dtype=torch.float32, device="cuda")
    ms = timeit(lambda: prop.intensity(f, 100e-4, out=out))
    res["cfg3_c64_512_ensemble"] = {"beams": B, "ms": ms, "beams_per_s": B / (ms * 1e-3),
                                    "frac_PB_plus_PR": B * n * n * 12 / (ms * 1e-3) / HBM}
    del f, out, prop
    # config 5: mode projection / synthesis (SIMT in this round), M modes x 512^2, batch Bt
    n, nb = 512, 512
    for M, Bt in ((256, 1), (256, 64)):
        modes = torch.randn(M, nb * nb, dtype=torch.complex64, device="cuda")
        fld = torch.randn(Bt, n, n, dtype=torch.complex64, device="cuda")
        b = lp.Beam2D(200e-4 * 2, n, 632e-7, init_field=fld, complex_bits=64, use_gpu=True, numpy_output=False)
        ms = timeit(lambda: b.deconstruct_by_modes(modes), iters=5)
        flops = 8.0 * Bt * M * nb * nb
        byts = (M * nb * nb + Bt * nb * nb + Bt * M) * 8
        res[f"cfg5_project_M{M}_Bt{Bt}"] = {"ms": ms, "TFLOPs": flops / (ms * 1e-3) / 1e12, "GBps_alg": byts / (ms * 1e-3) / 1e9}
        del b, modes, fld
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
