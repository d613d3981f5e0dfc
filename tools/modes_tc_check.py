"""Tensor-core mode projection against the SIMT kernel (FP64 accumulation) and numpy; timings."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightprop2d_b200 as lp  # noqa: E402


def timeit(fn, iters=10, warm=2):
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
    bad = 0
    for n, nb, M, Bt in ((64, 32, 5, 3), (128, 64, 130, 1), (256, 128, 96, 70), (512, 512, 256, 64), (512, 256, 300, 65)):
        g = torch.Generator(device="cuda"); g.manual_seed(n + M)
        modes = torch.randn(M, nb * nb, dtype=torch.complex64, device="cuda", generator=g)
        fld = torch.randn(Bt, n, n, dtype=torch.complex64, device="cuda", generator=g)
        b = lp.Beam2D(200e-4 * 2, n, 632e-7, init_field=fld, complex_bits=64, use_gpu=True, numpy_output=False)
        os.environ["LP2D_MODES_TC"] = "0"
        ref = b.deconstruct_by_modes(modes).clone()
        os.environ["LP2D_MODES_TC"] = "1"
        got = b.deconstruct_by_modes(modes).clone()
        torch.cuda.synchronize()
        dn = (n - nb) // 2
        crop = fld[:, dn:dn + nb, dn:dn + nb].reshape(Bt, -1).to(torch.complex128)
        exact = crop @ modes.to(torch.complex128).T
        e_tc = float((got.to(torch.complex128) - exact).norm() / exact.norm())
        e_simt = float((ref.to(torch.complex128) - exact).norm() / exact.norm())
        ok = e_tc <= 1e-5
        bad += (not ok)
        t_tc = timeit(lambda: b.deconstruct_by_modes(modes), iters=5)
        os.environ["LP2D_MODES_TC"] = "0"
        t_simt = timeit(lambda: b.deconstruct_by_modes(modes), iters=3)
        flops = 8.0 * Bt * M * nb * nb
        byts = (M * nb * nb + Bt * nb * nb) * 8
        print(f"n={n} nb={nb} M={M} Bt={Bt}: {'OK ' if ok else 'BAD'} rel_l2 tc={e_tc:.2e} simt={e_simt:.2e} | "
              f"tc {t_tc*1e3:8.1f} us ({flops/t_tc/1e9:7.2f} TFLOP/s, {byts/t_tc/1e6:7.1f} GB/s)  simt {t_simt*1e3:8.1f} us", flush=True)
        del b
    print("failures:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
