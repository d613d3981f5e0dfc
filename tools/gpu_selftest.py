"""Diagnostic sweep on a GPU box: every size / dtype / op against numpy, all results printed.

    python tools/gpu_selftest.py [quick]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, um, nm  # noqa: E402
from oracle import beam_oracle as orc  # noqa: E402


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print(torch.cuda.get_device_name(0), flush=True)
    sizes = [16, 32, 64, 128, 256, 512, 1024, 2048] + ([] if quick else [4096])
    if os.environ.get("LP2D_SELFTEST_SIZES"):
        sizes = [int(v) for v in os.environ["LP2D_SELFTEST_SIZES"].split(",")]
    bad = 0
    for bits in (64, 128):
        tol = 1e-5 if bits == 64 else 1e-12
        ct = np.complex64 if bits == 64 else np.complex128
        for n in sizes:
            if bits == 128 and n > 4096:
                continue
            rng = np.random.default_rng(n + bits)
            B = int(os.environ.get("LP2D_SELFTEST_BATCH", 2 if n <= 1024 else 1))
            f0 = (rng.standard_normal((B, n, n)) + 1j * rng.standard_normal((B, n, n))).astype(ct)
            area, wl = 200 * um * n / 256, 632 * nm
            z = 100 * um
            try:
                b = Beam2D(area, n, wl, init_field=f0.copy(), complex_bits=bits)
                o = orc.OracleBeam(area, n, wl, init_field=f0.copy(), complex_bits=bits)
                e_fft = rel(b.spectrum, o.spectrum)
                b.propagate(z); o.propagate(z)
                e_p1 = rel(b.field, o.field); e_s1 = rel(b.spectrum, o.spectrum)
                b.propagate(-0.4 * z); o.propagate(-0.4 * z)
                e_p2 = rel(b.field, o.field)
                e_i = rel(b.iprofile, o.iprofile)
                b.lens(0.05); o.lens(0.05)
                e_l = rel(b.field, o.field); e_ls = rel(b.spectrum, o.spectrum)
                # fresh chain (lazy spectrum) on a new object
                b2 = Beam2D(area, n, wl, init_field=f0.copy(), complex_bits=bits)
                b2.propagate(z)
                o2 = orc.OracleBeam(area, n, wl, init_field=f0.copy(), complex_bits=bits)
                o2.propagate(z)
                e_f = rel(b2.field, o2.field); e_fs = rel(b2.spectrum, o2.spectrum)
                errs = dict(fft2=e_fft, prop=e_p1, spec=e_s1, prop2=e_p2, ipro=e_i, lens=e_l, lens_s=e_ls,
                            fresh=e_f, fresh_s=e_fs)
                ok = all(v <= tol for v in errs.values())
                bad += (not ok)
                print(f"c{bits} n={n:5d} {'OK ' if ok else 'BAD'} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
            except Exception as ex:  # noqa: BLE001
                bad += 1
                print(f"c{bits} n={n:5d} EXC {type(ex).__name__}: {ex}", flush=True)
    print("selftest failures:", bad, flush=True)
    return bad


if __name__ == "__main__":
    t = time.time()
    rc = main()
    print(f"elapsed {time.time() - t:.1f}s")
    sys.exit(1 if rc else 0)
