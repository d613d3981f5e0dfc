"""Small fixed workload for ncu captures of the mode kernels (config 5): projection (tensor cores, complex64),
synthesis (tensor cores) and the SIMT complex128 projection, 256 modes x 512^2, 64 / 16 beams."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightprop2d_b200 as lp  # noqa: E402

n, M = 512, 256
for bits, Bt in ((64, 64), (128, 16)):
    ct = torch.complex64 if bits == 64 else torch.complex128
    modes = torch.randn(M, n * n, dtype=ct, device="cuda")
    fld = torch.randn(Bt, n, n, dtype=ct, device="cuda")
    b = lp.Beam2D(0.04, n, 632e-7, init_field=fld, complex_bits=bits, use_gpu=True, numpy_output=False)
    for _ in range(2):
        c = b.deconstruct_by_modes(modes)
    for _ in range(2):
        b.construct_by_modes(modes, c)
    torch.cuda.synchronize()
print("done")
