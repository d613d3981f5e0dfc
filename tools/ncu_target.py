"""Small fixed workload for ncu captures: stateful propagate + fresh chain, 1024^2 complex64."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, um, nm, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ct = torch.complex64 if bits == 64 else torch.complex128
f = torch.randn(batch, n, n, dtype=ct, device="cuda")
b = Beam2D(4e-1, n, 532e-7, init_field=f, complex_bits=bits, use_gpu=True)
lib = _lib.load()
plan = b._plan.handle
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
s = torch.empty_like(f); o = torch.empty_like(f)
lib.lp2d_fft2(plan, f.data_ptr(), s.data_ptr(), batch, 0, st)
for _ in range(iters):
    lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), None, batch, 0.02, st)
for _ in range(iters):
    lib.lp2d_propagate_field(plan, f.data_ptr(), o.data_ptr(), None, None, None, 0, batch, 0.02, st)
torch.cuda.synchronize()
print("done")
