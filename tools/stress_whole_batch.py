"""Whole-batch passes + H plane + two-step columns against explicitly chunked passes with the inline H, large batches."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, _lib  # noqa: E402
from lightprop2d_b200 import beam2d as _b  # noqa: E402

lib = _lib.load()
for n, bits, B in ((2048, 64, 301), (1024, 128, 257), (4096, 128, 9), (8192, 64, 5)):
    ct = torch.complex64 if bits == 64 else torch.complex128
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    s0 = torch.randn(B, n, n, dtype=ct, device="cuda", generator=g)
    res = {}
    for mode in ("old", "new"):
        for k in ("LP2D_CHUNK_MB", "LP2D_HPLANE", "LP2D_COLSTEP"):
            os.environ.pop(k, None)
        if mode == "old":
            os.environ.update({"LP2D_CHUNK_MB": "32", "LP2D_HPLANE": "0", "LP2D_COLSTEP": "0"})
        _b._PLAN_CACHE.clear()
        b = Beam2D(4e-1 * n / 1024, n, 532e-7, init_field=s0[:1], complex_bits=bits, use_gpu=True)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        s = s0.clone(); f = torch.empty_like(s)
        for z in (0.02, 0.02, 0.02, -0.01):
            _lib.check(lib.lp2d_propagate(b._plan.handle, s.data_ptr(), f.data_ptr(), None, B, z, st))
        torch.cuda.synchronize()
        res[mode] = (f, s)
        del b
    ef = float((res["new"][0] - res["old"][0]).norm() / res["old"][0].norm())
    es = float((res["new"][1] - res["old"][1]).norm() / res["old"][1].norm())
    print(f"n={n} c{bits} batch={B}: new defaults vs chunked / inline H / one-kernel columns: rel L2 field {ef:.2e} spectrum {es:.2e}", flush=True)
    assert ef < (3e-6 if bits == 64 else 1e-14) and es < (3e-6 if bits == 64 else 1e-14)
    del res, s0, s, f
    torch.cuda.empty_cache()
print("stress3 ok")
