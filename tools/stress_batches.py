import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from lightprop2d_b200 import Beam2D, _lib
from lightprop2d_b200 import beam2d as _b
lib = _lib.load()
for n, bits, B in ((256, 64, 30001), (512, 64, 4097), (256, 128, 9000)):
    ct = torch.complex64 if bits == 64 else torch.complex128
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    s0 = torch.randn(B, n, n, dtype=ct, device='cuda', generator=g)
    res = {}
    for mode in ('0', '2'):
        os.environ['LP2D_PIPE'] = mode
        _b._PLAN_CACHE.clear()
        b = Beam2D(4e-1 * n / 1024, n, 532e-7, init_field=s0[:1], complex_bits=bits, use_gpu=True)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        s = s0.clone(); f = torch.empty_like(s)
        for z in (0.02, -0.01):
            _lib.check(lib.lp2d_propagate(b._plan.handle, s.data_ptr(), f.data_ptr(), None, B, z, st))
        torch.cuda.synchronize()
        res[mode] = (f, s)
        del b
    ef = float((res['2'][0] - res['0'][0]).norm() / res['0'][0].norm())
    es = float((res['2'][1] - res['0'][1]).norm() / res['0'][1].norm())
    print(f"n={n} c{bits} batch={B}: fused vs per-pass rel L2 field {ef:.2e} spectrum {es:.2e}", flush=True)
    assert ef < (3e-6 if bits == 64 else 1e-14) and es < (3e-6 if bits == 64 else 1e-14)
    del res, s0, s, f
    torch.cuda.empty_cache()
print("stress ok")
