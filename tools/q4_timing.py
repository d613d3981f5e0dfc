"""Phase clocks of the persistent step kernel (LP2D_Q4_TIMING=1): where the transform warps spend their time."""
import ctypes as C
import os
import sys

os.environ["LP2D_Q4_TIMING"] = "1"
os.environ.setdefault("LP2D_PIPE_FORM", "4")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, _lib  # noqa: E402

n, batch = 1024, 256
f = torch.randn(batch, n, n, dtype=torch.complex64, device="cuda")
b = Beam2D(4e-1, n, 532e-7, init_field=f, complex_bits=64, use_gpu=True)
lib = _lib.load()
plan = b._plan.handle
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
s = torch.empty_like(f); o = torch.empty_like(f)
buf = (C.c_ulonglong * 16)()
for name, fn in (("fft2", lambda: lib.lp2d_fft2(plan, f.data_ptr(), s.data_ptr(), batch, 0, st)),
                 ("stateful", lambda: lib.lp2d_propagate(plan, s.data_ptr(), o.data_ptr(), None, batch, 0.02, st))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lib.lp2d_q4_timing(plan, buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    lib.lp2d_q4_timing(plan, buf, 1)
    v = list(buf)
    print(f"{name}: {e0.elapsed_time(e1) / 5 * 1e3 / batch:.2f} us/beam")
    for kind, o_ in (("row items", 0), ("col items", 8)):
        cnt = max(1, v[o_ + 4])
        w, l, fa, t = (v[o_ + i] / cnt for i in range(4))
        print(f"  {kind}: per warp and item, clock ticks: wait {w:7.0f} | tile->regs {l:7.0f} | factor {fa:7.0f} | transform+stores {t:7.0f} | total {w+l+fa+t:7.0f}   ({cnt} warp-items)")
