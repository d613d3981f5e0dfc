"""L2 bandwidth probe: repeated copy between two small (L2-resident) buffers vs large ones."""
import torch


def t(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for mb in (4, 8, 16, 32, 48, 64, 128, 512, 2048):
    n = mb << 20
    x = torch.empty(n, dtype=torch.uint8, device="cuda")
    y = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = t(lambda: y.copy_(x))
    print(f"copy {mb:5d} MB -> {mb:5d} MB: {ms*1e3:8.1f} us  {2*n/ms/1e6:8.1f} GB/s (read+write)")
    xs = x.view(torch.float32)
    ms = t(lambda: xs.sum())
    print(f"sum  {mb:5d} MB           : {ms*1e3:8.1f} us  {n/ms/1e6:8.1f} GB/s (read)")
