import sys, time, os
sys.path.insert(0, '/root/repo')
import torch
from lightprop2d_b200.ensemble import EnsemblePropagator
n, B = 1024, 64
prop = EnsemblePropagator(0.4, n, 532e-7, complex_bits=64)
host_in = torch.empty(B, n, n, dtype=torch.complex64, pin_memory=True); host_in.copy_(torch.randn(B, n, n, dtype=torch.complex64))
host_out = torch.empty(B, n, n, dtype=torch.float32, pin_memory=True)
for chunk in (4, 8, 16, 32, 64):
    for _ in range(2): prop.intensity_from_host(host_in, 0.02, chunk=chunk, out=host_out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): prop.intensity_from_host(host_in, 0.02, chunk=chunk, out=host_out)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"chunk {chunk:3d}: {dt*1e3:7.2f} ms per {B} beams -> {B/dt:8.0f} beams/s")
# raw copies
d = torch.empty(B, n, n, dtype=torch.complex64, device='cuda'); o = torch.empty(B, n, n, dtype=torch.float32, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(host_in, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s", 5 * host_in.numel() * 8 / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter()
for _ in range(5): host_out.copy_(o, non_blocking=True)
torch.cuda.synchronize(); print("D2H GB/s", 5 * host_out.numel() * 4 / (time.perf_counter() - t0) / 1e9)
