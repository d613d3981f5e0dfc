"""The reference's own timing script (tests/gaussbeam_performance.py:15-44) on this backend: Beam2D construction and ONE
propagate(10) at 2048 / 4096 / 8192 points, complex128 (the reference's default), host and device generators."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, rectangle_hole, device_profiles as dp  # noqa: E402

area_size, wl0 = 4e-1, 532e-7
for gen, what in ((rectangle_hole, "host generator"), (dp.rectangle_hole, "device generator")):
    for n in (2048, 4096, 8192):
        for rep in range(2):                      # second round: plans, twiddles and scratch exist
            torch.cuda.synchronize()
            t = time.time()
            beam = Beam2D(area_size, n, wl0, init_field_gen=gen, init_gen_args=(2e-1, 1e-1), use_gpu=1)
            torch.cuda.synchronize()
            t_c = time.time() - t
            t = time.time()
            beam.propagate(10)
            torch.cuda.synchronize()
            t_p = time.time() - t
            t = time.time()
            beam.propagate(10)
            torch.cuda.synchronize()
            t_p2 = time.time() - t
            print(f"{what:17s} {n:5d} points round {rep}: construction {t_c*1e3:9.2f} ms, first propagate {t_p*1e3:8.3f} ms, second {t_p2*1e3:8.3f} ms", flush=True)
            del beam
            torch.cuda.empty_cache()
