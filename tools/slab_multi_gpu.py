"""Slab-decomposed grid on several GPUs (torchrun, one process per GPU, NCCL all-to-all).

    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/slab_multi_gpu.py [n_check] [n_time]

1. parity: an n_check^2 grid against the numpy oracle (complex64 1e-5, complex128 1e-12), and the
   n_time^2 grid against the closed form for a sum of plane waves (tools/planewave_check.py);
2. timing: repeated propagate steps on an n_time^2 complex64 grid (spectrum resident, one exchange
   per step), CUDA events, max over ranks; prints one JSON line on rank 0.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200.slab import SlabBeam2D, slab_rows  # noqa: E402


def main():
    n_check = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {"world": world}
    from oracle import beam_oracle as orc
    for bits, tol in ((64, 1e-5), (128, 1e-12)):
        n = n_check
        area, wl = 200e-4 * n / 256, 632e-7
        rng = np.random.default_rng(11)
        ct = np.complex64 if bits == 64 else np.complex128
        full = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(ct)
        r0, r1 = slab_rows(n, rank, world)
        b = SlabBeam2D(area, n, wl, init_field=full[r0:r1].copy(), complex_bits=bits)
        b.propagate(100e-4)
        b.propagate(-37e-4)
        fld = b.gather_field().cpu().numpy()
        spec = b.gather_spectrum().cpu().numpy()
        if rank == 0:
            o = orc.OracleBeam(area, n, wl, init_field=full.copy(), complex_bits=bits)
            o.propagate(100e-4); o.propagate(-37e-4)
            ef = float(np.linalg.norm(fld - o.field) / np.linalg.norm(o.field))
            es = float(np.linalg.norm(spec - o.spectrum) / np.linalg.norm(o.spectrum))
            res[f"c{bits}_field_rel_l2"], res[f"c{bits}_spectrum_rel_l2"] = ef, es
            assert ef <= tol and es <= tol, (bits, ef, es)
        del b
    # full-size parity through the plane-wave closed form (the numpy oracle cannot hold the big grid)
    from tools import planewave_check as pw
    res["planewave_n"] = n_time
    res["planewave_c64_rel_l2"] = pw.check(n_time, 64, 0.4 * n_time / 1024, 532e-7, 0.02, rank, world)
    torch.cuda.empty_cache()
    if rank == 0:
        assert max(res["planewave_c64_rel_l2"].values()) <= 1e-5, res
    # timing
    n = n_time
    r0, r1 = slab_rows(n, rank, world)
    g = torch.Generator(device="cuda"); g.manual_seed(7 + rank)
    f = torch.polar(torch.rand(r1 - r0, n, generator=g, device="cuda"), torch.rand(r1 - r0, n, generator=g, device="cuda") * 6.2831853).to(torch.complex64)
    b = SlabBeam2D(0.4 * n / 1024, n, 532e-7, init_field=f, complex_bits=64)
    for _ in range(3):
        b.propagate(0.02)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        b.propagate(0.02)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        res.update({"n_time": n, "ms_per_propagate": float(ms.item()),
                    "hbm_bytes_alg": 3 * n * n * 8, "alg_GBps_per_gpu": 3 * n * n * 8 / world / (float(ms.item()) * 1e-3) / 1e9,
                    "nvlink_bytes_sent_per_rank": n * n * 8 * (world - 1) // (world * world)})
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
