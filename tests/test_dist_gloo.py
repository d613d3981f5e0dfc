"""CPU, world_size 2, gloo: the N>1 host logic of the ensemble path — block sharding and the
optional gather of per-beam results.  The data path itself has no collective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightprop2d_b200.ensemble import gather_to_rank0, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        # stand-in for "propagate my block": a per-beam map that depends only on the global beam index
        beams = torch.arange(lo, hi, dtype=torch.float64)
        local = beams[:, None, None] * torch.ones(1, 4, 4, dtype=torch.float64) + 0.5
        full = gather_to_rank0(local, n_total)
        # per-beam scalars too
        sc = gather_to_rank0(beams * 2, n_total)
        if rank == 0:
            q.put((full.numpy(), sc.numpy()))
        else:
            assert full is None and sc is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    world, n_total = 2, 7        # ragged: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, sc = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert full.shape == (n_total, 4, 4)
    assert np.array_equal(full[:, 0, 0], np.arange(n_total) + 0.5)
    assert np.array_equal(sc, np.arange(n_total) * 2.0)


def test_single_process_gather_is_identity():
    t = torch.arange(6.).reshape(3, 2)
    assert gather_to_rank0(t, 3) is t
