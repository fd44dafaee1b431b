"""world_size-2 gloo test of the multi-GPU host logic (SAI sharding, gather, max-over-ranks timing)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lfbm5d_b200 import dist as D
    asize = 9
    full = np.arange(asize * 6, dtype=np.float32).reshape(asize, 2, 3)
    lo, hi = D.shard_sais(asize, world, rank)
    mask = D.shard_mask(np.ones(asize), world, rank)
    assert mask.sum() == hi - lo and mask[lo:hi].all()
    out = D.gather_shards(full[lo:hi] * 2.0, asize, dist)
    ok = bool(np.array_equal(out.numpy(), full * 2.0))
    t = D.max_over_ranks(1.0 + rank, dist)
    ret[rank] = (ok, t, lo, hi)
    dist.destroy_process_group()


def test_shard_gather_max_over_ranks():
    from lfbm5d_b200 import dist as D
    for asize in (1, 9, 289):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = D.shard_sais(asize, world, r)
                cover += list(range(lo, hi))
            assert cover == list(range(asize))
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29613, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0] and ret[0][1] == 2.0 and ret[1][1] == 2.0
    assert (ret[0][2], ret[0][3], ret[1][2], ret[1][3]) == (0, 5, 5, 9)
