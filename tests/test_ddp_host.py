"""CPU (-m "not gpu"): the data-parallel host logic with world_size 2 over gloo — bucket averaging of
``training.BucketReducer`` and frame sharding — exactly the code the NCCL path runs on the B200 box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bdvs_b200.synthesis import shard_range
from bdvs_b200.training import BucketReducer


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        red = BucketReducer()
        buckets = {name: torch.full((n,), float(rank + 1) * (i + 1)) for i, (name, n) in
                   enumerate([("dd", 1000), ("du", 300), ("ed", 17), ("eu", 5)])}
        for name in ("dd", "du", "ed", "eu"):
            red.bucket_ready(name, buckets[name])
        red.bucket_ready("none", None)
        red.finish()
        ok = all(torch.allclose(buckets[name], torch.full_like(buckets[name], 1.5 * (i + 1)))
                 for i, name in enumerate(("dd", "du", "ed", "eu")))
        ok = ok and red.bytes_per_step == 4 * (1000 + 300 + 17 + 5)
        # frames of a 131-frame job: shards are disjoint, contiguous and cover everything
        lo, hi = shard_range(131, rank, world)
        t = torch.zeros(131)
        t[lo:hi] = 1
        dist.all_reduce(t)
        ok = ok and bool((t == 1).all())
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_bucket_reducer_and_sharding_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_range_balanced():
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
