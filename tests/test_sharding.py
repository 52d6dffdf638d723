"""Frame-set sharding: pure host logic + a world_size-2 gloo run on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import panob200

sh = panob200.sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, e = sh.shard_range(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [sh.shard_range(n, r, world)[1] - sh.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_round_robin_and_strips():
    allv = sorted(sum((sh.shard_round_robin(10, r, 4) for r in range(4)), []))
    assert allv == list(range(10))
    strips = sh.strip_columns(18176, 7, 8)      # BASELINE config 4: 142 units of 128
    assert strips[0][0] == 0 and strips[-1][1] == 18176
    assert all(a % 128 == 0 and b % 128 == 0 and b > a for a, b in strips)
    assert all(strips[i][1] == strips[i + 1][0] for i in range(7))
    assert sorted((b - a) // 128 for a, b in strips) == [17, 17, 18, 18, 18, 18, 18, 18]
    with pytest.raises(ValueError):
        sh.strip_columns(100, 3, 2)


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sh.shard_range(n, rank, world)
    # each rank "processes" its frame-sets: mark them, then check global coverage with no overlap
    mine = torch.zeros(n, dtype=torch.int32)
    mine[b:e] = 1
    dist.all_reduce(mine)
    assert bool((mine == 1).all())
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 37), nprocs=2, join=True)
