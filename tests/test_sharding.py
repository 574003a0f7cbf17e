"""Host logic of the multi-GPU path, exercised with world_size-2 gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imageflow_b200 import sharding


def test_contiguous_blocks_cover_batch_exactly():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                blk = sharding.shard_contiguous(n, world, r)
                seen += list(blk)
                assert abs(len(blk) - n / world) < 1
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_contiguous(4, 2, 2)


def test_lpt_balances_mixed_workload_and_keeps_chains_whole():
    import random
    rng = random.Random(7)
    costs = []
    for _ in range(2000):                       # config-5 shaped: long edge log-uniform 256..7680
        le = int(256 * (30 ** rng.random()))
        ar = rng.choice([(1, 1), (4, 3), (3, 2), (16, 9)])
        w, h = le, max(1, le * ar[1] // ar[0])
        costs.append(sum(a[0] * a[1] for a, _ in sharding.export_4_sizes_chain(w, h)) or 1)
    bins = sharding.shard_lpt(costs, 8)
    assert sorted(i for b in bins for i in b) == list(range(2000))
    assert sharding.lpt_imbalance(costs, bins) < 1.01          # >= 7.9x of ideal 8x scaling from balance alone


def test_export_4_sizes_chain_semantics():
    ch = sharding.export_4_sizes_chain(4000, 3000)
    assert ch == [((4000, 3000), (1600, 1200)), ((1600, 1200), (1200, 900)), ((1200, 900), (400, 300)), ((1600, 1200), (800, 600))]
    assert sharding.export_4_sizes_chain(300, 200) == []       # never upscales: every node deletes itself
    assert sharding.export_4_sizes_chain(1000, 500) == [((1000, 500), (400, 200)), ((1000, 500), (800, 400))]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blk = sharding.shard_contiguous(1024, world, rank)
    units = len(blk) * 3840 * 2160
    ms = 10.0 + rank                                  # rank 1 is slower: the max must win
    tot, mx = sharding.aggregate(units, ms)
    q.put((rank, list(blk)[:2], tot, mx))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregate():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1] and res[1][1] == [512, 513]
    for _, _, tot, mx in res:
        assert tot == 1024 * 3840 * 2160 and mx == 11.0
