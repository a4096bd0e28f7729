"""N > 1 path on CPU: two gloo ranks shard utterances round-robin, time a step as the max over ranks and
reassemble results in order (the same helpers bench.py uses under torchrun/NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svcc23_fastsvc_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_convert(i):
    # stands in for one utterance's conversion: deterministic, depends only on the utterance index
    g = torch.Generator().manual_seed(1000 + i)
    return torch.randn(8, generator=g)


def _worker(rank, world, port, n_utts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_utterances(n_utts, rank, world)
        outs = [_fake_convert(i) for i in mine]
        sharding.barrier()
        ms = sharding.max_over_ranks(10.0 + 5.0 * rank)          # rank 1 is the slow one
        full = sharding.gather_in_order(outs, mine, n_utts)
        if rank == 0:
            q.put((mine, ms, [t.tolist() for t in full]))
        else:
            q.put((mine, ms, None))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, n_utts = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_utts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(r[0] for r in res)
    assert shards == [[0, 2, 4, 6], [1, 3, 5]]                     # disjoint, complete, round-robin
    assert all(abs(r[1] - 15.0) < 1e-12 for r in res)              # step time = max over ranks
    full = next(r[2] for r in res if r[2] is not None)
    assert full == [_fake_convert(i).tolist() for i in range(n_utts)]
    assert sharding.aggregate_throughput(32 * 16000, world, 2.0) == 2 * 32 * 16000 / 2e-3


def test_single_process_fallbacks():
    assert sharding.shard_utterances(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.gather_in_order(["a", "b"], [1, 0], 2) == ["b", "a"]
