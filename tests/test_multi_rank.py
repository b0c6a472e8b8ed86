"""CPU, world_size 2 over gloo: the N>1 logic of bench.py / the worker -- request sharding without overlap and the
timing aggregation (sum of tokens over the max of the per-rank times).  The data path itself has no collective."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    from gridllm_b200 import multirank as M
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = M.shard_round_robin(7, rank, world)
        seeds = [s for step in range(3) for s in M.request_seeds(rank, step, 2)]
        dist.barrier()
        # rank r "generates" 128 tokens per request in (1 + r) seconds of device time
        agg = M.aggregate_throughput(tokens=128.0 * len(mine), device_seconds=1.0 + rank, wall_seconds=1.5 + rank, dist=dist)
        dist.barrier()
        q.put((rank, mine, seeds, agg))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_shard_and_aggregate():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    (r0, mine0, seeds0, agg0), (r1, mine1, seeds1, agg1) = out
    assert sorted(mine0 + mine1) == list(range(7)) and not set(mine0) & set(mine1)
    assert not set(seeds0) & set(seeds1) and len(seeds0) == len(seeds1) == 6
    for agg in (agg0, agg1):                       # every rank sees the same whole-job numbers
        assert agg["tokens"] == 128.0 * 7
        assert agg["device_s"] == 2.0 and agg["wall_s"] == 2.5
        assert abs(agg["value"] - 128.0 * 7 / 2.0) < 1e-9 and abs(agg["e2e"] - 128.0 * 7 / 2.5) < 1e-9


def test_single_process_is_identity():
    from gridllm_b200 import multirank as M
    assert M.reduce_max([1.0, 2.0], None) == [1.0, 2.0]
    assert M.aggregate_throughput(256.0, 2.0, 4.0, None) == {"tokens": 256.0, "device_s": 2.0, "wall_s": 4.0, "value": 128.0, "e2e": 64.0}
    assert M.shard_round_robin(5, 1, 2) == [1, 3]
