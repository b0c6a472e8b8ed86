"""Request sharding and result aggregation across one-process-per-GPU ranks.

The hot path has no exchange step: requests are independent, every rank holds a full weight replica and serves
its own requests (what the reference scheduler's least-loaded worker selection does across worker ids,
server/src/services/JobScheduler.ts:317-360).  The only collectives are for TIMING: a barrier on both sides of the
timed region and a max-reduce of the per-rank device times.  Works on `nccl` (GPU tensors) and `gloo` (CPU tensors,
used by the CPU tests with world_size 2)."""
from __future__ import annotations

from typing import Sequence


def request_seeds(rank: int, step: int, requests_per_step: int = 1) -> list[int]:
    """Deterministic, disjoint request ids per rank: rank r's i-th request of a step.  1000 requests per rank."""
    base = rank * 1000 + step * requests_per_step
    return [base + i for i in range(requests_per_step)]


def shard_round_robin(n_requests: int, rank: int, world: int) -> list[int]:
    """Static stand-in for the scheduler when all workers are equally loaded: request i -> rank i % world."""
    return [i for i in range(n_requests) if i % world == rank]


def _tensor(values: Sequence[float], dist):
    import torch
    dev = "cuda" if (dist is not None and dist.get_backend() == "nccl") else "cpu"
    return torch.tensor(list(values), dtype=torch.float64, device=dev)


def reduce_max(values: Sequence[float], dist) -> list[float]:
    """max over ranks of every entry (device / wall times: the job is as slow as its slowest rank)"""
    if dist is None:
        return list(values)
    t = _tensor(values, dist)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()


def reduce_sum(values: Sequence[float], dist) -> list[float]:
    """sum over ranks of every entry (tokens generated, requests served)"""
    if dist is None:
        return list(values)
    t = _tensor(values, dist)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def aggregate_throughput(tokens: float, device_seconds: float, wall_seconds: float, dist) -> dict:
    """whole-job tokens/s: all ranks' tokens over the slowest rank's time"""
    tot = reduce_sum([tokens], dist)[0]
    dev, wall = reduce_max([device_seconds, wall_seconds], dist)
    return {"tokens": tot, "device_s": dev, "wall_s": wall, "value": tot / dev, "e2e": tot / wall}
