"""Batch sharding helpers: problems are independent, so a batch is cut into
contiguous ranges, one per GPU / rank, and NOTHING crosses GPUs while iterating.  The only exchange is
the final gather of four statistics (converged count, sum and max of iterations, refactor count) --
an NCCL all-reduce over NVLink in the one-process-per-GPU deployment, gloo in the CPU tests."""
from __future__ import annotations


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous range [begin, begin+count) of `rank`; same arithmetic as shard_range() in
    csrc/admm_b200.cu (the in-process multi-GPU handle)."""
    per = (batch + world - 1) // world
    begin = min(batch, per * rank)
    end = min(batch, per * (rank + 1))
    return begin, end - begin


def gather_stats(stats, seconds: float, device=None):
    """All-reduce [converged, sum_iters, max_iters, refactors] (sum, sum, max, sum) and the per-rank
    time (max).  Works on any initialised torch.distributed backend; identity when uninitialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(s) for s in stats], float(seconds)
    t = torch.tensor([int(s) for s in stats], dtype=torch.int64, device=device)
    mx = t[2:3].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[2] = mx[0]
    sec = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    return [int(v) for v in t.tolist()], float(sec.item())
