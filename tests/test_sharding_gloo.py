"""N > 1 host logic on CPU: world_size-2 gloo ranks shard a batch with the product's shard_range(),
solve their shard with a CPU stand-in for the device (the oracle, as SURVEY.md 4.2 T4 prescribes --
the sharding and the statistics gather are what is under test here, not the kernels) and all-reduce
the four statistics exactly as bench.py does over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    from oracle import cpu
    pkg = graft.load_pkg()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob, opts = pkg.problems.cfg2_cw_batch(batch=21, N=8, seed=3)
    opts = dict(opts, max_iter=300)
    b, c = pkg.dist.shard_range(21, world, rank)
    sub = dict(prob, s0=prob["s0"][b:b + c])
    x, z, u, h = cpu.solve(sub, opts, nthreads=1)
    stats, sec = pkg.dist.gather_stats(h["stats"], 0.1 * (rank + 1))
    q.put((rank, b, c, stats, sec, h["iters"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly(pkg):
    for batch in (1, 7, 21, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [pkg.dist.shard_range(batch, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == batch
            pos = 0
            for b, c in spans:
                assert c >= 0 and (c == 0 or b == pos)
                pos += c


@pytest.mark.timeout(180)
def test_two_gloo_ranks_reproduce_the_single_process_statistics(pkg, cpu_oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    prob, opts = pkg.problems.cfg2_cw_batch(batch=21, N=8, seed=3)
    x, z, u, h = cpu_oracle.solve(prob, dict(opts, max_iter=300), nthreads=1)
    iters = sum((o[5] for o in out), [])
    assert iters == h["iters"].tolist()                      # contiguous shards, original order
    for o in out:
        assert o[3] == [int(v) for v in h["stats"]]           # every rank holds the global statistics
        assert abs(o[4] - 0.2) < 1e-12                        # time = max over ranks
