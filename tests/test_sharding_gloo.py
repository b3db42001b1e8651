"""CPU, world_size 2 over gloo: the multi-rank plumbing of the ensemble path (contiguous shards, counter-based inputs,
sum of step counters, max of times, final gather).  The per-shard solve is done by the oracle here (no GPU in this
container); on the GPU box the same plumbing wraps deq_solve (bench.py)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from diffeq_b200 import sharding, synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(total, rank, world)
    y0 = synth.lorenz_ics(b, e)
    r = O.ensemble_adaptive(O.ODE45, O.LORENZ, synth.LORENZ_PARAMS, y0, [0.0, 1.0], O.opts(1e-8, 1e-8), nthreads=1)
    counts, tmax = sharding.reduce_counts_and_time(dist, [int(r["accepted"].sum()), int(r["rejected"].sum())], 0.1 * (rank + 1))
    parts = sharding.gather_final(dist, r["yfinal"])
    if rank == 0:
        q.put((counts, tmax, np.concatenate([p.numpy() for p in parts])))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    sys.path.insert(0, ROOT)
    from diffeq_b200 import synth
    from oracle import oracle as O
    total, world = 301, 2   # odd: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    [p.start() for p in procs]
    counts, tmax, final = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ref = O.ensemble_adaptive(O.ODE45, O.LORENZ, synth.LORENZ_PARAMS, synth.lorenz_ics(0, total), [0.0, 1.0], O.opts(1e-8, 1e-8))
    assert counts == [int(ref["accepted"].sum()), int(ref["rejected"].sum())]
    assert abs(tmax - 0.2) < 1e-12                       # max over ranks, not the sum or the mean
    assert np.array_equal(final, ref["yfinal"])          # result of trajectory i does not depend on the partition
