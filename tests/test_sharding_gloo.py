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


def _worker_cyclic(rank, world, port, total, chunk, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from diffeq_b200 import sharding, synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y0, mu = synth.vdp_sweep(0, total, total)              # parameter-sorted sweep: the case block-cyclic sharding is for
    idx = sharding.cyclic_indices(total, rank, world, chunk)
    r = O.ensemble_adaptive(O.ODE45FE, O.VANDERPOL, mu[idx], y0[idx], [0.0, 5.0], O.opts(1e-6, 1e-8), nthreads=1)
    counts, _ = sharding.reduce_counts_and_time(dist, [int(r["accepted"].sum()), int(r["rejected"].sum())], 0.0)
    parts = sharding.gather_final(dist, r["yfinal"])
    if rank == 0:
        final = np.empty((total, 2))
        for rr, p in enumerate(parts):
            final[sharding.cyclic_indices(total, rr, world, chunk)] = p.numpy()
        q.put((counts, final))
    dist.destroy_process_group()


def test_two_rank_block_cyclic_sharding_matches_single_process():
    sys.path.insert(0, ROOT)
    from diffeq_b200 import synth
    from oracle import oracle as O
    total, world, chunk = 333, 2, 50   # ragged last chunk, unequal shard sizes (183 / 150)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_cyclic, args=(r, world, port, total, chunk, q)) for r in range(world)]
    [p.start() for p in procs]
    counts, final = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    y0, mu = synth.vdp_sweep(0, total, total)
    ref = O.ensemble_adaptive(O.ODE45FE, O.VANDERPOL, mu, y0, [0.0, 5.0], O.opts(1e-6, 1e-8))
    assert counts == [int(ref["accepted"].sum()), int(ref["rejected"].sum())]
    assert np.array_equal(final, ref["yfinal"])
