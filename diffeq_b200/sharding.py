"""Ensemble sharding over GPUs / ranks: contiguous equal blocks of the trajectory index range, one per
rank; no data-path collective — every trajectory is an independent OdeProblem (problem.rs:32-47).  The only
communication is the final reduction of counters / timings (and an optional gather of final states)."""


def shard_range(ntraj_total, rank, world):
    """[begin, end) of rank's contiguous block; the first (ntraj_total % world) ranks get one extra trajectory."""
    base, extra = divmod(int(ntraj_total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def cyclic_indices(ntraj_total, rank, world, chunk):
    """Block-cyclic alternative for parameter-sorted sweeps: chunk j of `chunk` consecutive trajectories
    belongs to rank j % world.  Returns the global indices of rank's trajectories, ascending.  `deq_solve_host` does the
    same inside the library when `deq_options.shard_chunk` > 0."""
    import numpy as np
    ntraj_total, chunk = int(ntraj_total), int(chunk)
    starts = np.arange(rank * chunk, ntraj_total, world * chunk, dtype=np.int64)
    if len(starts) == 0:
        return np.empty(0, dtype=np.int64)
    return np.concatenate([np.arange(s, min(s + chunk, ntraj_total), dtype=np.int64) for s in starts])


def weak_range(ntraj_per_rank, rank):
    """Weak scaling: every rank integrates the same number of trajectories, indices are globally unique."""
    return rank * int(ntraj_per_rank), (rank + 1) * int(ntraj_per_rank)


def reduce_counts_and_time(dist, counts, seconds, device=None):
    """Sum integer counters and take the max time over ranks with torch.distributed (any backend).
    Returns (list_of_summed_counts, max_seconds)."""
    import torch
    c = torch.tensor([float(x) for x in counts], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [int(round(x)) for x in c.tolist()], float(t.item())


def gather_final(dist, local_final, device=None):
    """Final gather of the sharded final states (the only inter-GPU traffic the path needs): list of per-rank arrays."""
    import torch
    x = torch.as_tensor(local_final, device=device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [x]
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([x.shape[0]], dtype=torch.int64, device=device))
    outs = [torch.empty((int(s.item()),) + tuple(x.shape[1:]), dtype=x.dtype, device=device) for s in sizes]
    dist.all_gather(outs, x) if len({int(s.item()) for s in sizes}) == 1 else [
        dist.broadcast(outs[r] if r != dist.get_rank() else outs[r].copy_(x), src=r) for r in range(dist.get_world_size())]
    return outs
