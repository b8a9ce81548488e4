"""Sharding of candidate windows over the GPUs of one box.

Windows are independent from ``Decision`` downward, so there is no collective on the data
path: every rank (one process per GPU) takes a cost-balanced subset, computes its records and
the parent gathers them on the host.  The only communication is that host-side gather."""
from __future__ import annotations

from typing import List, Sequence



def lpt_shards(costs: Sequence[float], n_shards: int) -> List[List[int]]:
    """Longest-processing-time-first assignment; deterministic (ties by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * n_shards
    shards: List[List[int]] = [[] for _ in range(n_shards)]
    for i in order:
        s = min(range(n_shards), key=lambda k: (load[k], k))
        shards[s].append(i)
        load[s] += float(costs[i])
    return [sorted(s) for s in shards]


def my_shard(costs: Sequence[float], rank: int, world_size: int) -> List[int]:
    return lpt_shards(costs, world_size)[rank]


def gather_records(local_indices: Sequence[int], local_records: Sequence[list], n_total: int, group=None):
    """Host-side gather to rank 0 (returns the full list there, None elsewhere).  Uses
    ``torch.distributed.gather_object`` on a CPU (gloo) group; with a single process it is the
    identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = [None] * n_total
        for i, r in zip(local_indices, local_records):
            out[i] = r
        return out
    rank = dist.get_rank(group)
    payload = (list(local_indices), list(local_records))
    buf = [None] * dist.get_world_size(group) if rank == 0 else None
    dist.gather_object(payload, buf, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * n_total
    for idx, recs in buf:
        for i, r in zip(idx, recs):
            out[i] = r
    return out


def host_group():
    """A gloo group for host-side object gathers when the default group is NCCL."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_backend() == "gloo":
        return None
    return dist.new_group(backend="gloo")
