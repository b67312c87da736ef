"""Read partitioning for 1/2/4/8 GPUs: reads are independent, the graph index is
replicated, so the path shards with no exchange step -- every rank aligns its own reads and only
the finished records are gathered on the host.  Work per read grows with its length (K1) and with
length x distance (K3), so reads are dealt by decreasing length in a snake order, which keeps the
ranks' total base counts within one read of each other."""
from __future__ import annotations

import numpy as np


def length_balanced_shards(lengths, world: int):
    """Indices of the reads of every rank: sorted by decreasing length and dealt 0..w-1, w-1..0, ..."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    shards = [[] for _ in range(world)]
    for k, idx in enumerate(order):
        lap, pos = divmod(k, world)
        shards[pos if lap % 2 == 0 else world - 1 - pos].append(int(idx))
    return [np.asarray(sorted(s), dtype=np.int64) for s in shards]


def gather_records(local: dict, rank: int, world: int):
    """Gather {read index: record} dictionaries on rank 0 (host-side result gathering; gloo or nccl group)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    if rank != 0:
        return None
    merged = {}
    for p in parts:
        merged.update(p)
    return merged
