"""Replica sharding across the GPUs of one box: saddle searches of a batch are independent
between the reference's `taskwait`s (src/saddle/find.cpp:697-732), so global replica r runs on rank r mod G,
every rank holds the shared immutable inputs, and only the finished records (exit code, saddle position,
axis, work counters) travel back to rank 0, which merges them in replica order exactly as Master does after its
taskwait (find.cpp:737-752). There is NO collective on the data path; the gather below runs once per batch over
host-side records, on whatever backend the process group has (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def replica_ids(total: int, world: int, rank: int) -> np.ndarray:
    """Global ids of the replicas rank `rank` owns: r mod world == rank (round-robin keeps long and short
    searches, which correlate with the seed chain, evenly spread)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return np.arange(rank, total, world, dtype=np.int64)


def shard(arrays, total: int, world: int, rank: int):
    """Slices of [total, ...] arrays for this rank (same order as replica_ids)."""
    ids = replica_ids(total, world, rank)
    return ids, tuple(np.ascontiguousarray(np.asarray(a)[ids]) for a in arrays)


def merge_records(ids: np.ndarray, records: dict, total: int, group=None):
    """Gather per-rank records to every rank and re-order them by global replica id.

    `records` maps a name to an array whose first axis matches `ids`. Returns {name: [total, ...] array}.
    With no initialised process group (single GPU) it only re-orders."""
    import torch.distributed as dist

    payload = (np.asarray(ids, dtype=np.int64), {k: np.asarray(v) for k, v in records.items()})
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, payload, group=group)
    else:
        parts = [payload]
    out = {}
    seen = np.zeros(total, dtype=bool)
    for pid, rec in parts:
        if seen[pid].any():
            raise ValueError("a replica was reported by two ranks")
        seen[pid] = True
        for k, v in rec.items():
            if k not in out:
                out[k] = np.zeros((total,) + v.shape[1:], dtype=v.dtype)
            out[k][pid] = v
    if not seen.all():
        raise ValueError(f"replicas {np.nonzero(~seen)[0].tolist()} were not reported by any rank")
    return out
