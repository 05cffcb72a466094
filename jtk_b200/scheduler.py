"""Host-side scheduler: shard independent chunks over the GPUs of one box, gather per-chunk results on the host.

The reference fans chunks out over rayon threads (`pileups.into_par_iter()`, local_clustering/mod.rs:64-72); chunks share
no state, so here every rank (one process per GPU, torch.distributed) takes a disjoint subset of the chunks, scores /
clusters them on its own device, and rank 0 gathers the small per-chunk results (consensus, score, k, assignments,
posteriors, ops).  There is no collective on the data path and no NCCL traffic for the tables: the only exchange is this
host gather.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def partition_chunks(weights: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of chunk indices to `world` ranks (deterministic: ties broken by index).
    weights[c] ~ cost of chunk c (reads x summed length is the cell-update count up to the band width)."""
    order = sorted(range(len(weights)), key=lambda c: (-float(weights[c]), c))
    load = [0.0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for c in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(c)
        load[r] += float(weights[c])
    for p in parts:
        p.sort()
    return parts


def chunk_weight(n_reads: int, chunk_len: int, mean_read_len: float, radius: int) -> float:
    """In-band cell updates of one modification-table pass over a pile-up: 2 * n * (Lt+Lr+1) * (2r+1)."""
    return 2.0 * n_reads * (chunk_len + mean_read_len + 1.0) * (2 * radius + 1)


def dist_info():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched under torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def gather_to_rank0(local: Dict[int, object], group=None) -> Optional[Dict[int, object]]:
    """Host gather of per-chunk results keyed by chunk id.  Returns the merged dict on rank 0, None elsewhere.
    Works on any backend (gloo on CPU in the tests, nccl under bench / production: objects are pickled by torch)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    if out is None:
        return None
    merged: Dict[int, object] = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"chunk {k} was processed by two ranks")
            merged[k] = v
    return merged


def run_sharded(chunk_ids: Sequence[int], weights: Sequence[float], process: Callable[[List[int]], Dict[int, object]],
                rank: int, world: int, group=None) -> Optional[Dict[int, object]]:
    """Partition, process this rank's chunks with `process(list of chunk ids) -> {chunk id: result}`, gather on rank 0."""
    parts = partition_chunks(weights, world)
    mine = [chunk_ids[c] for c in parts[rank]]
    local = process(mine) if mine else {}
    missing = set(mine) - set(local)
    if missing:
        raise RuntimeError(f"rank {rank}: no result for chunks {sorted(missing)[:5]}")
    # world == 1: nothing to gather, even inside a process that has a torch.distributed group for other work
    merged = dict(local) if world <= 1 else gather_to_rank0(local, group=group)
    if merged is not None and set(merged) != set(chunk_ids):
        raise RuntimeError("gather lost or duplicated chunks")
    return merged
