"""Per-image sharding of an editing sweep across the GPUs of one box (BASELINE config 4).

The reference shards per *config combination*: one OS process per (model, data, edit_cfg, method, edit_method) tuple,
samples serial inside (eval.py:112-133,155-181), so a single-config 700-image PIE-Bench sweep uses one GPU.  Here the
unit is the (image, prompt pair): sample i goes to rank i mod W, each rank walks its samples in lock-step groups of
`cobatch`, and NOTHING is exchanged inside the loops; only the per-sample result records are gathered at the end
(torch.distributed all_gather_object: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Sequence


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Samples owned by `rank`: i = rank, rank+world, ...  Every index in [0, n) is owned by exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n, world))


def group_indices(indices: Sequence[int], cobatch: int) -> List[List[int]]:
    """Lock-step groups of at most `cobatch` samples (the last group of a rank may be ragged)."""
    cobatch = max(1, int(cobatch))
    return [list(indices[i:i + cobatch]) for i in range(0, len(indices), cobatch)]


def gather_results(local: Dict[int, Any], n: int, world: int) -> List[Any]:
    """All-gather the per-sample records and return them ordered by sample index (same list on every rank)."""
    if world == 1:
        merged = dict(local)
    else:
        import torch.distributed as dist
        parts: List[Any] = [None] * world
        dist.all_gather_object(parts, local)
        merged = {}
        for p in parts:
            for k, v in p.items():
                if k in merged:
                    raise RuntimeError(f"sample {k} was processed by two ranks")
                merged[k] = v
    missing = [i for i in range(n) if i not in merged]
    if missing:
        raise RuntimeError(f"samples never processed: {missing[:8]}{'...' if len(missing) > 8 else ''}")
    return [merged[i] for i in range(n)]


def run_sweep(n: int, rank: int, world: int, cobatch: int, run_group: Callable[[List[int]], List[Any]],
              window: int = 1, run_groups: Callable[[List[List[int]]], List[List[Any]]] = None) -> List[Any]:
    """Drive a sweep: `run_group(indices)` edits one lock-step group and returns one record per index.
    With `run_groups` (and `window` > 1) the rank's groups are handed over `window` at a time, so that several groups
    can be in flight on the GPU (batching.run_pipelined); records come back per group, in order."""
    local: Dict[int, Any] = {}
    groups = list(group_indices(shard_indices(n, rank, world), cobatch))
    step = max(1, window) if run_groups is not None else 1
    for w0 in range(0, len(groups), step):
        chunk = groups[w0:w0 + step]
        recs_per_group = run_groups(chunk) if run_groups is not None else [run_group(chunk[0])]
        if len(recs_per_group) != len(chunk):
            raise RuntimeError("run_groups must return one record list per group")
        for grp, recs in zip(chunk, recs_per_group):
            if len(recs) != len(grp):
                raise RuntimeError("run_group must return one record per sample")
            local.update(dict(zip(grp, recs)))
    return gather_results(local, n, world)
