"""Multi-GPU plumbing: one process per GPU (torch.distributed), set A sharded, set B replicated,
partial matrices summed with one allreduce (NCCL over NVLink on GPUs; gloo in the CPU tests).

The path shards naturally — set-A sequences are independent units against a read-only set-B
table, which is how the reference threads it (src/overlap.cc:421-448) — so the only collective is
the sum of the R1 x R2 partial matrices (the reference's serial merge, overlap.cc:510-527).
Existence mode shards matrix ROWS, so it needs a gather, not a reduce; pairs stay per rank."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from .seqset import SeqSet


def probe_weights(lengths: np.ndarray, sigma: int, differences: int, indels: bool) -> np.ndarray:
    """Upper bound of the variants per sequence (runs are not inspected): the balancing weight."""
    L = lengths.astype(np.float64)
    w = np.ones_like(L)
    if differences >= 1:
        w += (sigma - 1) * L
        if indels:
            w += L + sigma * (L + 1) - L
    if differences == 2:
        w += (sigma - 1) ** 2 * L * (L - 1) / 2
    if differences > 2:
        w = np.ones_like(L)
    return w


def plan_shards(lengths: np.ndarray, world: int, sigma: int = 20, differences: int = 1,
                indels: bool = False) -> List[Tuple[int, int]]:
    """Contiguous (first, count) ranges of set A, one per rank, balanced by expected probes
    (cost per seed ~ L for d=1, ~ L^2 for d=2), covering every sequence exactly once."""
    n = int(lengths.shape[0])
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, 0)] * (world - 1)
    pre = np.concatenate([[0.0], np.cumsum(probe_weights(lengths, sigma, differences, indels))])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(pre, pre[-1] * r / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n)).tolist()
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


def allreduce_matrix(matrix, group=None):
    """Sum a partial matrix over all ranks, in place.  `matrix` is a torch tensor (CUDA for NCCL,
    CPU for gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(matrix, op=dist.ReduceOp.SUM, group=group)
    return matrix


def gather_rows(rows, counts: List[int], group=None):
    """Existence mode: concatenate the per-rank row blocks (rank order = shard order)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    cols = rows.shape[1]
    most = max(counts)   # all_gather wants equal shapes: pad to the largest shard, trim after
    mine = torch.zeros((most, cols), dtype=rows.dtype, device=rows.device)
    mine[: rows.shape[0]] = rows
    bufs = [torch.empty((most, cols), dtype=rows.dtype, device=rows.device) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    return torch.cat([bufs[r][: counts[r]] for r in range(world)], dim=0)


def sharded_overlap(a: SeqSet, b: Optional[SeqSet], compute: Callable, rank: int, world: int,
                    differences: int = 1, indels: bool = False, existence: bool = False,
                    device="cpu", group=None):
    """Rank-local driver: run `compute(a_shard, b) -> (rows x cols numpy matrix)` on this rank's
    shard of set A and combine.  In matrix mode every rank returns the full summed matrix; in
    existence mode the full row-concatenated matrix.  `compute` is the GPU engine in production
    (see bench.py) and anything with the same signature in tests."""
    import torch
    shards = plan_shards(a.lengths, world, a.sigma, differences, indels)
    first, count = shards[rank]
    part = compute(a.slice(first, count), b if b is not None else a)
    t = torch.as_tensor(np.ascontiguousarray(part), device=device)
    if existence:
        return gather_rows(t, [c for _, c in shards], group).cpu().numpy()
    return allreduce_matrix(t, group).cpu().numpy()
