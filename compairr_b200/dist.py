"""Multi-GPU driver: one rank per GPU, each with its own Engine joined in an NCCL communicator that
lives INSIDE the library (csrc/comm.cu).  The path shards by set-A sequences — they are independent
units against a read-only set-B table, which is how the reference threads it
(src/overlap.cc:421-448):

  set B   every rank uploads 1/world of it over its own PCIe link (shard_range), the ranks
          all-gather over NVLink and each builds the whole table          Engine.set_b_sharded
  set A   contiguous ranges balanced by expected probes (plan_shards)     Engine.run_a
  result  -m: sum of the partial matrices (the reference's merge, overlap.cc:510-527)
                                                                          Engine.allreduce_matrix
          -x: rows follow the A shards -> concatenated on the host; pairs stay per rank

`overlap_rank` is that sequence for one rank.  It talks to the engine through the five methods named
above, so the CPU test-suite can drive the same code with world_size-2 gloo processes and a host
stand-in for the engine (tests/test_dist_gloo.py); bench.py and the multi-GPU tests pass the real
Engine.  torch.distributed is only the host channel for the 128-byte communicator id."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .seqset import NarrowSet, SeqSet


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of set B that `rank` uploads: equal shards of ceil(n_total / world), the last
    short or empty — the layout cb_set_b_sharded's in-place all-gather needs (= cb_shard_range)."""
    per = -(-n_total // world) if world > 0 else n_total
    first = min(n_total, per * max(rank, 0))
    return first, min(per, n_total - first)


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


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 of the torch.distributed group makes the communicator id (make_id() -> 128 bytes,
    Engine.comm_unique_id), everybody gets it.  Works over gloo (CPU tensor) and nccl (CUDA tensor)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return make_id()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    if dist.get_rank(group) == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).to(dev)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def gather_rows(rows: np.ndarray, group=None) -> np.ndarray:
    """Existence mode: concatenate the per-rank row blocks in rank order (= shard order)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, rows, group=group)
    return np.concatenate(parts, axis=0)


def overlap_rank(eng, a: SeqSet, b: SeqSet, rank: int, world: int, differences: int = 1, indels: bool = False,
                 existence: bool = False, group=None) -> Optional[np.ndarray]:
    """This rank's part of `compairr -m/-x a b` on `world` GPUs; returns the complete matrix (on
    every rank).  `eng` has joined the communicator already (Engine.comm_init_rank)."""
    first, count = shard_range(b.n, rank, world)
    shard = NarrowSet.from_seqset(b.slice(first, count))
    shard.n_reps, shard.index_base = b.n_reps, b.index_base     # of the WHOLE set, the same on every rank
    eng.set_b_sharded(shard, b.n)
    f, c = plan_shards(a.lengths, world, a.sigma, differences, indels)[rank]
    eng.run_a(a.slice(f, c))        # the slice carries index_base = f: pairs and -x rows are global
    if existence:
        return gather_rows(eng.matrix(), group)
    eng.allreduce_matrix()
    return eng.matrix()
