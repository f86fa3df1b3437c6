"""The N>1 path on CPU: world_size-2 (and 3) gloo processes run the SAME per-rank driver as the GPU
ranks do (compairr_b200.dist.overlap_rank: shard of set B up, all-gather, shard of set A by expected
probes, all-reduce / row gather) and exchange the communicator id the same way
(dist.exchange_unique_id with the library's ncclGetUniqueId).  Only the engine is replaced: a host
stand-in (HostRank below, test infrastructure) that reassembles set B from the ranks' shards over
gloo exactly as cb_set_b_sharded lays them out, computes with the CPU oracle and reduces over gloo.
What this pins: the shard arithmetic (every sequence of B exactly once, in order; every seed of A
exactly once; global indices), and that partial results combine to the single-process result.  The
GPU engine itself under NCCL is tests/test_gpu_multi.py (needs >= 2 devices)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from _util import ROOT
from compairr_b200 import dist as cdist
from compairr_b200 import synth


def test_plan_shards_covers_everything_and_balances():
    rng = np.random.default_rng(0)
    lens = rng.integers(8, 23, 100000)
    for world in (1, 2, 3, 8):
        for d in (1, 2):
            sh = cdist.plan_shards(lens, world, 20, d, False)
            assert len(sh) == world and sh[0][0] == 0
            assert sum(c for _, c in sh) == lens.size
            for (f0, c0), (f1, _) in zip(sh, sh[1:]):
                assert f0 + c0 == f1
            w = cdist.probe_weights(lens, 20, d, False)
            loads = [w[f:f + c].sum() for f, c in sh]
            assert max(loads) / (sum(loads) / world) < 1.01
    assert cdist.plan_shards(np.zeros(0, np.int64), 4) == [(0, 0)] * 4
    assert cdist.plan_shards(np.array([5, 5]), 4)[-1][0] == 2  # more ranks than sequences


def test_shard_range_is_the_librarys():
    """dist.shard_range (pure Python, used by processes that never load the CUDA library) and
    cb_shard_range (what cb_set_b_sharded checks its argument against) agree, and tile the set."""
    from compairr_b200.engine import shard_range as c_shard_range
    for n in (0, 1, 7, 8, 9, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            at = 0
            for r in range(world):
                f, c = cdist.shard_range(n, r, world)
                assert (f, c) == c_shard_range(n, r, world)
                assert f == min(at, n) and c <= -(-n // world)
                at = f + c
            assert at == n


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class HostRank:
    """Host stand-in for compairr_b200.Engine in the CPU tests: the five methods dist.overlap_rank
    uses, over gloo + the CPU oracle."""

    def __init__(self, rank, world, kw, n_reps_a):
        self.rank, self.world, self.kw, self.n_reps_a = rank, world, kw, n_reps_a
        self.b = None
        self.m = None

    def set_b_sharded(self, shard, n_total):
        import torch.distributed as dist
        from compairr_b200.seqset import SeqSet
        f, c = cdist.shard_range(n_total, self.rank, self.world)
        assert shard.n == c                                        # what cb_set_b_sharded enforces
        parts = [None] * self.world
        dist.all_gather_object(parts, shard)
        assert sum(p.n for p in parts) == n_total
        lens = np.concatenate([p.lengths.astype(np.uint64) for p in parts])
        off = np.zeros(n_total + 1, np.uint64)
        np.cumsum(lens, out=off[1:])
        cat = lambda name, t: np.concatenate([getattr(p, name).astype(t) for p in parts])
        self.b = SeqSet(cat("residues", np.uint8), off, cat("v_gene", np.uint32), cat("j_gene", np.uint32),
                        cat("rep", np.uint32), cat("count", np.uint64), shard.n_reps)

    def run_a(self, a_shard):
        from oracle import oracle as orc
        m, _, _ = orc.overlap(a_shard, self.b, **self.kw)
        if not self.kw.get("existence"):
            full = np.zeros((self.n_reps_a, self.b.n_reps))
            full[: m.shape[0]] = m
            m = full
        self.m = m if self.m is None or self.kw.get("existence") else self.m + m

    def allreduce_matrix(self):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(self.m)
        dist.all_reduce(t)

    def matrix(self):
        return self.m


def _sets(existence):
    a = synth.small_dense_set(31, 1, 101) if existence else synth.small_dense_set(31, 5, 40)
    return a, synth.small_dense_set(32, 4, 91)


def _worker(rank, world, port, existence, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from compairr_b200 import Engine
    uid = cdist.exchange_unique_id(Engine.comm_unique_id)      # the id every rank would pass to cb_comm_init_rank
    a, b = _sets(existence)
    eng = HostRank(rank, world, dict(differences=1, indels=True, existence=existence), a.n_reps)
    full = cdist.overlap_rank(eng, a, b, rank, world, 1, True, existence)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    with open(os.path.join(out_dir, f"id{rank}.bin"), "wb") as f:
        f.write(uid)
    dist.destroy_process_group()


@pytest.mark.parametrize("existence", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_match_single_process(tmp_path, existence, world):
    from oracle import oracle as orc
    mp.spawn(_worker, args=(world, _free_port(), existence, str(tmp_path)), nprocs=world, join=True)
    a, b = _sets(existence)
    want, _, _ = orc.overlap(a, b, differences=1, indels=True, existence=existence)
    ids = {open(tmp_path / f"id{r}.bin", "rb").read() for r in range(world)}
    assert len(ids) == 1 and len(next(iter(ids))) == 128 and any(next(iter(ids)))   # one id, everywhere
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)
