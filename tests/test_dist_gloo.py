"""The N>1 path on CPU: world_size-2 gloo processes exercise the shard planner and the matrix
allreduce / row gather of compairr_b200.dist.  The per-rank compute is injected (here: the CPU
oracle, as the checker's stand-in for the GPU engine, which needs a device)."""
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


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, existence, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if existence:
        a = synth.small_dense_set(31, 1, 101)
    else:
        a = synth.small_dense_set(31, 5, 40)
    b = synth.small_dense_set(32, 4, 90)

    def compute(shard, bb):
        m, _, _ = orc.overlap(shard, bb, differences=1, indels=True, existence=existence)
        if not existence:          # matrix mode: rows are the repertoires of the WHOLE set A
            assert m.shape[0] == a.n_reps
        return m
    full = cdist.sharded_overlap(a, b, compute, rank, world, 1, True, existence)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    dist.destroy_process_group()


@pytest.mark.parametrize("existence", [False, True])
def test_world2_gloo_matches_single_process(tmp_path, existence):
    from oracle import oracle as orc
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), existence, str(tmp_path)), nprocs=world, join=True)
    a = synth.small_dense_set(31, 1, 101) if existence else synth.small_dense_set(31, 5, 40)
    b = synth.small_dense_set(32, 4, 90)
    want, _, _ = orc.overlap(a, b, differences=1, indels=True, existence=existence)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert np.array_equal(got, want)
