"""The N>1 GPU path against the oracle: contexts joined in an NCCL communicator inside the library
(csrc/comm.cu), set B uploaded in shards and all-gathered over NVLink, set A sharded, partial
matrices all-reduced — through the C ABI from one process (threads, cb_comm_init_all) and through
the CLI's --gpus.  Skipped below two devices; a single-GPU box still runs the world-1 degenerate
forms of the same entry points."""
import os
import subprocess
import threading

import numpy as np
import pytest

from _util import CLI
from compairr_b200 import Engine, NarrowSet, OverlapOptions, cabi, synth
from compairr_b200 import dist as cdist
from compairr_b200.engine import comm_init_all
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

N_DEV = cabi.lib.cb_device_count()
need2 = pytest.mark.skipif(N_DEV < 2, reason="needs >= 2 CUDA devices")


def _pairs(p):
    return sorted(map(tuple, np.asarray(p).tolist()))


def _run_world(a, b, world, kw, want_pairs=False):
    """One Engine per device in this process, one thread per rank running dist.overlap_rank."""
    opts = [OverlapOptions(device=r, want_pairs=want_pairs, **kw) for r in range(world)]
    engs = [Engine(o, n_reps_a=1 if kw.get("existence") else a.n_reps) for o in opts]
    if world > 1:
        comm_init_all(engs)
    out, pairs, errs = [None] * world, [None] * world, []

    def rank(r):
        try:
            first, count = cdist.shard_range(b.n, r, world)
            shard = NarrowSet.from_seqset(b.slice(first, count))
            shard.n_reps, shard.index_base = b.n_reps, b.index_base
            engs[r].set_b_sharded(shard, b.n)
            f, c = cdist.plan_shards(a.lengths, world, a.sigma, kw.get("differences", 0), kw.get("indels", False))[r]
            engs[r].run_a(a.slice(f, c))
            if not kw.get("existence"):
                engs[r].allreduce_matrix()
            out[r] = engs[r].matrix()
            pairs[r] = engs[r].drain_pairs() if want_pairs else None
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))
    th = [threading.Thread(target=rank, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert not errs, errs
    dups = [e.dups_b() for e in engs]
    [e.close() for e in engs]
    return out, pairs, dups


@pytest.fixture(scope="module")
def sets():
    pool = synth.make_pool(201, 6000)
    a = synth.make_set(202, 7, 3000, pool=pool, indel_mutants=True)
    b = synth.make_set(203, 9, 3001, pool=pool, indel_mutants=True)   # odd size: a short last shard
    return a, b


@pytest.mark.parametrize("kw", [dict(differences=1, indels=True), dict(differences=2, ignore_genes=True),
                                dict(differences=0, score="mh"), dict(differences=3)])
def test_world1_sharded_entry_points(sets, kw):
    """cb_set_b_sharded / cb_allreduce_matrix with one rank degenerate to cb_set_b_cols / a no-op."""
    a, b = sets
    (m,), (p,), (dups,) = _run_world(a, b, 1, kw, want_pairs=True)
    mo, po, _ = orc.overlap(a, b, want_pairs=True, threads=4, **kw)
    assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
    if kw["differences"] <= 2:
        assert dups == orc.count_dups(b, ignore_genes=kw.get("ignore_genes", False))


@need2
@pytest.mark.parametrize("kw", [dict(differences=1, indels=True), dict(differences=2, ignore_genes=True),
                                dict(differences=0, score="jaccard"), dict(differences=1, score="ratio"),
                                dict(differences=3)])
def test_two_gpus_matrix_and_pairs_vs_oracle(sets, kw):
    a, b = sets
    world = min(N_DEV, 4) if kw["differences"] == 1 and kw.get("indels") else 2
    ms, ps, dups = _run_world(a, b, world, kw, want_pairs=True)
    mo, po, _ = orc.overlap(a, b, want_pairs=True, threads=4, **kw)
    for m in ms:                                   # every rank holds the reduced matrix
        if kw.get("score") == "ratio":
            np.testing.assert_allclose(m, mo, rtol=1e-12, atol=0)
        else:
            assert np.array_equal(m, mo)
    assert _pairs(np.concatenate(ps)) == _pairs(po)    # pairs stay per rank, with global indices
    if kw["differences"] <= 2:
        assert len(set(dups)) == 1 and dups[0] == orc.count_dups(b, ignore_genes=kw.get("ignore_genes", False))


@need2
def test_two_gpus_tiled_build_in_one_process():
    """Set B large enough for the tiled build (4.4e6 keys, 2^24 slots), both contexts in ONE process: the
    tile kernel's 64 KiB of dynamic shared memory has to be granted on every device."""
    pool = synth.make_pool(121, 400_000)
    b = synth.make_set(122, 44, 100_000, pool=pool, indel_mutants=True, workers=4)
    a = synth.make_set(123, 8, 10_000, pool=pool, indel_mutants=True)
    kw = dict(differences=1, indels=True)
    ms, _, dups = _run_world(a, b, 2, kw)
    mo, _, _ = orc.overlap(a, b, threads=8, **kw)
    assert np.array_equal(ms[0], mo) and np.array_equal(ms[1], mo)
    assert dups[0] == dups[1] == orc.count_dups(b)


@need2
def test_two_gpus_existence_rows(sets):
    a, b = sets
    q = a.slice(0, 5000)
    q.rep = np.zeros(q.n, np.uint32)
    q.n_reps = 1
    kw = dict(differences=1, indels=True, existence=True)
    ms, _, _ = _run_world(q, b, 2, kw)
    mo, _, _ = orc.overlap(q, b, threads=4, **kw)
    assert np.array_equal(np.concatenate(ms, axis=0), mo)


@need2
def test_two_gpus_self_comparison_more_ranks_than_work():
    """Self-comparison through the resident set B, and a set-A shard that is empty."""
    s = synth.small_dense_set(211, 3, 50)
    engs = [Engine(OverlapOptions(device=r, differences=1, indels=True), n_reps_a=s.n_reps) for r in range(2)]
    comm_init_all(engs)
    res = [None, None]

    def rank(r):
        first, count = cdist.shard_range(s.n, r, 2)
        sh = NarrowSet.from_seqset(s.slice(first, count))
        sh.n_reps, sh.index_base = s.n_reps, s.index_base
        engs[r].set_b_sharded(sh, s.n)
        if r == 0:                                  # rank 0 takes everything, rank 1 nothing
            engs[r].run(engs[r].resident_b(), 0, s.n)
        engs[r].allreduce_matrix()
        res[r] = engs[r].matrix()
    th = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    mo, _, _ = orc.overlap(s, None, differences=1, indels=True)
    assert np.array_equal(res[0], mo) and np.array_equal(res[1], mo)
    [e.close() for e in engs]


@need2
@pytest.mark.parametrize("args", [["-m", "-d", "1", "-i"], ["-m", "-d", "2", "-g", "-s", "min"], ["-m", "-d", "0", "-s", "MH"],
                                  ["-x", "-d", "1"], ["-m", "-d", "3"]])
def test_cli_gpus_2_equals_gpus_1_and_the_reference(sets, tmp_path, args):
    a, b = sets
    if args[0] == "-x":
        a = a.slice(0, 4000)
        a.rep = np.zeros(a.n, np.uint32)
        a.n_reps = 1
    fa, fb = tmp_path / "a.tsv", tmp_path / "b.tsv"
    a.write_tsv(str(fa), "a")
    b.write_tsv(str(fb), "b")
    outs = {}
    for g in (1, 2):
        out, pairs, log = tmp_path / f"out{g}.tsv", tmp_path / f"pairs{g}.tsv", tmp_path / f"log{g}.txt"
        r = subprocess.run([CLI] + args + [str(fa), str(fb), "--gpus", str(g), "-o", str(out), "-p", str(pairs), "-l", str(log)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + log.read_text()
        outs[g] = (out.read_text(), sorted(pairs.read_text().splitlines()))
    assert outs[1] == outs[2]
    if orc.have_reference():
        out, pairs = tmp_path / "ref.tsv", tmp_path / "refp.tsv"
        r = orc.run_reference(args + [str(fa), str(fb), "-o", str(out), "-p", str(pairs), "-l", os.devnull, "-t", "4"])
        assert r.returncode == 0, r.stderr
        assert out.read_text() == outs[2][0]
        assert sorted(pairs.read_text().splitlines()) == outs[2][1]
