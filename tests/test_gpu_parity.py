"""GPU parity: the CUDA engine, through the C ABI, against the CPU oracle on the same seeded
inputs.  Bit-exact for integer-valued scores, counts and pair lists; 1e-12 relative for ratio."""
import numpy as np
import pytest

from compairr_b200 import OverlapOptions, SeqSet, overlap, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _pairs(p):
    return sorted(map(tuple, np.asarray(p).tolist()))


def _check(a, b, tol=None, **kw):
    eng_kw = dict(kw)
    m, p, info = overlap(a, b, OverlapOptions(want_pairs=True, nucleotides=a.nucleotides, **eng_kw))
    mo, po, io = orc.overlap(a, b, want_pairs=True, **kw)
    if tol is None:
        assert np.array_equal(m, mo)
    else:
        np.testing.assert_allclose(m, mo, rtol=tol, atol=0)
    assert _pairs(p) == _pairs(po)
    assert info["run"]["matches"] == io["matches"]
    if kw.get("differences", 0) <= 2:
        assert info["run"]["probes"] == io["probes"]
    return info


CASES = [(0, False), (1, False), (1, True), (2, False), (3, False), (4, False)]


@pytest.mark.parametrize("d,indels", CASES)
@pytest.mark.parametrize("ignore_genes", [False, True])
def test_dense_small_aa(d, indels, ignore_genes):
    a = synth.small_dense_set(3, 3, 120)
    b = synth.small_dense_set(4, 4, 120)
    _check(a, b, differences=d, indels=indels, ignore_genes=ignore_genes)


@pytest.mark.parametrize("d,indels", CASES)
def test_dense_small_nt(d, indels):
    a = synth.small_dense_set(5, 2, 150, alphabet="ACG", max_len=12, nucleotides=True)
    b = synth.small_dense_set(6, 3, 150, alphabet="ACG", max_len=12, nucleotides=True)
    _check(a, b, differences=d, indels=indels)


@pytest.mark.parametrize("score", ["product", "min", "max", "mean", "ratio"])
def test_scores(score):
    a = synth.small_dense_set(7, 3, 100)
    b = synth.small_dense_set(8, 3, 100)
    _check(a, b, tol=1e-12 if score == "ratio" else None, differences=1, indels=True, score=score)


def test_ignore_counts_and_self():
    a = synth.small_dense_set(9, 4, 100)
    _check(a, None, differences=1, ignore_counts=True)
    _check(a, None, differences=0, score="mh")
    _check(a, None, differences=0, score="jaccard")


@pytest.mark.parametrize("d,indels", [(0, False), (1, True), (2, False), (3, False)])
def test_existence(d, indels):
    a = synth.small_dense_set(10, 1, 90)
    b = synth.small_dense_set(11, 5, 100)
    _check(a, b, differences=d, indels=indels, existence=True)


@pytest.mark.parametrize("d,indels", [(0, False), (1, False), (1, True), (2, False)])
def test_cdr3_like(d, indels):
    pool = synth.make_pool(21, 3000)
    a = synth.make_set(22, 6, 1500, pool=pool, indel_mutants=True)
    b = synth.make_set(23, 7, 1500, pool=pool, indel_mutants=True)
    info = _check(a, b, differences=d, indels=indels)
    assert info["run"]["matches"] > 0


def test_dups():
    from compairr_b200 import Engine
    a = synth.small_dense_set(12, 3, 300, max_len=4)
    with Engine(OverlapOptions(differences=1), n_reps_a=a.n_reps) as eng:
        d = eng.upload(a)
        eng.build_b(d)
        assert eng.dups_b() == orc.count_dups(a) > 0
        assert eng.count_dups(d) == orc.count_dups(a)


@pytest.mark.parametrize("bpk", [24.0, 16.0, 4.0, 1.0])
@pytest.mark.parametrize("d,indels", [(0, False), (1, True), (2, False)])
def test_filter_geometries(bpk, d, indels):
    """Results must not depend on the filter geometry: from generous (24 bits per key in each class
    filter) down to saturated filters (1 bit per key: nearly every candidate reaches the table)."""
    pool = synth.make_pool(31, 2000)
    a = synth.make_set(32, 4, 1200, pool=pool, indel_mutants=True)
    b = synth.make_set(33, 5, 1200, pool=pool, indel_mutants=True)
    m, p, info = overlap(a, b, OverlapOptions(differences=d, indels=indels, want_pairs=True,
                                              bloom_bits_per_key=bpk))
    mo, po, io = orc.overlap(a, b, differences=d, indels=indels, want_pairs=True)
    assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
    assert 3 * info["build"]["bloom_bytes"] == info["build"]["bloom2_bytes"] > 0   # one class filter, the other three
    if d > 0 and bpk >= 16:
        assert info["run"]["bloom_pass"] < 0.05 * info["run"]["probes"]


def test_no_bloom_flag_gives_same_result():
    a = synth.small_dense_set(41, 3, 100)
    b = synth.small_dense_set(42, 3, 100)
    m0, _, _ = overlap(a, b, OverlapOptions(differences=1, indels=True))
    m1, _, _ = overlap(a, b, OverlapOptions(differences=1, indels=True, flags=2))
    assert np.array_equal(m0, m1)


@pytest.mark.parametrize("d,indels", [(0, False), (1, True), (2, False), (3, False)])
def test_narrow_columns_and_one_call_api(d, indels):
    """cb_set_b_cols / cb_run_a_cols (lengths + narrow dtypes, pipelined upload + fused build)
    give the same matrix, pairs and duplicate count as the reference-width path."""
    from compairr_b200 import Engine, NarrowSet
    pool = synth.make_pool(51, 3000)
    a = synth.make_set(52, 5, 2000, pool=pool, indel_mutants=True)
    b = synth.make_set(53, 6, 2000, pool=pool, indel_mutants=True)
    mo, po, io = orc.overlap(a, b, differences=d, indels=indels, want_pairs=True)
    for conv in (lambda s: s, NarrowSet.from_seqset):
        with Engine(OverlapOptions(differences=d, indels=indels, want_pairs=True), n_reps_a=a.n_reps) as eng:
            eng.set_b(conv(b))
            eng.run_a(conv(a))
            assert np.array_equal(eng.matrix(), mo)
            assert _pairs(eng.drain_pairs()) == _pairs(po)
            if d <= 2:
                assert eng.dups_b() == orc.count_dups(b)


def test_upload_of_a_slice_and_long_sequences():
    """A shard (offsets not starting at 0, index_base) and sequences longer than the 64 Zobrist
    rows the pipelined upload starts with (forces the re-hash / re-insert path)."""
    from compairr_b200 import Engine, NarrowSet
    a = synth.small_dense_set(61, 3, 200, alphabet="ACGT", min_len=60, max_len=90, nucleotides=True)
    b = synth.small_dense_set(61, 3, 200, alphabet="ACGT", min_len=60, max_len=90, nucleotides=True)
    b.residues[::7] = (b.residues[::7] + 1) % 4   # same sequences with scattered substitutions
    mo, po, _ = orc.overlap(a, b, differences=1, indels=True, want_pairs=True)
    with Engine(OverlapOptions(differences=1, indels=True, want_pairs=True, nucleotides=True), n_reps_a=a.n_reps) as eng:
        eng.set_b(NarrowSet.from_seqset(b))
        eng.run_a(a.slice(0, 250))
        eng.run_a(NarrowSet.from_seqset(a.slice(250, a.n - 250)))
        assert np.array_equal(eng.matrix(), mo)
        assert _pairs(eng.drain_pairs()) == _pairs(po)


def _long_mutant_sets(nucleotides, seed=71, n=240, lo=36, hi=90):
    """Set A: long random sequences; set B: for each of them one single-edit mutant (substitution,
    insertion or deletion at a position drawn over the WHOLE length, so most edits sit beyond
    position 32), plus exact copies and double-substitution mutants."""
    from compairr_b200.seqset import SeqSet
    rng = np.random.default_rng(seed)
    sigma = 4 if nucleotides else 20
    seqs_a = [rng.integers(0, sigma, int(rng.integers(lo, hi + 1))).astype(np.uint8) for _ in range(n)]
    seqs_b = []
    for k, q in enumerate(seqs_a):
        q = q.copy()
        kind = k % 5
        p = int(rng.integers(0, q.size))
        if kind == 0:
            q[p] = (q[p] + 1 + rng.integers(0, sigma - 1)) % sigma
        elif kind == 1:
            q = np.insert(q, p, rng.integers(0, sigma))
        elif kind == 2:
            q = np.delete(q, p)
        elif kind == 3:
            p2 = int(rng.integers(0, q.size))
            q[p] = (q[p] + 1) % sigma
            q[p2] = (q[p2] + 2) % sigma
        seqs_b.append(q.astype(np.uint8))

    def mk(seqs, reps):
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([x.size for x in seqs], out=off[1:])
        m = len(seqs)
        return SeqSet(np.concatenate(seqs), off, np.zeros(m, np.uint32), np.zeros(m, np.uint32),
                      (np.arange(m) % reps).astype(np.uint32), rng.integers(1, 5, m).astype(np.uint64), reps,
                      nucleotides=nucleotides)
    return mk(seqs_a, 3), mk(seqs_b, 2)


@pytest.mark.parametrize("nucleotides", [False, True])
@pytest.mark.parametrize("d,indels", [(1, False), (1, True), (2, False)])
def test_long_sequences_with_edits_beyond_position_32(d, indels, nucleotides):
    """True matches whose edit lies anywhere in sequences of 36-90 residues: the per-slot filter
    words of insertion and deletion slots >= 32 take a different path in the d=1 kernel (fetched
    when stored, not prefetched), and the parity of far positions must pick the right filter."""
    a, b = _long_mutant_sets(nucleotides)
    mo, po, io = orc.overlap(a, b, differences=d, indels=indels, want_pairs=True)
    assert io["matches"] >= (90 if d == 1 and not indels else 140)   # the construction really matches
    m, p, info = overlap(a, b, OverlapOptions(differences=d, indels=indels, want_pairs=True, nucleotides=nucleotides))
    assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
    assert info["run"]["probes"] == io["probes"]


@pytest.mark.parametrize("d", [3, 4])
@pytest.mark.parametrize("nucleotides", [False, True])
def test_d3_tensor_core_path(d, nucleotides):
    """Buckets big enough for the tcgen05 one-hot GEMM kernel (-g: length buckets only): same
    matrix and pairs as the oracle's pairwise definition, and as the CUDA-core kernel."""
    pool = synth.make_pool(71, 1500)
    a = synth.make_set(72, 4, 1500, pool=pool, nucleotides=nucleotides)
    b = synth.make_set(73, 5, 2000, pool=pool, nucleotides=nucleotides)
    kw = dict(differences=d, ignore_genes=True)
    m, p, info = overlap(a, b, OverlapOptions(want_pairs=True, nucleotides=nucleotides, **kw))
    mo, po, io = orc.overlap(a, b, want_pairs=True, threads=4, **kw)
    assert np.array_equal(m, mo)
    assert _pairs(p) == _pairs(po)
    assert info["run"]["matches"] == io["matches"]
    m2, p2, _ = overlap(a, b, OverlapOptions(want_pairs=True, nucleotides=nucleotides, flags=4, **kw))
    assert np.array_equal(m2, mo) and _pairs(p2) == _pairs(po)


def test_d4_every_pair_matches_tensor_core():
    """Worst case for the tcgen05 epilogue: every accumulator passes the threshold (all sequences
    are 2-substitution neighbours of one centre, so any two differ in <= 4 positions), which
    overflows the per-warp candidate queues while a TMEM stage is held; plus rows of residues
    16-19 (the (1,1,1,1) code) and a second bucket one residue shorter (partial last tiles)."""
    rng = np.random.default_rng(2024)
    def neighbours(centre, n):
        out = set()
        while len(out) < n:
            s = centre.copy()
            p = rng.choice(centre.size, 2, replace=False)
            s[p] = rng.integers(0, 20, 2)
            out.add(bytes(s))
        return [np.frombuffer(x, np.uint8) for x in sorted(out)]
    def make(n15, n14, n_reps, seed):
        c15 = np.array([16, 17, 18, 19, 0, 1, 2, 3, 16, 4, 19, 8, 12, 18, 7], np.uint8)
        seqs = neighbours(c15, n15) + neighbours(c15[:14], n14)
        r = np.random.default_rng(seed)
        off = np.zeros(len(seqs) + 1, np.uint64)
        np.cumsum([s.size for s in seqs], out=off[1:])
        n = len(seqs)
        return SeqSet(np.concatenate(seqs), off, np.zeros(n, np.uint32), np.zeros(n, np.uint32),
                      r.integers(0, n_reps, n).astype(np.uint32), r.integers(1, 9, n).astype(np.uint64), n_reps)
    a, b = make(600, 333, 3, 1), make(900, 415, 4, 2)
    kw = dict(differences=4, ignore_genes=True)
    m, p, info = overlap(a, b, OverlapOptions(want_pairs=True, **kw))
    mo, po, io = orc.overlap(a, b, want_pairs=True, threads=4, **kw)
    assert io["matches"] == 600 * 900 + 333 * 415
    assert np.array_equal(m, mo)
    assert _pairs(p) == _pairs(po)
    assert info["run"]["matches"] == io["matches"]
