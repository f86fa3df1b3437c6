"""Parity at BASELINE.json's sizes.  configs[1] (100 repertoires x 10 000 AA, self-comparison,
-m -d 0 and -d 1, with V/J genes, product score, with and without -f) is checked bit-exactly
against the CPU oracle; larger runs are checked through size-independent properties."""
import numpy as np
import pytest

from compairr_b200 import Engine, NarrowSet, OverlapOptions, overlap, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    return synth.make_set(1, 100, 10000)


@pytest.mark.parametrize("d", [0, 1])
@pytest.mark.parametrize("ignore_counts", [False, True])
def test_config2_self_comparison_bit_exact(c2, d, ignore_counts):
    m, _, info = overlap(c2, None, OverlapOptions(differences=d, ignore_counts=ignore_counts))
    mo, _, io = orc.overlap(c2, None, differences=d, ignore_counts=ignore_counts, threads=8)
    assert np.array_equal(m, mo)
    assert info["run"]["probes"] == io["probes"] and info["run"]["matches"] == io["matches"]
    assert np.array_equal(m, m.T)                      # a self-comparison is symmetric
    if ignore_counts:
        assert np.all(np.diag(m) >= 10000)             # every sequence matches itself


def test_config2_indels_and_d2_subset(c2):
    sub = c2.slice(0, 20000)
    for kw in (dict(differences=1, indels=True), dict(differences=2)):
        with Engine(OverlapOptions(**kw), n_reps_a=c2.n_reps) as eng:
            db = eng.upload(c2)
            eng.build_b(db)
            eng.run(db, 0, sub.n)
            m = eng.matrix()
        mo, _, _ = orc.overlap(sub, c2, threads=8, **kw)
        assert np.array_equal(m, mo)


@pytest.fixture(scope="module")
def big_pair():
    pool = synth.make_pool(5, 200000)
    a = synth.make_set(2, 40, 50000, pool=pool, indel_mutants=True)
    b = synth.make_set(3, 60, 50000, pool=pool, indel_mutants=True)
    return a, b


@pytest.mark.parametrize("kw", [dict(differences=1, indels=True), dict(differences=1), dict(differences=0)])
def test_large_run_properties(big_pair, kw):
    """2*10^6 x 3*10^6 sequences: (1) swapping the sets transposes the product matrix, (2) the
    -f matrix counts the pairs, so its sum equals the number of matches and of drained pairs,
    (3) d=0 matches are a subset of d=1, (4) the narrow-column one-call API agrees, (5) a
    two-shard run accumulates to the same matrix."""
    a, b = big_pair
    m_ab, _, i_ab = overlap(a, b, OverlapOptions(**kw))
    m_ba, _, i_ba = overlap(b, a, OverlapOptions(**kw))
    assert np.array_equal(m_ab, m_ba.T)
    assert i_ab["run"]["matches"] == i_ba["run"]["matches"]
    mf, pairs, i_f = overlap(a, b, OverlapOptions(ignore_counts=True, want_pairs=True, **kw))
    assert mf.sum() == i_f["run"]["matches"] == len(pairs) == i_ab["run"]["matches"]
    assert len(np.unique(pairs, axis=0)) == len(pairs)        # every pair exactly once
    if kw["differences"] == 1:
        m0, _, _ = overlap(a, b, OverlapOptions(differences=0, ignore_counts=True))
        assert np.all(m0 <= mf)
    with Engine(OverlapOptions(**kw), n_reps_a=a.n_reps) as eng:
        eng.set_b(NarrowSet.from_seqset(b))
        half = a.n // 2
        eng.run_a(NarrowSet.from_seqset(a.slice(0, half)))
        eng.run_a(a.slice(half, a.n - half))
        assert np.array_equal(eng.matrix(), m_ab)


def test_large_d2_and_d3_agree_on_shared_matches(big_pair):
    """d=2 (hash path) and d=3 (brute-force path) are different kernels: every d<=2 match is a
    d<=3 match, and restricted to Hamming distance <= 2 the pair sets coincide."""
    a, b = big_pair
    a_s, b_s = a.slice(0, 30000), b.slice(0, 200000)
    _, p2, _ = overlap(a_s, b_s, OverlapOptions(differences=2, want_pairs=True))
    _, p3, _ = overlap(a_s, b_s, OverlapOptions(differences=3, want_pairs=True))
    s2 = set(map(tuple, p2.tolist()))
    s3 = set(map(tuple, p3.tolist()))
    assert s2 <= s3

    def ham(x, y):
        sa = a.residues[int(a.offsets[x]):int(a.offsets[x + 1])]
        sb = b.residues[int(b.offsets[y]):int(b.offsets[y + 1])]
        return int(np.count_nonzero(sa != sb))
    assert {p for p in s3 if ham(*p) <= 2} == s2
