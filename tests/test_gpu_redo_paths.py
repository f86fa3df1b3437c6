"""GPU parity for the product paths that only fire at size or under pressure, each forced at a size
the CPU oracle finishes in seconds (reference semantics: src/overlap.cc:168-251 probe/accumulate,
:455-507 pairs, :861-873 build):

  * candidate-queue overflow -> the chunk is skipped by the table kernel and redone in smaller
    pieces (engine.cu run_chunks), forced with cb_config.queue_capacity
  * pair-buffer overflow -> second pass into an exact-size buffer (engine.cu cb_run), forced with
    cb_config.pairs_capacity
  * radix-partitioned table build (engine.cu cb_table_insert: table >= 256 MiB, >= 2^22 keys) — the
    build every C3-sized bench step takes: tiles built in shared memory (kernels.cu
    build_tile_kernel), against the swept (flags 64) and the direct (flags 8) builds
  * the BASELINE.json config shapes C4 (-d 2 -g -s min) and C5 (-x -n -d 2 -p --no-matrix, and the
    d = 3 tensor-core path with (length, V, J) buckets)
"""
import numpy as np
import pytest

from compairr_b200 import Engine, OverlapOptions, SeqSet, cluster, dedup, overlap, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _pairs(p):
    return sorted(map(tuple, np.asarray(p).tolist()))


@pytest.fixture(scope="module")
def dense_pair():
    pool = synth.make_pool(101, 4000)
    a = synth.make_set(102, 5, 3000, pool=pool, indel_mutants=True)
    b = synth.make_set(103, 6, 3000, pool=pool, indel_mutants=True)
    return a, b


@pytest.mark.parametrize("d,indels", [(1, False), (1, True), (2, False)])
@pytest.mark.parametrize("existence", [False, True])
def test_queue_overflow_redo(dense_pair, d, indels, existence):
    """A 256-entry candidate queue overflows on every chunk of more than a few dozen seeds: the
    result must not change, whatever the number of redo levels."""
    a, b = dense_pair
    if existence:   # -x: set A is one repertoire (a view of the fixture with its own repertoire column)
        a = a.slice(0, 4000)
        a.rep = np.zeros(a.n, np.uint32)
        a.n_reps = 1
    kw = dict(differences=d, indels=indels, existence=existence)
    mo, po, io = orc.overlap(a, b, want_pairs=True, threads=4, **kw)
    for cap in (256, 4096):
        m, p, info = overlap(a, b, OverlapOptions(want_pairs=True, queue_capacity=cap, **kw))
        assert np.array_equal(m, mo)
        assert _pairs(p) == _pairs(po)
        assert info["run"]["matches"] == io["matches"] and info["run"]["probes"] == io["probes"]
    m, _, info0 = overlap(a, b, OverlapOptions(**kw))
    assert np.array_equal(m, mo)
    # the redo really happened: more launches than the plain run
    assert info["run"]["kernel_launches"] >= info0["run"]["kernel_launches"]


@pytest.mark.parametrize("d,indels", [(0, False), (1, True), (2, False), (3, False)])
@pytest.mark.parametrize("no_matrix", [False, True])
def test_pairs_overflow_second_pass(dense_pair, d, indels, no_matrix):
    """pairs_capacity = 16: every run overflows the device pair buffer and is redone for the pairs
    alone; matrix and match counts come from the first pass and must not be touched by the second."""
    a, b = dense_pair
    kw = dict(differences=d, indels=indels)
    mo, po, io = orc.overlap(a, b, want_pairs=True, threads=4, **kw)
    assert len(po) > 16
    m, p, info = overlap(a, b, OverlapOptions(want_pairs=True, pairs_capacity=16, no_matrix=no_matrix,
                                              queue_capacity=4096 if d in (1, 2) else 0, **kw))
    assert _pairs(p) == _pairs(po)
    assert info["run"]["matches"] == io["matches"] and info["run"]["pairs"] == len(po)
    if no_matrix:
        assert m is None
    else:
        assert np.array_equal(m, mo)


def test_pairs_overflow_across_chunked_runs(dense_pair):
    """Several cb_run calls on one context, each overflowing the pair buffer: pending pairs
    accumulate across calls, the matrix accumulates once per call."""
    a, b = dense_pair
    mo, po, _ = orc.overlap(a, b, differences=1, indels=True, want_pairs=True, threads=4)
    with Engine(OverlapOptions(differences=1, indels=True, want_pairs=True, pairs_capacity=64,
                               queue_capacity=1024), n_reps_a=a.n_reps) as eng:
        db, da = eng.upload(b), eng.upload(a)
        eng.build_b(db)
        for first in range(0, a.n, 4000):
            eng.run(da, first, min(4000, a.n - first))
        assert np.array_equal(eng.matrix(), mo)
        assert _pairs(eng.drain_pairs()) == _pairs(po)


@pytest.mark.parametrize("d,indels", [(1, True), (2, False)])
def test_cluster_network_with_tiny_queue_and_pair_buffer(d, indels):
    """-c builds its network through the same kernels in pairs mode (cluster.cu): tiny queue and
    pair buffer must give the reference's clusters."""
    pool = synth.make_pool(111, 1500)
    s = synth.make_set(112, 3, 1500, pool=pool, indel_mutants=True)
    o_order, o_no, o_size, o_ncl, o_edges = orc.cluster(s, d, indels, False)
    order, no, size, info = cluster(s, OverlapOptions(differences=d, indels=indels, queue_capacity=512,
                                                      pairs_capacity=32))
    assert info["clusters"] == o_ncl and info["edges"] == o_edges
    assert np.array_equal(order, o_order) and np.array_equal(no, o_no) and np.array_equal(size, o_size)


@pytest.fixture(scope="module")
def partition_sized():
    """Set B large enough for the partitioned build: 4.4e6 keys -> 2^24 slots = 256 MiB."""
    pool = synth.make_pool(121, 400_000)
    b = synth.make_set(122, 44, 100_000, pool=pool, indel_mutants=True, workers=4)
    a = synth.make_set(123, 8, 10_000, pool=pool, indel_mutants=True)
    return a, b


@pytest.mark.parametrize("kw", [dict(differences=0), dict(differences=1, indels=True), dict(differences=2, ignore_genes=True)])
def test_partitioned_build_vs_oracle(partition_sized, kw):
    a, b = partition_sized
    if kw["differences"] == 2:
        a = a.slice(0, 20_000)
    with Engine(OverlapOptions(**kw), n_reps_a=a.n_reps) as eng:
        db = eng.upload(b)
        eng.build_b(db)
        st = eng.stats()
        assert st["table_slots"] * 16 >= 256 << 20
        assert st["kernel_launches"] >= 3 + 4       # clear, reset, insert, dups + iota and the radix passes
        dups = eng.dups_b()
        da = eng.upload(a)
        eng.run(da)
        m, run = eng.matrix(), eng.stats()
        # the same set through the swept build (sorted keys into a cleared table) and WITHOUT the
        # partition sort gives the same table contents
        for flags in (64, 8):
            with Engine(OverlapOptions(flags=flags, **kw), n_reps_a=a.n_reps) as eng2:
                db2 = eng2.upload(b)
                eng2.build_b(db2)
                if flags == 8:
                    assert eng2.stats()["kernel_launches"] < st["kernel_launches"]
                assert eng2.dups_b() == dups
                da2 = eng2.upload(a)
                eng2.run(da2)
                assert np.array_equal(eng2.matrix(), m)
                assert eng2.stats()["matches"] == run["matches"]
        # built twice in place (links reset, table buffers reused, never cleared): same answers
        eng.build_b(db)
        assert eng.dups_b() == dups
        eng.clear_matrix()
        eng.run(da)
        assert np.array_equal(eng.matrix(), m)
    mo, _, io = orc.overlap(a, b, threads=8, **kw)
    assert np.array_equal(m, mo)
    assert run["matches"] == io["matches"] and run["probes"] == io["probes"]
    assert dups == orc.count_dups(b, ignore_genes=kw.get("ignore_genes", False))


def test_tiled_build_occurrence_lists_dedup_and_pairs(partition_sized):
    """The occurrence lists the tiled build links (shared-memory atomicExch per duplicate) carry
    -z and the pair list: same leaders, counts and pairs as the direct build, spilled keys included
    (a table at 94 % load spills a good part of every tile)."""
    _, b = partition_sized
    s = b.slice(0, 4_300_000)
    s.rep = (s.rep % 3).astype(np.uint32)
    s.n_reps = 3
    got = {}
    for flags in (0, 8):
        lead, cnt, merged = dedup(s, OverlapOptions(flags=flags))
        got[flags] = (lead, cnt, merged)
    assert got[0][2] == got[8][2] > 0
    assert np.array_equal(got[0][0], got[8][0]) and np.array_equal(got[0][1], got[8][1])
    o_lead, o_cnt, o_merged = orc.dedup(s.slice(0, 300_000))
    lead, cnt, merged = dedup(s.slice(0, 300_000), OverlapOptions())
    assert merged == o_merged and np.array_equal(lead, o_lead) and np.array_equal(cnt, o_cnt)


def _with_hub(b, k, src, at=0):
    """b plus k copies of sequence `at` of src (its V and J too), spread over the repertoires."""
    lo, hi = int(src.offsets[at]), int(src.offsets[at + 1])
    lens = np.concatenate([b.lengths, np.full(k, hi - lo, dtype=b.lengths.dtype)])
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return SeqSet(np.concatenate([b.residues, np.tile(src.residues[lo:hi], k)]), off,
                  np.concatenate([b.v_gene, np.full(k, src.v_gene[at], np.uint32)]),
                  np.concatenate([b.j_gene, np.full(k, src.j_gene[at], np.uint32)]),
                  np.concatenate([b.rep, (np.arange(k) % b.n_reps).astype(np.uint32)]),
                  np.concatenate([b.count, np.ones(k, np.uint64)]), b.n_reps)


def test_tiled_build_hub_sequence(partition_sized):
    """One sequence present many thousand times: all its copies meet in one slot of one tile (the tile
    kernel sets them aside and links them in its second pass, one atomicExch each on the same word) and
    every matching seed walks the whole list."""
    a, b = partition_sized
    kw = dict(differences=1, indels=True)
    q = a.slice(0, 20_000)
    s = _with_hub(b, 20_000, q)
    m, _, info = overlap(q, s, OverlapOptions(**kw))
    assert info["build"]["table_slots"] * 16 >= 256 << 20
    mo, _, io = orc.overlap(q, s, threads=8, **kw)
    assert np.array_equal(m, mo) and info["run"]["matches"] == io["matches"] > 20_000
    assert info["dups_b"] == orc.count_dups(s) > 19_000
    # a hub too large for the serial oracle (its insert is quadratic in the copies): tiled against direct build
    s = _with_hub(b, 200_000, q)
    res = []
    for flags in (0, 8):
        m, _, info = overlap(q, s, OverlapOptions(flags=flags, **kw))
        res.append((m, info["run"]["matches"], info["dups_b"]))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]
    assert res[0][2] > 199_000


def test_tiled_build_crowded_table():
    """1.5e7 keys in 2^24 slots (89 % load): probe runs are long and cross tile ends all the time,
    so a good part of the keys takes the spill route; matrix, pairs and duplicate count against the
    direct build and the oracle."""
    pool = synth.make_pool(121, 400_000)
    s = synth.make_set(124, 150, 100_000, pool=pool, indel_mutants=True, workers=4)
    q = synth.make_set(125, 4, 5_000, pool=pool, indel_mutants=True)
    res = {}
    for flags in (0, 8):
        m, p, info = overlap(q, s, OverlapOptions(differences=1, indels=True, want_pairs=True, table_load_pct=90,
                                                  flags=flags))
        assert info["build"]["table_slots"] == 1 << 24
        res[flags] = (m, _pairs(p), info["run"]["matches"], info["dups_b"])
    assert np.array_equal(res[0][0], res[8][0]) and res[0][1:] == res[8][1:]
    mo, po, _ = orc.overlap(q, s, differences=1, indels=True, want_pairs=True, threads=8)
    assert np.array_equal(res[0][0], mo) and res[0][1] == _pairs(po)
    assert res[0][3] == orc.count_dups(s)


def test_c4_shape_d2_ignore_genes_min(partition_sized):
    """BASELINE config 4: -m -d 2 -g -s min (Jaccard's summand), two sets."""
    a, b = partition_sized
    a_s, b_s = a.slice(0, 30_000), b.slice(0, 600_000)
    m, _, info = overlap(a_s, b_s, OverlapOptions(differences=2, ignore_genes=True, score="min"))
    mo, _, io = orc.overlap(a_s, b_s, differences=2, ignore_genes=True, score="min", threads=8)
    assert np.array_equal(m, mo)
    assert info["run"]["matches"] == io["matches"] and info["run"]["probes"] == io["probes"]
    assert info["run"]["matches"] > 1000


@pytest.fixture(scope="module")
def c5_sets():
    """BASELINE config 5: a nucleotide query set in ONE repertoire against a multi-repertoire set."""
    pool = synth.make_pool(131, 30_000)
    b = synth.make_set(132, 20, 10_000, pool=pool, nucleotides=True)
    q = synth.make_set(133, 10, 10_000, pool=pool, nucleotides=True, single_repertoire=True)
    return q, b


def test_c5_shape_existence_nt_d2_pairs_no_matrix(c5_sets):
    """-x -n -d 2 -p --no-matrix on 10^5 queries: the pair list is the whole result."""
    q, b = c5_sets
    kw = dict(differences=2, existence=True)
    m, p, info = overlap(q, b, OverlapOptions(nucleotides=True, want_pairs=True, no_matrix=True,
                                              pairs_capacity=1 << 12, **kw))
    _, po, io = orc.overlap(q, b, want_pairs=True, want_matrix=False, threads=8, **kw)
    assert m is None
    assert _pairs(p) == _pairs(po)
    assert info["run"]["matches"] == io["matches"] and info["run"]["probes"] == io["probes"]
    # with the matrix: rows are query sequences (overlap.cc:226)
    m, _, _ = overlap(q.slice(0, 20_000), b, OverlapOptions(nucleotides=True, **kw))
    mo, _, _ = orc.overlap(q.slice(0, 20_000), b, threads=8, **kw)
    assert m.shape == (20_000, b.n_reps) and np.array_equal(m, mo)


@pytest.mark.parametrize("nucleotides", [True, False])
def test_d3_tensor_core_with_vj_buckets(nucleotides):
    """d = 3 WITHOUT -g: the joins are (length, V, J) buckets.  Few genes make them large enough for
    the tcgen05 kernel; the result must equal the CUDA-core kernel's and the oracle's."""
    pool = synth.make_pool(141, 3000)
    a = synth.make_set(142, 4, 4000, pool=pool, nucleotides=nucleotides, single_repertoire=False)
    b = synth.make_set(143, 5, 6000, pool=pool, nucleotides=nucleotides)
    for s in (a, b):
        s.v_gene %= 3
        s.j_gene %= 2
    kw = dict(differences=3)
    m, p, info = overlap(a, b, OverlapOptions(want_pairs=True, nucleotides=nucleotides, **kw))
    mo, po, io = orc.overlap(a, b, want_pairs=True, threads=8, **kw)
    assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
    assert info["run"]["matches"] == io["matches"]
    m2, p2, info2 = overlap(a, b, OverlapOptions(want_pairs=True, nucleotides=nucleotides, flags=4, **kw))
    assert np.array_equal(m2, mo) and _pairs(p2) == _pairs(po)
    # the two runs really took different kernels
    assert info["run"]["kernel_launches"] != info2["run"]["kernel_launches"]


@pytest.mark.parametrize("score", ["product", "ratio", "min"])
@pytest.mark.parametrize("d,indels", [(0, False), (1, True), (2, False)])
def test_matrix_tile_equals_global_atomics(d, indels, score):
    """Small matrices are accumulated in CTA-private shared-memory tiles with warp-aggregated
    atomics (device_utils.cuh accumulate_warp); CB_FLAG_NO_SMEM_TILE sends every match straight to
    the global matrix.  Low-complexity sets: many matches per cell, many lanes on the same cell."""
    a = synth.small_dense_set(301, 3, 4000, max_len=6)
    b = synth.small_dense_set(302, 2, 5000, max_len=6)
    kw = dict(differences=d, indels=indels, score=score)
    mo, _, io = orc.overlap(a, b, threads=8, **kw)
    assert io["matches"] > 20 * a.n                      # dense: tens to hundreds of matches per seed
    for flags in (0, 1):
        m, _, info = overlap(a, b, OverlapOptions(flags=flags, **kw))
        if score == "ratio":
            np.testing.assert_allclose(m, mo, rtol=1e-12, atol=0)
        else:
            assert np.array_equal(m, mo)
        assert info["run"]["matches"] == io["matches"]


def test_matrix_too_large_for_a_tile():
    """More cells than the tile holds (> 1024): the global-atomics path, same answer."""
    a = synth.small_dense_set(311, 120, 40, max_len=5)
    b = synth.small_dense_set(312, 110, 40, max_len=5)
    m, _, info = overlap(a, b, OverlapOptions(differences=1, indels=True))
    mo, _, io = orc.overlap(a, b, differences=1, indels=True, threads=4)
    assert m.shape == (120, 110) and np.array_equal(m, mo) and info["run"]["matches"] == io["matches"]


def test_degenerate_family_saturates_one_filter_word_only():
    """20 000 keys that differ ONLY at positions of one class (p = 0 mod 4) share their word in
    class filter 0 (its index is blind to exactly those positions): that word saturates and every
    candidate looked up there reaches the table stage.  Results must stay exact, and the damage
    bounded: the other three filters are unaffected, so at most the slots of one class in four
    lose their filter."""
    rng = np.random.default_rng(77)
    L, n = 16, 20_000
    base = rng.integers(0, 20, L).astype(np.uint8)
    fam = np.tile(base, (n, 1))
    fam[:, [0, 4, 8, 12]] = rng.integers(0, 20, (n, 4))
    fam = np.unique(fam, axis=0)
    n = fam.shape[0]

    def mk(rows, reps, seed):
        r = np.random.default_rng(seed)
        m = rows.shape[0]
        off = np.arange(m + 1, dtype=np.uint64) * np.uint64(L)
        return SeqSet(rows.reshape(-1), off, np.zeros(m, np.uint32), np.zeros(m, np.uint32),
                      r.integers(0, reps, m).astype(np.uint32), r.integers(1, 4, m).astype(np.uint64), reps)
    b = mk(fam, 3, 1)
    a = mk(fam[rng.permutation(n)[:4000]], 2, 2)
    healthy = rng.integers(0, 20, (n, L)).astype(np.uint8)
    b_h, a_h = mk(healthy, 3, 1), mk(healthy[:4000], 2, 2)
    for kw in (dict(differences=1, indels=True), dict(differences=2)):
        mo, _, io = orc.overlap(a, b, threads=8, **kw)
        m, _, info = overlap(a, b, OverlapOptions(**kw))
        assert np.array_equal(m, mo) and info["run"]["matches"] == io["matches"]
        _, _, info_h = overlap(a_h, b_h, OverlapOptions(**kw))
        frac = info["run"]["bloom_pass"] / info["run"]["probes"]
        # only slots of class 0 can be hit by the saturated word; the family's true neighbours pass anyway
        assert frac < 0.30, frac
        assert info_h["run"]["bloom_pass"] / info_h["run"]["probes"] < 0.02
