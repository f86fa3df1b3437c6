"""Seed lengths across the three enumeration launches (variant.cu): <= 30 residues the ZP = 32
kernels, 31..94 the ZP = 96 kernels, longer the generic kernel — the reference has no length limit
(a single 250-nt junction used to make every cb_run fail).  Sets mix all three
classes, with true matches (single edits, double substitutions, copies) in every class; the generic
kernel is also forced onto ordinary CDR3-like sets (CB_FLAG_GENERIC_KERNEL)."""
import numpy as np
import pytest

from compairr_b200 import OverlapOptions, SeqSet, overlap, synth
from oracle import oracle as orc
from test_gpu_parity import _long_mutant_sets

pytestmark = pytest.mark.gpu
GENERIC = 16


def _pairs(p):
    return sorted(map(tuple, np.asarray(p).tolist()))


def _concat(sets):
    off = [np.zeros(1, np.uint64)]
    base = 0
    for s in sets:
        off.append(s.offsets[1:] + np.uint64(base))
        base += int(s.offsets[-1])
    cat = lambda f: np.concatenate([getattr(s, f) for s in sets])
    return SeqSet(cat("residues"), np.concatenate(off), cat("v_gene"), cat("j_gene"), cat("rep"), cat("count"),
                  sets[0].n_reps, nucleotides=sets[0].nucleotides)


def _mixed(nucleotides, classes=((5, 30), (31, 94), (95, 230))):
    parts = [_long_mutant_sets(nucleotides, seed=80 + k, n=120, lo=lo, hi=hi) for k, (lo, hi) in enumerate(classes)]
    a = _concat([p[0] for p in parts])
    b = _concat([p[1] for p in parts])
    rng = np.random.default_rng(5)
    perm = rng.permutation(a.n)       # classes interleaved: every warp batch sees seeds it must skip
    def take(s, idx):
        lens = np.diff(s.offsets).astype(np.int64)[idx]
        off = np.zeros(idx.size + 1, np.uint64)
        np.cumsum(lens, out=off[1:])
        res = np.concatenate([s.residues[int(s.offsets[i]):int(s.offsets[i + 1])] for i in idx])
        return SeqSet(res, off, s.v_gene[idx], s.j_gene[idx], s.rep[idx], s.count[idx], s.n_reps, nucleotides=s.nucleotides)
    return take(a, perm), take(b, rng.permutation(b.n))


@pytest.mark.parametrize("nucleotides", [False, True])
@pytest.mark.parametrize("d,indels", [(1, False), (1, True), (2, False)])
def test_all_three_length_classes_in_one_run(d, indels, nucleotides):
    a, b = _mixed(nucleotides)
    assert int(np.diff(a.offsets).max()) > 94 and int(np.diff(a.offsets).min()) <= 30
    mo, po, io = orc.overlap(a, b, differences=d, indels=indels, want_pairs=True, threads=4)
    assert io["matches"] > 100
    m, p, info = overlap(a, b, OverlapOptions(differences=d, indels=indels, want_pairs=True, nucleotides=nucleotides))
    assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
    assert info["run"]["probes"] == io["probes"] and info["run"]["matches"] == io["matches"]
    assert info["run"]["kernel_launches"] >= 3 + 1          # three enumeration launches + the table stage


@pytest.mark.parametrize("nucleotides,hi", [(False, 510), (True, 400)])
def test_longest_supported_sequences(nucleotides, hi):
    """Up to the descriptor's limit (510 residues): the generic kernel with the Zobrist table in
    global memory (amino acids) or shared memory (nucleotides)."""
    a, b = _mixed(nucleotides, classes=((hi - 60, hi),))
    for d, indels in ((1, True), (2, False)):
        if d == 2:
            a, b = a.slice(0, 16), b
        mo, po, io = orc.overlap(a, b, differences=d, indels=indels, want_pairs=True, threads=8)
        m, p, info = overlap(a, b, OverlapOptions(differences=d, indels=indels, want_pairs=True, nucleotides=nucleotides))
        assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
        assert info["run"]["probes"] == io["probes"]


def test_sequence_beyond_the_limit_is_an_error_not_a_wrong_answer():
    from compairr_b200.engine import EngineError
    a, b = _mixed(False, classes=((511, 520),))
    with pytest.raises(EngineError, match="510"):
        overlap(a, b, OverlapOptions(differences=1))
    m, _, _ = overlap(a, b, OverlapOptions(differences=0))       # d = 0 and d >= 3 have no such limit
    mo, _, _ = orc.overlap(a, b, differences=0)
    assert np.array_equal(m, mo)


@pytest.mark.parametrize("d,indels", [(1, False), (1, True), (2, False)])
@pytest.mark.parametrize("nucleotides", [False, True])
def test_generic_kernel_on_ordinary_sets(d, indels, nucleotides):
    pool = synth.make_pool(91, 2000)
    a = synth.make_set(92, 3, 1200 if d < 2 else 300, pool=pool, indel_mutants=True, nucleotides=nucleotides)
    b = synth.make_set(93, 4, 1500, pool=pool, indel_mutants=True, nucleotides=nucleotides)
    mo, po, io = orc.overlap(a, b, differences=d, indels=indels, want_pairs=True, threads=4)
    for flags in (0, GENERIC):
        m, p, info = overlap(a, b, OverlapOptions(differences=d, indels=indels, want_pairs=True, nucleotides=nucleotides, flags=flags))
        assert np.array_equal(m, mo) and _pairs(p) == _pairs(po)
        assert info["run"]["probes"] == io["probes"] and info["run"]["matches"] == io["matches"]
