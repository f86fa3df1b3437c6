"""Pins the CPU oracle (oracle/overlap_oracle.c) to the reference: every golden case was produced
by the unmodified reference binary (tests/golden/make_golden.py), including the reference's own
fixture test/expected.tsv (case ref_ab_d1_i) and the README worked examples."""
import os

import numpy as np
import pytest

from _util import GOLDEN_DIR, assert_matrix_text, golden_cases, hot_opts, is_integer_score, parse_args
from compairr_b200 import report, synth
from compairr_b200.seqset import encode_sequences, read_airr_pair
from oracle import oracle as orc

CASES = [c for c in golden_cases() if c["rc"] == 0]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("method", [0, 1])
def test_oracle_matches_reference_output(case, method):
    o = parse_args(case["args"])
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    a, b = read_airr_pair(files[0], files[1] if len(files) > 1 else None, o["nucleotides"])
    m, pairs, _ = orc.overlap(a, b, want_pairs=case["pairs"], method=method, **hot_opts(o))
    text = report.format_matrix(m, a, b or a, o["score"], o["existence"], o["alternative"])
    assert_matrix_text(text, case["output"], exact=is_integer_score(o))
    if case["pairs"]:
        header, rows = report.format_pairs(pairs, a, b or a, o["distance"])
        assert header == case["pairs_header"]
        assert sorted(rows) == case["pairs_sorted"]


def test_expected_tsv_of_the_reference():
    """test/test.sh:9-11: -m seta setb -d 1 -i must give test/expected.tsv."""
    case = next(c for c in CASES if c["name"] == "ref_ab_d1_i")
    assert case["output"] == "#\tB1\tB2\nA1\t0\t7\nA2\t45\t0\n"


# known-answer variant counts (from the reference's generate_variants)
KNOWN = [("C", (1, 20, 59, 20)), ("CASSF", (1, 96, 215, 3706)), ("AAAAA", (1, 96, 212, 3706)),
         ("CAASSF", (1, 115, 253, 5530)), ("CASSLRVGGYGYTF", (1, 267, 565, 33118))]


@pytest.mark.parametrize("seq,want", KNOWN)
def test_variant_counts(seq, want):
    codes = encode_sequences([seq])[0]
    got = tuple(len(orc.enumerate_variants(codes, 20, d, i)[0]) for d, i in [(0, 0), (1, 0), (1, 1), (2, 0)])
    assert got == want


@pytest.mark.parametrize("seq", ["C", "AAAAA", "CAASSF", "ACCA", "CASSLRVGGYGYTF"])
@pytest.mark.parametrize("d,indels", [(1, False), (1, True), (2, False)])
def test_variants_are_distinct_sequences(seq, d, indels):
    """Each variant is enumerated exactly once (SURVEY.md warning 3) and is what its record says."""
    codes = encode_sequences([seq])[0]
    recs, seqs = orc.enumerate_variants(codes, 20, d, indels)
    assert len(set(seqs)) == len(seqs)
    base = tuple(int(x) for x in codes)
    for (kind, p1, r1, p2, r2), s in zip(recs.tolist(), seqs):
        if kind == 0:
            want = base
        elif kind == 1:
            want = base[:p1] + (r1,) + base[p1 + 1:]
        elif kind == 2:
            want = base[:p1] + base[p1 + 1:]
        elif kind == 3:
            want = base[:p1] + (r1,) + base[p1:]
        else:
            t = list(base)
            t[p1], t[p2] = r1, r2
            want = tuple(t)
        assert s == want


@pytest.mark.parametrize("d,indels", [(0, False), (1, False), (1, True), (2, False), (3, False)])
@pytest.mark.parametrize("ignore_genes", [False, True])
def test_hash_path_equals_definition(d, indels, ignore_genes):
    """The restated hash/Bloom/variant machinery and the pairwise definition agree (and both agree
    with a pure-numpy brute force)."""
    a = synth.small_dense_set(3, 3, 80)
    b = synth.small_dense_set(4, 4, 80)
    kw = dict(differences=d, indels=indels, ignore_genes=ignore_genes)
    m0, p0, _ = orc.overlap(a, b, want_pairs=True, **kw)
    m1, p1, _ = orc.overlap(a, b, want_pairs=True, method=1, threads=3, **kw)
    m2, p2 = orc.brute_force(a, b, **kw)
    assert np.array_equal(m0, m1) and np.array_equal(m0, m2)
    assert sorted(map(tuple, p0.tolist())) == sorted(map(tuple, p1.tolist())) == sorted(map(tuple, p2.tolist()))


def test_threads_do_not_change_integer_results():
    a = synth.make_set(5, 4, 400)
    m1, _, i1 = orc.overlap(a, None, differences=1, threads=1)
    m4, _, i4 = orc.overlap(a, None, differences=1, threads=4)
    assert np.array_equal(m1, m4) and i1["probes"] == i4["probes"]


def test_dups_restatement():
    a = synth.small_dense_set(12, 3, 300, max_len=4)
    # definition: sum over groups of identical (rep, v, j, sequence) of (size - 1)
    keys = {}
    for i in range(a.n):
        k = (int(a.rep[i]), int(a.v_gene[i]), int(a.j_gene[i]), a.sequence(i))
        keys[k] = keys.get(k, 0) + 1
    assert orc.count_dups(a) == sum(v - 1 for v in keys.values()) > 0


@pytest.mark.skipif(not orc.have_reference(), reason="oracle/_ref/compairr not built")
@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("args", [["-d", "1", "-i"], ["-d", "2", "-g"], ["-d", "0", "-s", "MH"], ["-d", "3"]])
def test_live_differential_vs_reference(tmp_path, seed, args):
    """Seeded inputs through the unmodified reference binary vs the oracle, here and now."""
    pool = synth.make_pool(seed, 300)
    a = synth.make_set(seed * 10 + 1, 3, 150, pool=pool, indel_mutants=True)
    b = synth.make_set(seed * 10 + 2, 4, 150, pool=pool, indel_mutants=True)
    fa, fb, out = tmp_path / "a.tsv", tmp_path / "b.tsv", tmp_path / "o.tsv"
    a.write_tsv(str(fa), "a")
    b.write_tsv(str(fb), "b")
    r = orc.run_reference(["-m"] + args + [str(fa), str(fb), "-o", str(out), "-l", "/dev/null", "-t", "2"])
    assert r.returncode == 0, r.stderr
    o = parse_args(args)
    a2, b2 = read_airr_pair(str(fa), str(fb))
    m, _, _ = orc.overlap(a2, b2, **hot_opts(o))
    assert_matrix_text(report.format_matrix(m, a2, b2, o["score"]), out.read_text(), exact=is_integer_score(o))


# ---- `-c` / `-z`: the restatements of src/cluster.cc and src/dedup.cc against the reference's files ----

def _cz_cases():
    import json
    with open(os.path.join(GOLDEN_DIR, "golden_cz.json")) as f:
        return json.load(f)["cases"]


CZ = _cz_cases()


def _cz_opts(args):
    o = dict(differences=0, indels=False, ignore_genes=False, ignore_counts=False, nucleotides=False)
    it = iter(args)
    for a in it:
        if a == "-d":
            o["differences"] = int(next(it))
        elif a == "-t":
            next(it)
        else:
            o[{"-i": "indels", "-g": "ignore_genes", "-f": "ignore_counts", "-n": "nucleotides"}.get(a, "_")] = True
    o.pop("_", None)
    return o


@pytest.mark.parametrize("case", CZ, ids=[c["name"] for c in CZ])
def test_cluster_and_dedup_oracle_match_reference_files(case):
    from compairr_b200 import report
    from compairr_b200.seqset import read_airr_pair
    o = _cz_opts(case["args"])
    s, _ = read_airr_pair(os.path.join(GOLDEN_DIR, case["file"]), None, o["nucleotides"])
    if case["args"][0] == "-c":
        order, no, size, ncl, _ = orc.cluster(s, o["differences"], o["indels"], o["ignore_genes"])
        assert report.format_clusters(order, no, size, s) == case["output"]
        assert case["log"] == [f"Clusters:          {ncl}"]
    else:
        lead, cnt, merged = orc.dedup(s, o["ignore_genes"], o["ignore_counts"])
        assert report.format_dedup(lead, cnt, s, o["ignore_genes"]) == case["output"]
        assert case["log"] == [f"Duplicates merged: {merged}"]
