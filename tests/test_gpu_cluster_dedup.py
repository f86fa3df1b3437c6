"""GPU parity for the `-c` (cluster) and `-z` (deduplicate) commands: the compiled CLI and the C ABI
against the reference binary's golden files (byte for byte, tests/golden/golden_cz.json), against
the oracle's restatements of src/cluster.cc / src/dedup.cc on seeded CDR3-like sets, and through
size-independent properties at 10^6 sequences."""
import json
import os
import subprocess

import numpy as np
import pytest

from _util import CLI, GOLDEN_DIR
from compairr_b200 import OverlapOptions, cluster, dedup, overlap, report, synth
from compairr_b200.seqset import read_airr_pair
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN_DIR, "golden_cz.json")) as f:
    CZ = json.load(f)["cases"]


def _opts(args):
    o = dict(differences=0, indels=False, ignore_genes=False, ignore_counts=False, nucleotides=False)
    it = iter(args)
    for a in it:
        if a == "-d":
            o["differences"] = int(next(it))
        elif a == "-t":
            next(it)
        elif a in ("-i", "-g", "-f", "-n"):
            o[{"-i": "indels", "-g": "ignore_genes", "-f": "ignore_counts", "-n": "nucleotides"}[a]] = True
    return o


@pytest.mark.parametrize("case", CZ, ids=[c["name"] for c in CZ])
def test_cli_reproduces_reference_files(case, tmp_path):
    out, log = tmp_path / "out.tsv", tmp_path / "log.txt"
    cmd = [CLI] + case["args"] + [os.path.join(GOLDEN_DIR, case["file"]), "-o", str(out), "-l", str(log)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + log.read_text()
    assert out.read_text() == case["output"]
    assert [ln for ln in log.read_text().splitlines() if ln.startswith(("Clusters:", "Duplicates merged:"))] == case["log"]


@pytest.mark.parametrize("case", CZ[::2], ids=[c["name"] for c in CZ[::2]])
def test_cabi_reproduces_reference_files(case):
    o = _opts(case["args"])
    s, _ = read_airr_pair(os.path.join(GOLDEN_DIR, case["file"]), None, o["nucleotides"])
    eo = OverlapOptions(**o)
    if case["args"][0] == "-c":
        order, no, size, info = cluster(s, eo)
        assert report.format_clusters(order, no, size, s) == case["output"]
        assert case["log"] == [f"Clusters:          {info['clusters']}"]
    else:
        lead, cnt, merged = dedup(s, eo)
        assert report.format_dedup(lead, cnt, s, o["ignore_genes"]) == case["output"]
        assert case["log"] == [f"Duplicates merged: {merged}"]


@pytest.mark.parametrize("d,indels,n", [(0, False, 4000), (1, False, 4000), (1, True, 4000), (2, False, 400), (3, False, 1500)])
@pytest.mark.parametrize("ignore_genes", [False, True])
def test_cluster_cdr3_like_equals_oracle(d, indels, n, ignore_genes):
    pool = synth.make_pool(31, max(n // 4, 10))
    s = synth.make_set(32, 4, n // 4, pool=pool, indel_mutants=True)
    order, no, size, info = cluster(s, OverlapOptions(differences=d, indels=indels, ignore_genes=ignore_genes))
    o_order, o_no, o_size, o_ncl, o_edges = orc.cluster(s, d, indels, ignore_genes)
    assert info["clusters"] == o_ncl and info["edges"] == o_edges
    assert np.array_equal(order, o_order) and np.array_equal(no, o_no) and np.array_equal(size, o_size)


@pytest.mark.parametrize("ignore_genes,ignore_counts", [(False, False), (True, False), (False, True)])
def test_dedup_equals_oracle(ignore_genes, ignore_counts):
    pool = synth.make_pool(41, 3000)
    s = synth.make_set(42, 5, 20000, pool=pool, indel_mutants=False)
    # fold the five repertoires into two so that many sequences repeat inside a repertoire
    s.rep = (s.rep % 2).astype(np.uint32)
    s.n_reps = 2
    lead, cnt, merged = dedup(s, OverlapOptions(ignore_genes=ignore_genes, ignore_counts=ignore_counts))
    o_lead, o_cnt, o_merged = orc.dedup(s, ignore_genes, ignore_counts)
    assert merged == o_merged and merged > 0
    assert np.array_equal(lead, o_lead) and np.array_equal(cnt, o_cnt)


def test_large_cluster_and_dedup_properties():
    """10^6 sequences: the cluster rows are a permutation, sizes are consistent and non-increasing,
    every d=1 match of the self-overlap joins two members of one cluster, and nothing else does
    (clusters = connected components, checked with a union-find over the pair list)."""
    pool = synth.make_pool(51, 100_000)
    s = synth.make_set(52, 10, 100_000, pool=pool, indel_mutants=True)
    n = s.n
    o = OverlapOptions(differences=1, indels=True)
    order, no, size, info = cluster(s, o)
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    assert np.all(np.diff(no.astype(np.int64)) >= 0) and no[0] == 1 and no[-1] == info["clusters"]
    assert np.all(np.diff(size.astype(np.int64)) <= 0)
    assert np.array_equal(np.bincount(no)[1:][no - 1], size)
    label = np.empty(n, dtype=np.int64)
    label[order] = no
    _, pairs, _ = overlap(s, None, OverlapOptions(differences=1, indels=True, want_pairs=True, no_matrix=True))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]].astype(np.int64)
    assert pairs.shape[0] == info["edges"]
    assert np.array_equal(label[pairs[:, 0]], label[pairs[:, 1]])
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    g = coo_matrix((np.ones(pairs.shape[0], np.int8), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    ncomp, _ = connected_components(g, directed=False)
    assert ncomp == info["clusters"]
    # first row of every cluster is its smallest index (the reference seeds clusters in index order)
    firsts = order[np.concatenate(([0], np.nonzero(np.diff(no))[0] + 1))]
    mins = np.full(info["clusters"] + 1, n, dtype=np.int64)
    np.minimum.at(mins, label, np.arange(n))
    assert np.array_equal(firsts, mins[1:])

    s.rep = (s.rep % 3).astype(np.uint32)
    s.n_reps = 3
    lead, cnt, merged = dedup(s, OverlapOptions())
    leaders = lead == np.arange(n)
    assert merged == n - int(leaders.sum()) and merged > 0
    assert int(cnt.sum()) == int(s.count.sum()) and np.all(cnt[~leaders] == 0)
    assert np.all(lead <= np.arange(n)) and np.all(lead[lead] == lead)
    assert np.array_equal(s.rep[lead], s.rep) and np.array_equal(s.v_gene[lead], s.v_gene)
    assert np.array_equal(np.diff(s.offsets)[lead], np.diff(s.offsets))
