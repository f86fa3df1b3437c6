"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY — importable from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs, never from compairr_b200/.

  overlap(a, b, ...)        the plain-C restatement (overlap_oracle.c) through ctypes
  brute_force(a, b, ...)    a pure-numpy/Python statement of the DEFINITION,
                            hash-free, for small cases
  dedup(s, ...), cluster(s, ...)   restatements of src/dedup.cc and src/cluster.cc (`-z`, `-c`)
  run_reference(args)       the unmodified reference binary oracle/_ref/compairr (built from
                            /root/reference/src by oracle/Makefile) on TSV files
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "compairr")

SCORES = {"product": 0, "ratio": 1, "min": 2, "max": 3, "mean": 4, "mh": 5, "jaccard": 6}


class orc_set(C.Structure):
    _fields_ = [("n", C.c_uint64), ("residues", C.c_void_p), ("offsets", C.c_void_p),
                ("v_gene", C.c_void_p), ("j_gene", C.c_void_p), ("rep", C.c_void_p),
                ("count", C.c_void_p), ("n_reps", C.c_uint32)]


class orc_opts(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("alphabet_size", "differences", "indels", "ignore_genes",
                                         "ignore_counts", "score", "existence", "threads", "method",
                                         "want_pairs")]


class orc_result(C.Structure):
    _fields_ = [("probes", C.c_uint64), ("bloom_pass", C.c_uint64), ("matches", C.c_uint64),
                ("n_pairs", C.c_uint64), ("pairs", C.POINTER(C.c_uint64)),
                ("seconds_build", C.c_double), ("seconds_probe", C.c_double)]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "overlap_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_overlap.restype = C.c_int
        _lib.orc_overlap.argtypes = [C.POINTER(orc_set), C.POINTER(orc_set), C.POINTER(orc_opts),
                                     C.c_uint32, C.c_void_p, C.POINTER(orc_result)]
        _lib.orc_count_dups.restype = C.c_uint64
        _lib.orc_count_dups.argtypes = [C.POINTER(orc_set), C.c_int, C.c_int]
        _lib.orc_enumerate.restype = C.c_uint64
        _lib.orc_enumerate.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_uint32, C.c_uint64]
        _lib.orc_free.argtypes = [C.c_void_p]
        _lib.orc_write_tsv.restype = C.c_int
        _lib.orc_write_tsv.argtypes = [C.POINTER(orc_set), C.c_char_p, C.c_char_p, C.c_int, C.c_uint64]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _set(s) -> orc_set:
    o = orc_set()
    o.n = s.n
    o.residues, o.offsets = _p(s.residues), _p(s.offsets)
    o.v_gene, o.j_gene, o.rep, o.count = _p(s.v_gene), _p(s.j_gene), _p(s.rep), _p(s.count)
    o.n_reps = s.n_reps
    return o


def overlap(a, b=None, differences=0, indels=False, ignore_genes=False, ignore_counts=False,
            score="product", existence=False, threads=1, method=0, want_pairs=False,
            want_matrix=True):
    """-> (matrix [rows x b.n_reps] or None, pairs [n,2] or None, info)"""
    L = lib()
    sa = _set(a)
    sb = sa if b is None else _set(b)
    bb = a if b is None else b
    o = orc_opts(a.sigma, differences, int(indels), int(ignore_genes), int(ignore_counts),
                 SCORES[score.lower()], int(existence), threads, method, int(want_pairs))
    rows = a.n if existence else a.n_reps
    m = np.zeros((rows, bb.n_reps), dtype=np.float64) if want_matrix else None
    res = orc_result()
    rc = L.orc_overlap(C.byref(sa), C.byref(sb), C.byref(o), a.n_reps, _p(m) if m is not None else None,
                       C.byref(res))
    if rc:
        raise RuntimeError("oracle failed")
    pairs = None
    if want_pairs:
        pairs = np.ctypeslib.as_array(res.pairs, shape=(res.n_pairs, 2)).copy() if res.n_pairs else \
            np.zeros((0, 2), np.uint64)
    if res.n_pairs:
        L.orc_free(res.pairs)
    info = {"probes": res.probes, "bloom_pass": res.bloom_pass, "matches": res.matches,
            "seconds_build": res.seconds_build, "seconds_probe": res.seconds_probe}
    return m, pairs, info


def write_tsv(s, path: str, id_prefix: str = "s") -> None:
    """SeqSet -> AIRR TSV, the same bytes as SeqSet.write_tsv for sets with default names, ~100x
    faster (10^8-line bench inputs)."""
    assert s.rep_names is None and s.v_names is None and s.j_names is None and s.seq_ids is None
    ss = _set(s)
    if lib().orc_write_tsv(C.byref(ss), path.encode(), id_prefix.encode(), int(s.nucleotides), s.index_base):
        raise OSError(f"cannot write {path}")


def count_dups(s, ignore_genes=False) -> int:
    ss = _set(s)
    return int(lib().orc_count_dups(C.byref(ss), s.sigma, int(ignore_genes)))


def enumerate_variants(codes, sigma, differences, indels):
    """-> (records [n,5] = kind,pos1,res1,pos2,res2 ; list of variant sequences as tuples)"""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    L = lib()
    n = L.orc_enumerate(_p(codes), codes.size, sigma, differences, int(indels), None, None, 0, 0)
    recs = np.zeros((n, 5), dtype=np.uint32)
    stride = codes.size + 2
    seqs = np.zeros((n, stride), dtype=np.uint8)
    L.orc_enumerate(_p(codes), codes.size, sigma, differences, int(indels), _p(recs), _p(seqs), stride, n)
    return recs, [tuple(int(x) for x in row[1:1 + row[0]]) for row in seqs]


def _variant_bytes(codes, sigma, differences, indels):
    """the variants of one sequence, in enumeration order, as bytes objects"""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    L = lib()
    n = L.orc_enumerate(_p(codes), codes.size, sigma, differences, int(indels), None, None, 0, 0)
    stride = codes.size + 2
    seqs = np.zeros((n, stride), dtype=np.uint8)
    L.orc_enumerate(_p(codes), codes.size, sigma, differences, int(indels), None, _p(seqs), stride, n)
    out = [b""] * n
    for ln in np.unique(seqs[:, 0]).tolist():
        idx = np.nonzero(seqs[:, 0] == ln)[0]
        if ln == 0:
            continue
        rows = np.ascontiguousarray(seqs[idx, 1:1 + ln]).view(f"V{ln}").ravel().tolist()
        for k, r in zip(idx.tolist(), rows):
            out[k] = r
    return out


# ---- hash-free statement of the definition, small inputs only ----------

def _within(x, y, d, indels):
    if len(x) == len(y):
        return int(np.count_nonzero(x != y)) <= d
    if not indels or abs(len(x) - len(y)) != 1:
        return False
    lo, hi = (x, y) if len(x) < len(y) else (y, x)
    p = 0
    while p < len(lo) and lo[p] == hi[p]:
        p += 1
    return bool(np.array_equal(lo[p:], hi[p + 1:]))


def _score(name, ignore_counts, a, b):
    if ignore_counts:
        return 1.0
    name = name.lower()
    if name in ("product", "mh"):
        return float(a) * float(b)
    if name == "ratio":
        return float(a) / float(b)
    if name in ("min", "jaccard"):
        return float(min(a, b))
    if name == "max":
        return float(max(a, b))
    return (float(a) + float(b)) / 2


def brute_force(a, b=None, differences=0, indels=False, ignore_genes=False, ignore_counts=False,
                score="product", existence=False):
    b = a if b is None else b
    rows = a.n if existence else a.n_reps
    m = np.zeros((rows, b.n_reps))
    pairs = []
    sa = [a.residues[int(a.offsets[i]):int(a.offsets[i + 1])] for i in range(a.n)]
    sb = [b.residues[int(b.offsets[i]):int(b.offsets[i + 1])] for i in range(b.n)]
    for i in range(a.n):
        for k in range(b.n):
            if not ignore_genes and (a.v_gene[i] != b.v_gene[k] or a.j_gene[i] != b.j_gene[k]):
                continue
            if _within(sa[i], sb[k], differences, indels):
                m[i if existence else a.rep[i], b.rep[k]] += _score(score, ignore_counts, int(a.count[i]), int(b.count[k]))
                pairs.append((i, k))
    return m, np.array(pairs, dtype=np.uint64).reshape(-1, 2)


# ---- `-z` and `-c`: restatements of src/dedup.cc and src/cluster.cc, small inputs only ----------
# Pinned by tests/test_oracle_golden.py against tests/golden/golden_cz.json (outputs of the
# unmodified reference binary, byte for byte).  The reference's open-addressing table becomes a
# dict keyed by what its probe loop compares; a dict value lists indices in insertion (= index)
# order, which is the order a linear-probing chain filled in index order is walked in.

def _seq(s, i):
    return bytes(s.residues[int(s.offsets[i]):int(s.offsets[i + 1])])


def dedup(s, ignore_genes=False, ignore_counts=False):
    """dedup(), src/dedup.cc:139-215.  process() (:62-137) links every sequence to the latest
    earlier one with the same repertoire, V, J (unless -g) and residues; report() (:27-59) prints
    each chain once, at its first member, with the summed count (1 per member with -f).
    -> (leader index per sequence, summed count at each leader else 0, duplicates merged)"""
    last = {}
    lead = np.arange(s.n, dtype=np.uint32)
    cnt = np.zeros(s.n, dtype=np.uint64)
    merged = 0
    for i in range(s.n):
        key = (int(s.rep[i]), _seq(s, i)) if ignore_genes else \
              (int(s.rep[i]), int(s.v_gene[i]), int(s.j_gene[i]), _seq(s, i))
        if key in last:                      # dedup.cc:128-132: next_seq[last] = seed, returns true
            lead[i] = lead[last[key]]
            merged += 1
        last[key] = i
        cnt[lead[i]] += np.uint64(1 if ignore_counts else int(s.count[i]))   # report(), :34-43
    return lead, cnt, merged


def cluster(s, differences=0, indels=False, ignore_genes=False):
    """cluster(), src/cluster.cc:302-475.
    network (:71-274): the hits of seed x are, for each variant of x in generate_variants order
    (variants.cc:402-428), the table entries equal to that variant with equal V and J (unless -g)
    and index != x (:105); for d > 2, all other sequences within Hamming distance d in index order
    (process_trad, :147-196).  clustering (:277-300, 356-407): breadth-first from every still
    unclustered seed in index order, a cluster's members chained in the order they are reached;
    clusters sorted by decreasing size (:41-55, 411; glibc qsort = stable merge sort).
    -> (order, cluster_no (1-based), cluster_size) per output row, n_clusters, n_edges"""
    n = s.n
    seqs = [_seq(s, i) for i in range(n)]
    vj = [(0, 0) if ignore_genes else (int(s.v_gene[i]), int(s.j_gene[i])) for i in range(n)]
    network = []
    if differences <= 2:
        table = {}
        for i in range(n):
            table.setdefault((seqs[i], vj[i]), []).append(i)
        cache = {}
        for x in range(n):
            if seqs[x] not in cache:
                cache[seqs[x]] = _variant_bytes(np.frombuffer(seqs[x], np.uint8), s.sigma, differences, indels)
            hits = []
            for v in cache[seqs[x]]:
                hits.extend(h for h in table.get((v, vj[x]), ()) if h != x)
            network.append(hits)
    else:
        arr = [np.frombuffer(q, np.uint8) for q in seqs]
        for x in range(n):
            network.append([h for h in range(n) if h != x and vj[h] == vj[x] and len(seqs[h]) == len(seqs[x])
                            and int(np.count_nonzero(arr[h] != arr[x])) <= differences])
    cid = [-1] * n
    clusters = []
    for seed in range(n):
        if cid[seed] >= 0:
            continue
        members = [seed]
        cid[seed] = len(clusters)
        k = 0
        while k < len(members):
            for h in network[members[k]]:
                if cid[h] < 0:
                    cid[h] = len(clusters)
                    members.append(h)
            k += 1
        clusters.append(members)
    clusters.sort(key=lambda m: -len(m))     # list.sort is stable
    order = np.array([i for m in clusters for i in m], dtype=np.uint32)
    no = np.array([k + 1 for k, m in enumerate(clusters) for _ in m], dtype=np.uint32)
    size = np.array([len(m) for m in clusters for _ in m], dtype=np.uint32)
    return order, no, size, len(clusters), sum(len(h) for h in network)


# ---- the unmodified reference binary ---------------------------------------------------------

def have_reference() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_reference(args, cwd=None, timeout=600):
    """Run oracle/_ref/compairr with the given argument list -> CompletedProcess."""
    return subprocess.run([REF_BIN] + list(args), cwd=cwd, capture_output=True, text=True, timeout=timeout)


def parse_matrix_alt(path):
    """Three-column (-a) output -> dict {(row_id, col_id): float}"""
    out = {}
    with open(path) as f:
        next(f)
        for line in f:
            r, c, v = line.rstrip("\n").split("\t")
            out[(r, c)] = float(v)
    return out
