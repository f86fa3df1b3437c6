"""Seeded synthetic repertoires: CDR3-like amino-acid junctions
"CAS" + random + "F" with length ~ N(14.5, 1.8) clipped to [8, 22], 60 V x 13 J genes,
Pareto(1.2) duplicate counts, 20 % of every repertoire drawn from a shared public pool and another
20 % as 1-2 edit mutants of pool members, so that d=0/1/2 overlaps are non-trivial.
Nucleotide sets encode every amino acid by a random codon (length x 3).

Vectorised numpy, generated in blocks so that 10^8-sequence sets stay within host memory."""
from __future__ import annotations

import numpy as np

from .seqset import SeqSet

_CODE = {ch: i for i, ch in enumerate("ACDEFGHIKLMNPQRSTVWY")}
N_V, N_J = 60, 13

# one codon table: amino-acid code -> list of codons (as 3 nt codes A0 C1 G2 T3)
_CODONS = {
    "A": ["GCT", "GCC", "GCA", "GCG"], "C": ["TGT", "TGC"], "D": ["GAT", "GAC"], "E": ["GAA", "GAG"],
    "F": ["TTT", "TTC"], "G": ["GGT", "GGC", "GGA", "GGG"], "H": ["CAT", "CAC"],
    "I": ["ATT", "ATC", "ATA"], "K": ["AAA", "AAG"], "L": ["TTA", "TTG", "CTT", "CTC", "CTA", "CTG"],
    "M": ["ATG"], "N": ["AAT", "AAC"], "P": ["CCT", "CCC", "CCA", "CCG"], "Q": ["CAA", "CAG"],
    "R": ["CGT", "CGC", "CGA", "CGG", "AGA", "AGG"], "S": ["TCT", "TCC", "TCA", "TCG", "AGT", "AGC"],
    "T": ["ACT", "ACC", "ACA", "ACG"], "V": ["GTT", "GTC", "GTA", "GTG"], "W": ["TGG"], "Y": ["TAT", "TAC"],
}


def _lengths(rng, n):
    return np.clip(np.rint(rng.normal(14.5, 1.8, n)), 8, 22).astype(np.int64)


def _fresh(rng, lens):
    """random CDR3-like sequences for the given lengths -> (residues, offsets)"""
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    res = rng.integers(0, 20, int(off[-1]), dtype=np.uint8)
    res[off[:-1]] = _CODE["C"]
    res[off[:-1] + 1] = _CODE["A"]
    res[off[:-1] + 2] = _CODE["S"]
    res[off[1:] - 1] = _CODE["F"]
    return res, off


def _counts(rng, n):
    return np.maximum(np.floor(rng.pareto(1.2, n) + 1.0), 1).astype(np.uint64)


def _derive(rng, pool, src, kind, sigma=20):
    """sequences derived from pool members src: kind 0 copy, 1 one substitution, 2 two
    substitutions, 3 one deletion, 4 one insertion -> (residues, offsets)"""
    poff = pool["off"]
    slen = poff[src + 1] - poff[src]
    dlen = slen + (kind == 4) - (kind == 3)
    off = np.zeros(src.size + 1, dtype=np.int64)
    np.cumsum(dlen, out=off[1:])
    total = int(off[-1])
    seq_of = np.repeat(np.arange(src.size), dlen)
    pos = np.arange(total, dtype=np.int64) - off[seq_of]          # position in the derived sequence
    p1 = (rng.random(src.size) * np.maximum(np.where(kind == 4, slen + 1, slen), 1)).astype(np.int64)
    kk, pp = kind[seq_of], p1[seq_of]
    spos = pos + ((kk == 3) & (pos >= pp)) - ((kk == 4) & (pos > pp))
    spos = np.minimum(spos, slen[seq_of] - 1)
    res = pool["res"][poff[src][seq_of] + spos].copy()
    # substitutions: change the residue at p1 (and p2) to a different one
    def substitute(mask, where):
        idx = off[:-1][mask] + where[mask]
        res[idx] = (res[idx] + rng.integers(1, sigma, idx.size, dtype=np.uint8)) % sigma
    substitute((kind == 1) | (kind == 2), p1)
    p2 = (p1 + 1 + (rng.random(src.size) * np.maximum(slen - 1, 1)).astype(np.int64)) % np.maximum(slen, 1)
    substitute(kind == 2, p2)
    ins = kind == 4
    res[off[:-1][ins] + p1[ins]] = rng.integers(0, sigma, int(ins.sum()), dtype=np.uint8)
    return res, off


_POOLS = {}


def make_pool(seed: int, n: int):
    key = (seed, n)
    if key not in _POOLS:
        rng = np.random.default_rng([seed, 0x9001])
        lens = _lengths(rng, n)
        res, off = _fresh(rng, lens)
        _POOLS.clear()   # keep at most one pool alive per process
        _POOLS[key] = {"res": res, "off": off, "v": rng.integers(0, N_V, n, dtype=np.uint32),
                       "j": rng.integers(0, N_J, n, dtype=np.uint32), "key": key}
    return _POOLS[key]


def _make_block(args):
    """nr consecutive repertoires starting at r0 -> (residues, lengths, v, j, counts)"""
    seed, r0, nr, per_rep, pool_key, k_pool, k_mut, indel_mutants = args
    pool = make_pool(*pool_key)
    n_pool = pool["v"].size
    k_new = per_rep - k_pool - k_mut
    rng = np.random.default_rng([seed, r0])
    # pool members: without replacement inside a repertoire
    src_pool = np.concatenate([rng.choice(n_pool, k_pool, replace=False) for _ in range(nr)]) if k_pool else np.zeros(0, np.int64)
    src_mut = rng.integers(0, n_pool, nr * k_mut)
    kinds_allowed = np.array([1, 2, 3, 4] if indel_mutants else [1, 2])
    kind = np.concatenate([np.zeros(src_pool.size, np.int64), kinds_allowed[rng.integers(0, kinds_allowed.size, src_mut.size)]])
    src = np.concatenate([src_pool, src_mut]).astype(np.int64)
    dres, doff = _derive(rng, pool, src, kind)
    flen = _lengths(rng, nr * k_new)
    fres, foff = _fresh(rng, flen)
    # per repertoire: [pool | mutants | fresh]
    dl = np.diff(doff)
    lens, chunks, v, j = [], [], [], []
    for i in range(nr):
        p0, p1 = i * k_pool, (i + 1) * k_pool
        m0, m1 = nr * k_pool + i * k_mut, nr * k_pool + (i + 1) * k_mut
        f0, f1 = i * k_new, (i + 1) * k_new
        lens += [dl[p0:p1], dl[m0:m1], flen[f0:f1]]
        chunks += [dres[doff[p0]:doff[p1]], dres[doff[m0]:doff[m1]], fres[foff[f0]:foff[f1]]]
        v += [pool["v"][src_pool[p0:p1]], pool["v"][src_mut[i * k_mut:(i + 1) * k_mut]], rng.integers(0, N_V, k_new, dtype=np.uint32)]
        j += [pool["j"][src_pool[p0:p1]], pool["j"][src_mut[i * k_mut:(i + 1) * k_mut]], rng.integers(0, N_J, k_new, dtype=np.uint32)]
    return (np.concatenate(chunks), np.concatenate(lens).astype(np.int64), np.concatenate(v).astype(np.uint32),
            np.concatenate(j).astype(np.uint32), _counts(rng, nr * per_rep))


def make_set(seed: int, n_reps: int, per_rep: int, pool=None, pool_frac=0.2, mut_frac=0.2,
             indel_mutants=False, nucleotides=False, single_repertoire=False,
             block_reps: int = 8, workers: int = 1, first_rep: int = 0) -> SeqSet:
    """A set of n_reps repertoires with per_rep sequences each.  Repertoire r of the set is a
    pure function of (seed, first_rep + r, pool), so a shard of a bigger set can be generated on
    its own (first_rep must be a multiple of block_reps)."""
    if pool is None:
        pool = make_pool(seed ^ 0x5EED, max(int(per_rep * (pool_frac + mut_frac)) * 4, 16))
    n_pool = pool["v"].size
    k_pool = min(int(per_rep * pool_frac), n_pool)
    k_mut = int(per_rep * mut_frac)
    jobs = [(seed, first_rep + r0, min(block_reps, n_reps - r0), per_rep, pool["key"], k_pool, k_mut, indel_mutants)
            for r0 in range(0, n_reps, block_reps)]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pl:
            parts = pl.map(_make_block, jobs)
    else:
        parts = [_make_block(jb) for jb in jobs]
    lens = np.concatenate([p[1] for p in parts])
    res = np.concatenate([p[0] for p in parts])
    off = np.zeros(lens.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    rep = np.repeat(np.arange(n_reps, dtype=np.uint32), per_rep)
    if single_repertoire:
        rep[:] = 0
    s = SeqSet(res, off, np.concatenate([p[2] for p in parts]), np.concatenate([p[3] for p in parts]), rep,
               np.concatenate([p[4] for p in parts]), 1 if single_repertoire else n_reps)
    return to_nucleotides(s, seed) if nucleotides else s


def to_nucleotides(s: SeqSet, seed: int) -> SeqSet:
    """Back-translate with random codons (length x 3), nucleotide codes A0 C1 G2 T3."""
    nt = {"A": 0, "C": 1, "G": 2, "T": 3}
    aa = "ACDEFGHIKLMNPQRSTVWY"
    maxc = 6
    table = np.zeros((20, maxc, 3), dtype=np.uint8)
    ncod = np.zeros(20, dtype=np.int64)
    for a, ch in enumerate(aa):
        ncod[a] = len(_CODONS[ch])
        for c, cod in enumerate(_CODONS[ch]):
            table[a, c] = [nt[x] for x in cod]
    rng = np.random.default_rng([seed, 0xC0D0])
    pick = (rng.random(s.residues.size) * ncod[s.residues]).astype(np.int64)
    res = table[s.residues, pick].reshape(-1)
    return SeqSet(res, s.offsets * np.uint64(3), s.v_gene, s.j_gene, s.rep, s.count, s.n_reps,
                  nucleotides=True, rep_names=s.rep_names, seq_ids=s.seq_ids)


def small_dense_set(seed: int, n_reps: int, per_rep: int, alphabet="ACS", min_len=1, max_len=9,
                    n_v=3, n_j=2, nucleotides=False) -> SeqSet:
    """Low-complexity sets for parity tests: short sequences over a tiny alphabet, so runs of
    equal residues, indel neighbours and duplicates are dense."""
    rng = np.random.default_rng([seed, 0xDE5E])
    n = n_reps * per_rep
    codes = np.array([("ACGT" if nucleotides else "ACDEFGHIKLMNPQRSTVWY").index(c) for c in alphabet], np.uint8)
    lens = rng.integers(min_len, max_len + 1, n)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    res = codes[rng.integers(0, codes.size, int(off[-1]))]
    return SeqSet(res, off, rng.integers(0, n_v, n, dtype=np.uint32), rng.integers(0, n_j, n, dtype=np.uint32),
                  rng.permutation(np.repeat(np.arange(n_reps, dtype=np.uint32), per_rep)),
                  rng.integers(1, 6, n).astype(np.uint64), n_reps, nucleotides=nucleotides)
