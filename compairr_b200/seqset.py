"""Sequence sets in structure-of-arrays form — what the reference's db_read() (src/db.cc:708-901)
leaves in memory, minus the C strings: one residue arena, offsets, gene / repertoire numbers and
duplicate counts.  Residue codes follow src/db.cc:33-71 (ACDEFGHIKLMNPQRSTVWY -> 0..19,
ACGT/U -> 0..3, case-insensitive)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

AA_ALPHABET = "ACDEFGHIKLMNPQRSTVWY"
NT_ALPHABET = "ACGT"


def _code_table(nucleotides: bool) -> np.ndarray:
    t = np.full(256, 255, dtype=np.uint8)
    alpha = NT_ALPHABET if nucleotides else AA_ALPHABET
    for i, ch in enumerate(alpha):
        t[ord(ch)] = i
        t[ord(ch.lower())] = i
    if nucleotides:
        t[ord("U")] = t[ord("u")] = 3
    return t


def encode_sequences(seqs: Sequence[str], nucleotides: bool = False):
    """strings -> (residue arena uint8, offsets uint64[n+1]); raises on an illegal symbol."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    raw = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8)
    res = _code_table(nucleotides)[raw]
    if res.size and res.max() == 255:
        bad = int(np.argmax(res == 255))
        raise ValueError(f"illegal character {chr(raw[bad])!r} in sequence")
    return np.ascontiguousarray(res), offsets


@dataclass
class SeqSet:
    residues: np.ndarray            # uint8, codes 0..sigma-1
    offsets: np.ndarray             # uint64, n+1
    v_gene: np.ndarray              # uint32
    j_gene: np.ndarray              # uint32
    rep: np.ndarray                 # uint32, repertoire number 0..n_reps-1
    count: np.ndarray               # uint64, duplicate_count >= 1
    n_reps: int
    nucleotides: bool = False
    rep_names: Optional[List[str]] = None
    v_names: Optional[List[str]] = None
    j_names: Optional[List[str]] = None
    seq_ids: Optional[List[str]] = None
    index_base: int = 0
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        self.residues = np.ascontiguousarray(self.residues, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(self.offsets, dtype=np.uint64)
        self.v_gene = np.ascontiguousarray(self.v_gene, dtype=np.uint32)
        self.j_gene = np.ascontiguousarray(self.j_gene, dtype=np.uint32)
        self.rep = np.ascontiguousarray(self.rep, dtype=np.uint32)
        self.count = np.ascontiguousarray(self.count, dtype=np.uint64)
        n = self.n
        assert self.offsets.shape == (n + 1,)
        assert self.v_gene.shape == self.j_gene.shape == self.rep.shape == self.count.shape == (n,)

    @property
    def n(self) -> int:
        return int(self.offsets.shape[0] - 1)

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets).astype(np.int64)

    @property
    def sigma(self) -> int:
        return 4 if self.nucleotides else 20

    def sequence(self, i: int) -> str:
        alpha = NT_ALPHABET if self.nucleotides else AA_ALPHABET
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return "".join(alpha[c] for c in self.residues[a:b])

    def slice(self, first: int, count: int) -> "SeqSet":
        """A shard sharing the residue arena (offsets stay absolute), reporting global indices."""
        sl = slice(first, first + count)
        return SeqSet(self.residues, self.offsets[first:first + count + 1], self.v_gene[sl],
                      self.j_gene[sl], self.rep[sl], self.count[sl], self.n_reps, self.nucleotides,
                      self.rep_names, self.v_names, self.j_names,
                      None if self.seq_ids is None else self.seq_ids[sl],
                      index_base=self.index_base + first)

    @staticmethod
    def from_records(records, nucleotides=False, gene_maps=None) -> "SeqSet":
        """records: iterable of (repertoire_id, sequence_id, count, v_call, j_call, sequence).
        Numbering is first-seen order like the reference (db.cc:510-520, 592-631); gene_maps =
        (v_map, j_map) dicts shared between the two sets of a comparison (db.cc:119-125)."""
        v_map, j_map = gene_maps if gene_maps is not None else ({}, {})
        rep_map: Dict[str, int] = {}
        reps, ids, cnt, vs, js, seqs = [], [], [], [], [], []
        for rid, sid, c, v, j, s in records:
            reps.append(rep_map.setdefault(rid, len(rep_map)))
            ids.append(sid)
            cnt.append(int(c))
            vs.append(v_map.setdefault(v, len(v_map)))
            js.append(j_map.setdefault(j, len(j_map)))
            seqs.append(s)
        res, off = encode_sequences(seqs, nucleotides)
        return SeqSet(res, off, np.array(vs, np.uint32), np.array(js, np.uint32),
                      np.array(reps, np.uint32), np.array(cnt, np.uint64), len(rep_map), nucleotides,
                      list(rep_map), None, None, ids)

    # ---- AIRR TSV (the reference's file contract, README.md "Input files") ------------------
    def names(self):
        rn = self.rep_names or [f"R{r:04d}" for r in range(self.n_reps)]
        nv = int(self.v_gene.max()) + 1 if self.n else 0
        nj = int(self.j_gene.max()) + 1 if self.n else 0
        vn = self.v_names or [f"TRBV{v + 1:02d}" for v in range(nv)]
        jn = self.j_names or [f"TRBJ{j + 1:02d}" for j in range(nj)]
        return rn, vn, jn

    def write_tsv(self, path: str, id_prefix: str = "s") -> None:
        alpha = np.frombuffer((NT_ALPHABET if self.nucleotides else AA_ALPHABET).encode(), np.uint8)
        text = alpha[self.residues].tobytes().decode("ascii")
        rn, vn, jn = self.names()
        col = "junction" if self.nucleotides else "junction_aa"
        off = self.offsets
        with open(path, "w") as f:
            f.write(f"repertoire_id\tsequence_id\tduplicate_count\tv_call\tj_call\t{col}\n")
            lines = []
            for i in range(self.n):
                sid = self.seq_ids[i] if self.seq_ids is not None else f"{id_prefix}{i + self.index_base}"
                lines.append(f"{rn[self.rep[i]]}\t{sid}\t{self.count[i]}\t{vn[self.v_gene[i]]}\t"
                             f"{jn[self.j_gene[i]]}\t{text[int(off[i]):int(off[i + 1])]}\n")
                if len(lines) >= 65536:
                    f.write("".join(lines))
                    lines = []
            f.write("".join(lines))


def _smallest_uint(a: np.ndarray, choices=(np.uint8, np.uint16, np.uint32, np.uint64)):
    mx = int(a.max()) if a.size else 0
    for t in choices:
        if mx <= np.iinfo(t).max:
            return np.ascontiguousarray(a, dtype=t)
    raise ValueError("value out of range")


@dataclass
class NarrowSet:
    """The same set with lengths instead of offsets and the smallest lossless column types
    (cb_set_cols): fewer bytes across PCIe.  Sequences must be contiguous in `residues`."""
    residues: np.ndarray
    lengths: np.ndarray
    v_gene: np.ndarray
    j_gene: np.ndarray
    rep: np.ndarray
    count: np.ndarray
    n_reps: int
    longest: int
    index_base: int = 0

    @property
    def n(self) -> int:
        return int(self.lengths.shape[0])

    @staticmethod
    def from_seqset(s: "SeqSet") -> "NarrowSet":
        off = s.offsets
        lens = np.diff(off)
        a, b = int(off[0]), int(off[-1])
        return NarrowSet(np.ascontiguousarray(s.residues[a:b]), _smallest_uint(lens, (np.uint8, np.uint16, np.uint32)),
                         _smallest_uint(s.v_gene, (np.uint8, np.uint16, np.uint32)),
                         _smallest_uint(s.j_gene, (np.uint8, np.uint16, np.uint32)),
                         _smallest_uint(s.rep, (np.uint8, np.uint16, np.uint32)),
                         _smallest_uint(s.count), s.n_reps, int(lens.max()) if lens.size else 0, s.index_base)

    def nbytes(self) -> int:
        return sum(x.nbytes for x in (self.residues, self.lengths, self.v_gene, self.j_gene, self.rep, self.count))


def read_airr_tsv(path: str, nucleotides: bool = False, gene_maps=None, default_rep: str = "1",
                  cdr3: bool = False) -> SeqSet:
    """Minimal Python mirror of the reference reader (src/db.cc:172-901) for tests and tools:
    header-driven columns, ids numbered in first-seen order.  The C++ CLI has the full reader."""
    col = ("cdr3" if nucleotides else "cdr3_aa") if cdr3 else ("junction" if nucleotides else "junction_aa")
    recs = []
    with open(path) as f:
        header = None
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if header is None:
                if line.startswith("#") or line.startswith("@"):
                    continue
                header = {name: i for i, name in enumerate(line.split("\t"))}
                continue
            t = line.split("\t")

            def get(name, default=""):
                i = header.get(name)
                return t[i] if i is not None and i < len(t) else default
            recs.append((get("repertoire_id", default_rep) if "repertoire_id" in header else default_rep,
                         get("sequence_id"), int(get("duplicate_count", "1") or 1), get("v_call"),
                         get("j_call"), get(col)))
    return SeqSet.from_records(recs, nucleotides, gene_maps)


def read_airr_pair(path_a: str, path_b=None, nucleotides: bool = False):
    """Both sets of a comparison with the shared V/J gene numbering (db.cc:119-125).
    path_b None or equal to path_a is the self-comparison: returns (a, None)."""
    maps = ({}, {})
    a = read_airr_tsv(path_a, nucleotides, maps, default_rep="1")
    b = None
    if path_b is not None and path_b != path_a:
        b = read_airr_tsv(path_b, nucleotides, maps, default_rep="2")
    for s in (a, b):
        if s is not None:
            s.v_names, s.j_names = list(maps[0]), list(maps[1])
    return a, b
