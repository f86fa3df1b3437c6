"""Result formatting — the reference's output contract (src/overlap.cc:540-577 values,
:944-1039 matrix layouts, :455-507 and :908-925 pairs file), as text, from a raw matrix.
Host-side; used by the Python API and by the parity tests to compare against golden files."""
from __future__ import annotations

import numpy as np

from .seqset import AA_ALPHABET, SeqSet


def _fmt(x: float) -> str:
    return "%.10g" % x   # C's %.10lg


def rep_totals(s: SeqSet):
    cnt = np.zeros(s.n_reps, dtype=np.uint64)
    np.add.at(cnt, s.rep, s.count)
    sq = np.zeros(s.n_reps, dtype=np.float64)
    np.add.at(sq, s.rep, (s.count * s.count).astype(np.float64))  # u64 product then double (overlap.cc:654)
    return cnt, sq


def finalize(matrix: np.ndarray, a: SeqSet, b: SeqSet, score: str) -> np.ndarray:
    """MH / Jaccard post-processing of the summed products / minima (overlap.cc:548-570)."""
    score = score.lower()
    if score not in ("mh", "jaccard"):
        return matrix
    ca, qa = rep_totals(a)
    cb, qb = rep_totals(b)
    ca_f, cb_f = ca.astype(np.float64), cb.astype(np.float64)
    if score == "mh":
        lx = qa / ca_f / ca_f
        ly = qb / cb_f / cb_f
        xy = 1.0 * ca_f[:, None] * cb_f[None, :]
        return (2.0 * matrix) / ((lx[:, None] + ly[None, :]) * xy)
    return matrix / (ca_f[:, None] + cb_f[None, :] - matrix)


def format_matrix(matrix: np.ndarray, a: SeqSet, b: SeqSet, score="product", existence=False,
                  alternative=False) -> str:
    b_names, _, _ = b.names()
    a_names, _, _ = a.names()
    cols = sorted(range(b.n_reps), key=lambda t: b_names[t].encode())      # strcmp order
    if existence:
        rows = list(range(a.n))
        row_names = [a.seq_ids[i] for i in rows]
        vals = matrix
    else:
        rows = sorted(range(a.n_reps), key=lambda s: a_names[s].encode())
        row_names = [a_names[s] for s in rows]
        vals = finalize(matrix, a, b, score)
    out = []
    if alternative:
        out.append("#sequence_id_1\trepertoire_id_2\tmatches" if existence else "#repertoire_id_1\trepertoire_id_2\tmatches")
        for r, rn in zip(rows, row_names):
            for t in cols:
                out.append(f"{rn}\t{b_names[t]}\t{_fmt(vals[r, t])}")
    else:
        out.append("#" + "".join("\t" + b_names[t] for t in cols))
        for r, rn in zip(rows, row_names):
            out.append(rn + "".join("\t" + _fmt(vals[r, t]) for t in cols))
    return "\n".join(out) + "\n"


def format_pairs(pairs: np.ndarray, a: SeqSet, b: SeqSet, distance=False):
    """-> (header line, list of row lines) of the pairs file."""
    col = "junction" if a.nucleotides else "junction_aa"
    header = (f"#repertoire_id_1\tsequence_id_1\tduplicate_count_1\tv_call_1\tj_call_1\t{col}_1"
              f"\trepertoire_id_2\tsequence_id_2\tduplicate_count_2\tv_call_2\tj_call_2\t{col}_2")
    if distance:
        header += "\tdistance"
    alpha = "acgt" if a.nucleotides else AA_ALPHABET   # db.cc:73-74: nucleotides print lower-case

    def side(s: SeqSet, i: int, names):
        rn, vn, jn = names
        seq = "".join(alpha[c] for c in s.residues[int(s.offsets[i]):int(s.offsets[i + 1])])
        sid = s.seq_ids[i] if s.seq_ids is not None else ""
        return f"{rn[s.rep[i]]}\t{sid}\t{s.count[i]}\t{vn[s.v_gene[i]]}\t{jn[s.j_gene[i]]}\t{seq}"
    na, nb = a.names(), b.names()
    rows = []
    for x, y in np.asarray(pairs).tolist():
        line = side(a, x, na) + "\t" + side(b, y, nb)
        if distance:
            la, lb = int(a.offsets[x + 1] - a.offsets[x]), int(b.offsets[y + 1] - b.offsets[y])
            d = 1
            if la == lb:
                d = int(np.count_nonzero(a.residues[int(a.offsets[x]):int(a.offsets[x + 1])] !=
                                         b.residues[int(b.offsets[y]):int(b.offsets[y + 1])]))
            line += f"\t{d}"
        rows.append(line)
    return header, rows


def _alpha(s: SeqSet):
    return "acgt" if s.nucleotides else AA_ALPHABET


def _seq_text(s: SeqSet, i: int) -> str:
    al = _alpha(s)
    return "".join(al[c] for c in s.residues[int(s.offsets[i]):int(s.offsets[i + 1])])


def format_clusters(order, cluster_no, cluster_size, s: SeqSet) -> str:
    """The `-c` output file (src/cluster.cc:417-446)."""
    rn, vn, jn = s.names()
    col = "junction" if s.nucleotides else "junction_aa"
    out = [f"#cluster_no\tcluster_size\trepertoire_id\tsequence_id\tduplicate_count\tv_call\tj_call\t{col}"]
    for i, no, size in zip(np.asarray(order).tolist(), np.asarray(cluster_no).tolist(), np.asarray(cluster_size).tolist()):
        sid = s.seq_ids[i] if s.seq_ids is not None else ""
        out.append(f"{no}\t{size}\t{rn[s.rep[i]]}\t{sid}\t{s.count[i]}\t{vn[s.v_gene[i]]}\t{jn[s.j_gene[i]]}\t{_seq_text(s, i)}")
    return "\n".join(out) + "\n"


def format_dedup(leader, count, s: SeqSet, ignore_genes=False) -> str:
    """The `-z` output file (src/dedup.cc:27-59, 169-173, 185-190): one row per group, at its
    first member, in file order."""
    rn, vn, jn = s.names()
    col = "junction" if s.nucleotides else "junction_aa"
    out = ["repertoire_id\tduplicate_count" + ("" if ignore_genes else "\tv_call\tj_call") + f"\t{col}"]
    leader = np.asarray(leader)
    for i in np.nonzero(leader == np.arange(leader.size))[0].tolist():
        genes = "" if ignore_genes else f"\t{vn[s.v_gene[i]]}\t{jn[s.j_gene[i]]}"
        out.append(f"{rn[s.rep[i]]}\t{int(count[i])}{genes}\t{_seq_text(s, i)}")
    return "\n".join(out) + "\n"
