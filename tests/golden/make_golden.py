"""Generates tests/golden/golden.json by running the UNMODIFIED reference binary
(oracle/_ref/compairr, built from /root/reference/src by oracle/Makefile) in this container.

Inputs: the reference's own fixtures (test/seta.tsv, setb.tsv, setc.tsv — copied as DATA into
tests/golden/ref_*.tsv so the GPU box, which has no /root/reference, can replay them) and small
seeded synthetic sets written to tests/golden/syn_*.tsv.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from compairr_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF_TEST = "/root/reference/test"


def main():
    assert orc.have_reference(), "build oracle/_ref/compairr first (make -C oracle ref)"
    for f in ("seta", "setb", "setc"):
        shutil.copy(os.path.join(REF_TEST, f + ".tsv"), os.path.join(HERE, f"ref_{f}.tsv"))
    synth.small_dense_set(101, 3, 60).write_tsv(os.path.join(HERE, "syn_a.tsv"), id_prefix="a")
    synth.small_dense_set(102, 4, 60).write_tsv(os.path.join(HERE, "syn_b.tsv"), id_prefix="b")
    synth.small_dense_set(103, 1, 50).write_tsv(os.path.join(HERE, "syn_q.tsv"), id_prefix="q")
    synth.small_dense_set(104, 2, 70, alphabet="ACG", max_len=12, nucleotides=True).write_tsv(os.path.join(HERE, "syn_na.tsv"), id_prefix="a")
    synth.small_dense_set(105, 3, 70, alphabet="ACG", max_len=12, nucleotides=True).write_tsv(os.path.join(HERE, "syn_nb.tsv"), id_prefix="b")

    cases = []

    def add(name, args, files, pairs=False, expect_rc=0):
        argv = list(args) + [os.path.join(HERE, f) for f in files]
        out = os.path.join("/tmp", "golden_out.tsv")
        pr = os.path.join("/tmp", "golden_pairs.tsv")
        full = argv + ["-o", out, "-l", "/dev/null"] + (["-p", pr] if pairs else [])
        r = orc.run_reference(full)
        case = {"name": name, "args": list(args), "files": list(files), "pairs": pairs, "rc": r.returncode}
        assert r.returncode == expect_rc, (name, r.returncode, r.stderr)
        if r.returncode == 0:
            case["output"] = open(out).read()
            if pairs:
                lines = open(pr).read().splitlines()
                case["pairs_header"] = lines[0]
                case["pairs_sorted"] = sorted(lines[1:])
        else:
            case["stderr"] = r.stderr
        cases.append(case)

    AB = ["ref_seta.tsv", "ref_setb.tsv"]
    for d in ("0", "1", "2", "3"):
        add(f"ref_ab_d{d}", ["-m", "-d", d], AB)
    add("ref_ab_d1_i", ["-m", "-d", "1", "-i"], AB)                      # test/test.sh:9 -> expected.tsv
    add("ref_ab_d1_pairs", ["-m", "-d", "1"], AB, pairs=True)             # README.md:332-457
    for s in ("ratio", "min", "max", "mean"):
        add(f"ref_ab_d1_{s}", ["-m", "-d", "1", "-s", s], AB)
    add("ref_ab_n_d1_g_f", ["-m", "-n", "-d", "1", "-g", "-f"], AB)
    add("ref_ab_n_d3_g", ["-m", "-n", "-d", "3", "-g"], AB)
    for s in ("product", "ratio", "min", "MH", "Jaccard"):
        add(f"ref_b_self_{s}", ["-m", "-d", "0", "-s", s], ["ref_setb.tsv"])
    add("ref_b_self_MH_f", ["-m", "-s", "MH", "-f"], ["ref_setb.tsv"])
    add("ref_b_self_Jaccard_f", ["-m", "-s", "Jaccard", "-f"], ["ref_setb.tsv"])
    add("ref_x_cb_d1_a", ["-x", "-d", "1", "-a"], ["ref_setc.tsv", "ref_setb.tsv"])
    add("ref_x_cb_d1_f_pairs", ["-x", "-d", "1", "-f"], ["ref_setc.tsv", "ref_setb.tsv"], pairs=True)  # README.md:466-577
    add("ref_err_mh_d1", ["-m", "-d", "1", "-s", "MH"], AB, expect_rc=1)
    add("ref_err_x_multi", ["-x"], ["ref_setb.tsv", "ref_setb.tsv"], expect_rc=1)
    add("ref_err_indels_d2", ["-m", "-d", "2", "-i"], AB, expect_rc=1)

    SAB = ["syn_a.tsv", "syn_b.tsv"]
    for d, extra in [("0", []), ("1", []), ("1", ["-i"]), ("2", []), ("3", []), ("4", [])]:
        for g in ([], ["-g"]):
            tag = f"syn_ab_d{d}{'_i' if extra else ''}{'_g' if g else ''}"
            add(tag, ["-m", "-d", d, "-a"] + extra + g, SAB, pairs=(d in ("1", "3")))
    for s in ("ratio", "min", "max", "mean"):
        add(f"syn_ab_d1_i_{s}", ["-m", "-d", "1", "-i", "-s", s], SAB)
    add("syn_ab_d2_f", ["-m", "-d", "2", "-f"], SAB)
    add("syn_a_self_d1", ["-m", "-d", "1"], ["syn_a.tsv"])
    add("syn_a_self_MH", ["-m", "-s", "MH"], ["syn_a.tsv"])
    add("syn_a_self_Jaccard", ["-m", "-s", "Jaccard"], ["syn_a.tsv"])
    add("syn_ab_MH", ["-m", "-s", "MH"], SAB)
    add("syn_x_qb_d1_i", ["-x", "-d", "1", "-i"], ["syn_q.tsv", "syn_b.tsv"], pairs=True)
    add("syn_x_qb_d2_a", ["-x", "-d", "2", "-a"], ["syn_q.tsv", "syn_b.tsv"])
    add("syn_x_qb_d3", ["-x", "-d", "3"], ["syn_q.tsv", "syn_b.tsv"])
    NAB = ["syn_na.tsv", "syn_nb.tsv"]
    for d, extra in [("0", []), ("1", ["-i"]), ("2", []), ("3", [])]:
        add(f"syn_nt_d{d}", ["-m", "-n", "-d", d] + extra, NAB, pairs=(d == "2"))
    add("syn_nt_d1_dist", ["-m", "-n", "-d", "1", "-i", "--distance"], NAB, pairs=True)

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"reference": "CompAIRR 1.13.0 (oracle/_ref/compairr)", "cases": cases}, f, indent=1)
    print(len(cases), "cases written")


def main_cz():
    """tests/golden/golden_cz.json: `-c` (cluster) and `-z` (deduplicate) outputs of the reference
    binary.  Both commands take ONE file and print rows in an order the algorithm defines
    (src/cluster.cc:356-446, src/dedup.cc:185-190), so the whole output is compared byte for byte."""
    assert orc.have_reference(), "build oracle/_ref/compairr first (make -C oracle ref)"
    synth.small_dense_set(106, 2, 150, alphabet="ACDEFGHIK", min_len=3, max_len=6).write_tsv(os.path.join(HERE, "syn_c.tsv"), id_prefix="c")
    synth.small_dense_set(107, 3, 120, alphabet="ACGT", min_len=4, max_len=7, nucleotides=True).write_tsv(os.path.join(HERE, "syn_nc.tsv"), id_prefix="n")
    cases = []

    def add(name, args, file):
        out = os.path.join("/tmp", "golden_out.tsv")
        log = os.path.join("/tmp", "golden_log.txt")
        r = orc.run_reference(list(args) + [os.path.join(HERE, file), "-o", out, "-l", log])
        assert r.returncode == 0, (name, r.stderr)
        keep = [ln for ln in open(log).read().splitlines() if ln.startswith(("Clusters:", "Duplicates merged:"))]
        cases.append({"name": name, "args": list(args), "file": file, "output": open(out).read(), "log": keep})

    for f, nt in (("syn_a.tsv", []), ("syn_b.tsv", []), ("syn_c.tsv", []), ("ref_setb.tsv", []), ("syn_nc.tsv", ["-n"])):
        tag = f.split(".")[0]
        for d, extra in [("0", []), ("1", []), ("1", ["-i"]), ("2", []), ("3", [])]:
            for g in ([], ["-g"]):
                add(f"c_{tag}_d{d}{'_i' if extra else ''}{'_g' if g else ''}", ["-c", "-d", d] + extra + g + nt, f)
        for opt in ([], ["-g"], ["-f"], ["-g", "-f"]):
            add(f"z_{tag}{''.join('_' + o[1] for o in opt)}", ["-z"] + opt + nt, f)
    add("c_syn_c_d1_t3", ["-c", "-d", "1", "-t", "3"], "syn_c.tsv")
    with open(os.path.join(HERE, "golden_cz.json"), "w") as f:
        json.dump({"reference": "CompAIRR 1.13.0 (oracle/_ref/compairr)", "cases": cases}, f, indent=1)
    print(len(cases), "cluster / dedup cases written")


if __name__ == "__main__":
    if sys.argv[1:] == ["cz"]:
        main_cz()
    else:
        main()
