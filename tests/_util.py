"""Shared helpers of the test-suite: golden cases, option parsing, comparisons."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN_DIR = os.path.join(HERE, "golden")
CLI = os.path.join(ROOT, "compairr_b200", "bin", "compairr_b200")


def golden_cases():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)["cases"]


def parse_args(args):
    """compairr argv fragment -> dict of hot-path options"""
    o = dict(differences=0, indels=False, ignore_genes=False, ignore_counts=False, score="product",
             existence=False, nucleotides=False, alternative=False, distance=False)
    it = iter(args)
    for a in it:
        if a == "-d":
            o["differences"] = int(next(it))
        elif a == "-s":
            o["score"] = next(it).lower()
        elif a == "-i":
            o["indels"] = True
        elif a == "-g":
            o["ignore_genes"] = True
        elif a == "-f":
            o["ignore_counts"] = True
        elif a == "-n":
            o["nucleotides"] = True
        elif a == "-x":
            o["existence"] = True
        elif a == "-a":
            o["alternative"] = True
        elif a == "--distance":
            o["distance"] = True
        elif a == "-m":
            pass
        else:
            raise ValueError(a)
    return o


def hot_opts(o):
    return {k: o[k] for k in ("differences", "indels", "ignore_genes", "ignore_counts", "score", "existence")}


def is_integer_score(o):
    return o["ignore_counts"] or o["score"] in ("product", "min", "max", "mean")


def assert_matrix_text(got: str, want: str, exact: bool):
    if exact:
        assert got == want
        return
    gl, wl = got.splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    for g, w in zip(gl, wl):
        gt, wt = g.split("\t"), w.split("\t")
        assert len(gt) == len(wt)
        for x, y in zip(gt, wt):
            if x == y:
                continue
            # the printed value has 10 significant digits; 1e-12 relative is far inside that
            np.testing.assert_allclose(float(x), float(y), rtol=1e-9, atol=0)


def pair_set(p):
    return sorted(map(tuple, np.asarray(p).tolist()))
