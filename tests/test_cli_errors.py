"""CLI contract that needs no GPU: option grammar, error texts and exit codes of the reference
(src/compairr.cc:561-689, src/overlap.cc:699-703), checked against golden stderr captured from
the unmodified reference binary."""
import os
import subprocess

import pytest

from _util import CLI, GOLDEN_DIR, golden_cases

ERR = [c for c in golden_cases() if c["rc"] != 0]


def _run(args):
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=60)


@pytest.mark.parametrize("case", ERR, ids=[c["name"] for c in ERR])
def test_error_cases_match_reference(case, tmp_path):
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    r = _run(case["args"] + files + ["-o", str(tmp_path / "o.tsv"), "-l", os.devnull])
    assert r.returncode == case["rc"] == 1
    assert r.stderr == case["stderr"]


@pytest.mark.parametrize("args,msg", [
    ([], "Please specify a command (--help, --version, --matrix, --existence, --cluster, or --deduplicate)"),
    (["-m", "-x", "a", "b"], "Please specify just one command (--help, --version, --matrix, --existence, --cluster, or --deduplicate)"),
    (["-m"], "Incorrect number of arguments. One or two input files must be specified."),
    (["-x", "a"], "Incorrect number of arguments. Two input files must be specified."),
    (["-m", "a", "-d", "-1"], "Differences specified with -d or -differences cannot be negative."),
    (["-m", "a", "-i"], "Indels are only allowed when d=1"),
    (["-m", "a", "-s", "foo"], "Argument to -s or --score must be MH, Jaccard, product, ratio, min, max or mean"),
    (["-x", "a", "b", "-s", "MH"], "The Morisita-Horn index is only allowed when computing repertoire overlap"),
    (["-m", "a", "-d", "1", "-s", "jaccard"], "The Jaccard index is not defined when d>0"),
    (["-m", "a", "-k", "x"], "Option --keep-columns only allowed with --pairs options."),
])
def test_option_errors(args, msg):
    r = _run(args)
    assert r.returncode == 1
    assert r.stderr == f"\nError: {msg}\n"


def test_option_twice_and_threads():
    r = _run(["-m", "-d", "1", "-d", "2", "a"])
    assert r.returncode == 1 and r.stderr == "Error: Option -d or --differences specified more than once.\n"
    r = _run(["-m", "a", "-t", "0"])
    assert r.returncode == 1 and "Illegal number of threads" in r.stderr


def test_version_and_help():
    assert _run(["-v"]).returncode == 0
    r = _run(["-h"])
    assert r.returncode == 0 and "Usage: compairr [OPTIONS] TSVFILE1 [TSVFILE2]" in r.stderr
