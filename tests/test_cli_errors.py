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


@pytest.mark.parametrize("case", ERR, ids=[c["name"] for c in ERR])
def test_error_cases_match_reference_with_parallel_reader(case, tmp_path):
    """The same malformed inputs read by 4 threads over ranges of ~100 bytes: the reported line is
    still the earliest bad line, numbered from the top of the file."""
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    env = dict(os.environ, COMPAIRR_B200_READ_MIN_BYTES="100")
    r = subprocess.run([CLI] + case["args"] + files + ["-t", "4", "-o", str(tmp_path / "o.tsv"), "-l", os.devnull],
                       capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == case["rc"] == 1
    assert r.stderr == case["stderr"]


def test_parallel_reader_reports_earliest_bad_line(tmp_path):
    """Two malformed lines far apart in a file read by 8 threads: the first one wins, with its
    absolute line number; comment lines before the header count."""
    lines = ["# comment", "@ another", "repertoire_id\tsequence_id\tduplicate_count\tv_call\tj_call\tjunction_aa"]
    for i in range(4000):
        lines.append(f"R{i % 7}\ts{i}\t{i % 9 + 1}\tV{i % 5}\tJ{i % 3}\tCASS{'ADEF'[i % 4]}YEQYF")
    lines[3 + 2500] = "R1\tsX\t0\tV1\tJ1\tCASSF"       # illegal duplicate_count on line 2504
    lines[3 + 3900] = "R1\tsY\t3\tV1\tJ1\tCAS1F"       # illegal character later on
    f = tmp_path / "bad.tsv"
    f.write_text("\n".join(lines) + "\n")
    env = dict(os.environ, COMPAIRR_B200_READ_MIN_BYTES="1000")
    r = subprocess.run([CLI, "-m", str(f), "-t", "8", "-o", str(tmp_path / "o.tsv")], capture_output=True, text=True,
                       timeout=60, env=env)
    assert r.returncode == 1
    assert r.stderr.endswith("\n\nError: Illegal duplicate_count on line 2504: 0\n")


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
