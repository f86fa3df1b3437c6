"""GPU parity against the reference's golden outputs, two ways: through the compiled CLI
(file in -> file out, the reference's outer contract) and through the C ABI from Python."""
import os
import subprocess

import pytest

from _util import CLI, GOLDEN_DIR, assert_matrix_text, golden_cases, hot_opts, is_integer_score, parse_args
from compairr_b200 import OverlapOptions, overlap, report
from compairr_b200.seqset import read_airr_pair

pytestmark = pytest.mark.gpu
CASES = [c for c in golden_cases() if c["rc"] == 0]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cli_reproduces_reference_files(case, tmp_path):
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    out, pairs = tmp_path / "out.tsv", tmp_path / "pairs.tsv"
    cmd = [CLI] + case["args"] + files + ["-o", str(out), "-l", str(tmp_path / "log.txt")]
    if case["pairs"]:
        cmd += ["-p", str(pairs)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + open(tmp_path / "log.txt").read()
    o = parse_args(case["args"])
    assert_matrix_text(out.read_text(), case["output"], exact=is_integer_score(o))
    if case["pairs"]:
        lines = pairs.read_text().splitlines()
        assert lines[0] == case["pairs_header"]
        assert sorted(lines[1:]) == case["pairs_sorted"]


SPLIT = CASES[::3]


@pytest.mark.parametrize("case", SPLIT, ids=[c["name"] for c in SPLIT])
def test_cli_parallel_reader_reproduces_reference_files(case, tmp_path):
    """Same golden cases with the input cut into ranges of a few hundred bytes parsed by 5 reader
    threads: sequence order, id numbering and therefore every output byte must not change."""
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    out, pairs = tmp_path / "out.tsv", tmp_path / "pairs.tsv"
    cmd = [CLI] + case["args"] + files + ["-t", "5", "-o", str(out), "-l", str(tmp_path / "log.txt")]
    if case["pairs"]:
        cmd += ["-p", str(pairs)]
    env = dict(os.environ, COMPAIRR_B200_READ_MIN_BYTES="200")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr + open(tmp_path / "log.txt").read()
    o = parse_args(case["args"])
    assert_matrix_text(out.read_text(), case["output"], exact=is_integer_score(o))
    if case["pairs"]:
        lines = pairs.read_text().splitlines()
        assert lines[0] == case["pairs_header"]
        assert sorted(lines[1:]) == case["pairs_sorted"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cabi_reproduces_reference_files(case):
    o = parse_args(case["args"])
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    a, b = read_airr_pair(files[0], files[1] if len(files) > 1 else None, o["nucleotides"])
    m, pairs, _ = overlap(a, b, OverlapOptions(nucleotides=o["nucleotides"], want_pairs=case["pairs"], **hot_opts(o)))
    text = report.format_matrix(m, a, b or a, o["score"], o["existence"], o["alternative"])
    assert_matrix_text(text, case["output"], exact=is_integer_score(o))
    if case["pairs"]:
        header, rows = report.format_pairs(pairs, a, b or a, o["distance"])
        assert header == case["pairs_header"]
        assert sorted(rows) == case["pairs_sorted"]


def test_cli_multi_gpu_flag_single_device(tmp_path):
    """--gpus larger than the box has is clamped; results unchanged."""
    case = next(c for c in CASES if c["name"] == "syn_ab_d1_i")
    files = [os.path.join(GOLDEN_DIR, f) for f in case["files"]]
    out = tmp_path / "out.tsv"
    r = subprocess.run([CLI] + case["args"] + files + ["-o", str(out), "-l", os.devnull, "--gpus", "2"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == case["output"]


def test_cli_write_failure_is_not_success(tmp_path):
    """A full disk must not look like success: /dev/full accepts the open and fails
    every flush; the command has to end with a non-zero status and say so."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    files = [os.path.join(GOLDEN_DIR, f) for f in ("ref_seta.tsv", "ref_setb.tsv")]
    r = subprocess.run([CLI, "-m", "-d", "1"] + files + ["-o", "/dev/full", "-l", str(tmp_path / "log.txt")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "Unable to write" in r.stderr
    r = subprocess.run([CLI, "-m", "-d", "1", "-t", "1"] + files + ["-o", str(tmp_path / "o.tsv"), "-l", str(tmp_path / "log.txt")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Threads (t):       1" in (tmp_path / "log.txt").read_text()
