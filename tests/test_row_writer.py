"""The CLI's multi-threaded row writer (pairs, cluster and dedup files) produces the bytes of a serial
loop for any thread count and block size.  Compiles tests/csrc/row_writer_check.cpp; no GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_row_writer_is_order_preserving(tmp_path):
    exe = tmp_path / "row_writer_check"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++17", os.path.join(HERE, "csrc", "row_writer_check.cpp"), "-o", str(exe), "-lpthread"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "row_writer_check ok" in r.stdout
