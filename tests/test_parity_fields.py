"""Host-side check of compairr_b200/csrc/common.cuh (the integer arithmetic shared by the kernels
and the engine): the position-class structure of the Zobrist values, the class-filter addressing, the
home-slot multiplier and the closed-form variant count.  Compiles tests/csrc/hd_check.cpp with g++; no GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_common_cuh_invariants(tmp_path):
    exe = tmp_path / "hd_check"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++17", "-x", "c++", os.path.join(HERE, "csrc", "hd_check.cpp"), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "hd_check ok" in r.stdout
