"""The C-ABI library loads and exports every symbol include/compairr_b200.h declares (no GPU)."""
import ctypes
import os
import re

from _util import ROOT
from compairr_b200 import cabi


def _declared():
    text = open(os.path.join(ROOT, "include", "compairr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = ctypes.CDLL(cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 24
    bound = {n for n, _, _ in cabi.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in bound, f"{n} declared in the header but missing from the ctypes binding"
    assert bound <= set(names)


def test_struct_layouts_match_header():
    # sizes are fixed by the header's field lists (LP64)
    assert ctypes.sizeof(cabi.cb_config) == 80
    assert ctypes.sizeof(cabi.cb_set) == 72
    assert ctypes.sizeof(cabi.cb_set_cols) == 16 + 6 * 16 + 16
    assert ctypes.sizeof(cabi.cb_stats) == 96


def test_abi_version_and_no_cpu_fallback():
    assert cabi.lib.cb_abi_version() == cabi.ABI_VERSION
    if cabi.lib.cb_device_count() == 0:
        # without a GPU the engine must refuse loudly, never compute on the CPU
        from compairr_b200 import Engine, OverlapOptions
        import pytest
        with pytest.raises(Exception) as e:
            Engine(OverlapOptions())
        assert "no CPU fallback" in str(e.value)


def test_probe_count_closed_form():
    import numpy as np
    from compairr_b200.engine import probe_count
    from compairr_b200.seqset import encode_sequences
    for seq, want in [("C", (1, 20, 59, 20)), ("CASSF", (1, 96, 215, 3706)), ("AAAAA", (1, 96, 212, 3706)),
                      ("CAASSF", (1, 115, 253, 5530)), ("CASSLRVGGYGYTF", (1, 267, 565, 33118))]:
        codes = encode_sequences([seq])[0]
        got = tuple(probe_count(codes, 20, d, i) for d, i in [(0, False), (1, False), (1, True), (2, False)])
        assert got == want
