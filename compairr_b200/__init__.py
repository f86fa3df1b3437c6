"""compairr_b200 — B200-native engine for the repertoire-overlap hot path of CompAIRR.

Layout (only what the path needs):
  csrc/            CUDA kernels (sm_100a), the C ABI (include/compairr_b200.h) and the
                   CompAIRR-compatible C++ CLI
  cabi.py          ctypes binding of the C ABI (fails loudly if the library is not built)
  engine.py        host-side mirror of the reference's overlap() seam on numpy arrays
  seqset.py        sequence sets in structure-of-arrays form, AIRR TSV <-> arrays
  synth.py         seeded synthetic repertoires

The engine names (Engine, overlap, ...) load libcompairr_b200.so when first touched; the data
modules (seqset, synth, report) are plain numpy and import without it, so that a process which
only generates inputs — the reference arm of bench.py — never maps the CUDA library."""
from .seqset import SeqSet, NarrowSet, encode_sequences, AA_ALPHABET, NT_ALPHABET  # noqa: F401

_ENGINE_NAMES = ("Engine", "OverlapOptions", "overlap", "dedup", "cluster", "SCORES", "EngineError")
__all__ = ["SeqSet", "NarrowSet", "encode_sequences", *_ENGINE_NAMES]


def __getattr__(name):
    if name in _ENGINE_NAMES or name in ("engine", "cabi"):
        import importlib
        mod = importlib.import_module(".engine" if name != "cabi" else ".cabi", __name__)
        return mod if name in ("engine", "cabi") else getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
