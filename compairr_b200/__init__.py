"""compairr_b200 — B200-native engine for the repertoire-overlap hot path of CompAIRR.

Layout (only what the path needs):
  csrc/            CUDA kernels (sm_100a), the C ABI (include/compairr_b200.h) and the
                   CompAIRR-compatible C++ CLI
  cabi.py          ctypes binding of the C ABI (fails loudly if the library is not built)
  engine.py        host-side mirror of the reference's overlap() seam on numpy arrays
  seqset.py        sequence sets in structure-of-arrays form, AIRR TSV <-> arrays
  synth.py         seeded synthetic repertoires
"""
from .seqset import SeqSet, NarrowSet, encode_sequences, AA_ALPHABET, NT_ALPHABET  # noqa: F401
from .engine import Engine, OverlapOptions, overlap, dedup, cluster, SCORES  # noqa: F401

__all__ = ["SeqSet", "Engine", "OverlapOptions", "overlap", "dedup", "cluster", "SCORES", "encode_sequences"]
