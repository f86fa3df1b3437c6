"""The C3-sized sets (B = 10^8, A = 10^7 of the bench generator), cached in /dev/shm so that several
processes of one gpurun call share one generation."""
import os
import numpy as np
from compairr_b200 import synth
from compairr_b200.seqset import SeqSet
F = ("residues", "offsets", "v_gene", "j_gene", "rep", "count")
_pool = None


def _cached(name, make):
    d = f"/dev/shm/cbig_{name}"
    if os.path.isdir(d):
        arr = {f: np.load(f"{d}/{f}.npy") for f in F}
        return SeqSet(arr["residues"], arr["offsets"], arr["v_gene"], arr["j_gene"], arr["rep"], arr["count"], int(arr["rep"].max()) + 1)
    s = make()
    os.makedirs(d)
    for f in F:
        np.save(f"{d}/{f}.npy", getattr(s, f))
    return s


def _mk(seed, reps):
    global _pool
    _pool = _pool if _pool is not None else synth.make_pool(5, 4_000_000)
    return synth.make_set(seed, reps, 100000, pool=_pool, indel_mutants=True, workers=14)


def sets():
    b = _cached("b", lambda: _mk(3, 1000))
    a = _cached("a", lambda: _mk(2, 100))
    return a, b
