"""Run-to-run determinism: the same 2e6 x 3e6 overlap three times with and without the filters (flag 2), each against
the oracle (the build races for slots, the results must not depend on who wins)."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from compairr_b200 import Engine, OverlapOptions, synth
from oracle import oracle as orc
pool = synth.make_pool(5, 400000)
a = synth.make_set(2, 20, 100000, pool=pool); b = synth.make_set(3, 30, 100000, pool=pool)
mo, _, io = orc.overlap(a, b, differences=1, threads=14)
print("oracle matches", io["matches"], "sum", mo.sum())
for flags in (0, 2):
    for rep in range(3):
        with Engine(OverlapOptions(differences=1, flags=flags), n_reps_a=a.n_reps) as eng:
            db = eng.upload(b); eng.build_b(db); da = eng.upload(a)
            eng.run(da); s = eng.stats(); m = eng.matrix()
            print("flags", flags, "rep", rep, "matches", s["matches"], "sum", m.sum(), "equal", np.array_equal(m, mo), "dups_b", eng.dups_b(), "absdiff", np.abs(m - mo).sum())
