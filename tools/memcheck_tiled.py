"""compute-sanitizer target: word-wide hash loads and the tiled build at the smallest size that takes it
(4.4e6 keys -> 2^24 slots), then a short d=1 run.  usage: compute-sanitizer --tool memcheck python tools/memcheck_tiled.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from compairr_b200 import Engine, OverlapOptions, synth
pool = synth.make_pool(121, 400_000)
b = synth.make_set(122, 44, 100_000, pool=pool, indel_mutants=True, workers=4)
a = synth.make_set(123, 1, 2_000, pool=pool, indel_mutants=True)
with Engine(OverlapOptions(differences=1, indels=True), n_reps_a=a.n_reps) as eng:
    db = eng.upload(b)
    eng.build_b(db)
    st = eng.stats()
    eng.build_b(db)                      # rebuild in place: link reset + tiled build again
    eng.run(eng.upload(a))
    print("slots", st["table_slots"], "launches", st["kernel_launches"], "dups", eng.dups_b(), "matrix sum", float(eng.matrix().sum()))
