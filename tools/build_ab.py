"""Set-B build at full size (10^8 keys) under the three build routes — tiled (default), swept (flags 64),
direct (flags 8) — in one process: ms_build_b (best of 3), duplicate count and the matrix of a 10^6 set A
must agree.  usage: [COMPAIRR_B200_LIB=...] python tools/build_ab.py [flags ...]"""
import sys, json, os
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import numpy as np
from compairr_b200 import Engine, OverlapOptions
import bigcase_sets
a, b = bigcase_sets.sets()
a = a.slice(0, 1_000_000)
ref = None
for flags in [int(x) for x in sys.argv[1:]] or [0, 64, 8]:
    with Engine(OverlapOptions(differences=1, indels=True, flags=flags), n_reps_a=a.n_reps) as eng:
        db = eng.upload(b)
        ms = []
        for _ in range(3):
            eng.build_b(db)
            st = eng.stats()
            ms.append(round(st["ms_build_b"], 2))
        dups = eng.dups_b()
        eng.run(eng.upload(a))
        m = eng.matrix()
        got = (dups, float(m.sum()))
        ref = ref or (got, m)
        print(json.dumps({"tag": os.environ.get("TAG", ""), "flags": flags, "ms_build_b": ms, "ms_dups": round(st["ms_dups_b"], 2),
                          "launches": st["kernel_launches"], "dups": dups, "matrix_sum": got[1],
                          "same_as_first": got == ref[0] and bool(np.array_equal(m, ref[1]))}), flush=True)
