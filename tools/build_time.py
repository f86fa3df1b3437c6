import sys, json, os, time
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth
pool = synth.make_pool(5, 4_000_000)
b = synth.make_set(3, 1000, 100000, pool=pool, indel_mutants=True, workers=14)
with Engine(OverlapOptions(differences=1, indels=True), n_reps_a=100) as eng:
    db = eng.upload(b)
    for it in range(5):
        t0 = time.perf_counter(); eng.rehash(db); t1 = time.perf_counter(); eng.build_b(db); t2 = time.perf_counter(); s = eng.stats()
        print(json.dumps({"rehash_wall_ms": round(1e3*(t1-t0),1), "build_wall_ms": round(1e3*(t2-t1),1), "ms_build_b": round(s["ms_build_b"],1), "ms_dups_b": round(s["ms_dups_b"],1), "dups": eng.dups_b()}), flush=True)
