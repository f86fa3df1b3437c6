"""Full-size set B (10^8) vs a slice of set A: the kernels alone, for quick rates and for ncu.
usage: prof_big.py [bits_per_key [differences [seed_repertoires_of_100k]]]   (d = 1 runs with -i)"""
import sys
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth
bpk = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else (20 if d == 1 else 2)
pool = synth.make_pool(5, 4_000_000)
b = synth.make_set(3, 1000, 100000, pool=pool, indel_mutants=True, workers=14)
a = synth.make_set(2, reps, 100000, pool=pool, indel_mutants=True, workers=14)
for dd in ([1, 2] if d == 12 else [d]):
    with Engine(OverlapOptions(differences=dd, indels=dd == 1, bloom_bits_per_key=bpk), n_reps_a=a.n_reps) as eng:
        db = eng.upload(b); eng.build_b(db); da = eng.upload(a)
        n = a.n if dd == 1 else min(a.n, 200000)
        for _ in range(2):
            eng.run(da, 0, n); s = eng.stats()
            print(f"d={dd}", s["probes"], "probes", round(s["ms_probe"], 3), "ms", round(s["probes"] / s["ms_probe"] / 1e6, 2), "Gprobes/s",
                  "pass", round(s["bloom_pass"] / s["probes"], 5), "build ms", round(eng.stats()["ms_build_b"], 2), flush=True)
