"""Full-size set B (10^8) vs 2*10^6 seeds, d=1 -i: the HBM-resident Bloom regime, for ncu."""
import sys
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth
bpk = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
pool = synth.make_pool(5, 4_000_000)
b = synth.make_set(3, 1000, 100000, pool=pool, indel_mutants=True, workers=14)
a = synth.make_set(2, 20, 100000, pool=pool, indel_mutants=True, workers=14)
with Engine(OverlapOptions(differences=1, indels=True, bloom_bits_per_key=bpk), n_reps_a=a.n_reps) as eng:
    db = eng.upload(b); eng.build_b(db); da = eng.upload(a)
    for _ in range(2):
        eng.run(da); s = eng.stats()
        print(s["probes"], "probes", round(s["ms_probe"], 3), "ms", round(s["probes"] / s["ms_probe"] / 1e6, 2), "Gprobes/s")
