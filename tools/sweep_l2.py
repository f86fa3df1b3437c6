"""Bloom stage alone (CB_BLOOM_LOAD=9 skips the table probe): throughput vs Bloom size at fixed B."""
import sys, json, os
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth
pool = synth.make_pool(5, 4_000_000)
b = synth.make_set(3, 1000, 100000, pool=pool, indel_mutants=True, workers=14)
a = synth.make_set(2, 40, 100000, pool=pool, indel_mutants=True, workers=14)
for bpk in (0.5, 1, 1.5, 2, 2.5, 3, 4, 5, 6, 8, 10, 16):
    with Engine(OverlapOptions(differences=1, indels=True, bloom_bits_per_key=bpk), n_reps_a=a.n_reps) as eng:
        db = eng.upload(b); eng.build_b(db); sb = eng.stats(); da = eng.upload(a)
        best = None
        for _ in range(2):
            eng.run(da); s = eng.stats()
            if best is None or s["ms_probe"] < best["ms_probe"]: best = s
        print(json.dumps({"bpk": bpk, "bloom_MiB": round(sb["bloom_bytes"] / 2**20, 1), "ms": round(best["ms_probe"], 2),
                          "Gprobes_s": round(best["probes"] / best["ms_probe"] / 1e6, 1), "pass_pct": round(100 * best["bloom_pass"] / best["probes"], 2)}), flush=True)
