"""Full-size set B (10^8), A shard 10^7: sweep Bloom bits/key; prints kernel and build timings."""
import sys, json, time
sys.path.insert(0, ".")
import numpy as np
from compairr_b200 import Engine, OverlapOptions, synth
pool = synth.make_pool(5, 4_000_000)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
b = synth.make_set(3, nb, 100000, pool=pool, indel_mutants=True, workers=14)
a = synth.make_set(2, 100, 100000, pool=pool, indel_mutants=True, workers=14)
for d, ind in [(1, True), (2, False)]:
    for bpk in [float(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "6,8,10,12,16,24".split(","))]:
        with Engine(OverlapOptions(differences=d, indels=ind, bloom_bits_per_key=bpk), n_reps_a=a.n_reps) as eng:
            db = eng.upload(b); eng.build_b(db); sb = eng.stats()
            da = eng.upload(a)
            n = a.n if d == 1 else 200000
            best = None
            for _ in range(2):
                eng.clear_matrix(); eng.run(da, 0, n); s = eng.stats()
                if best is None or s["ms_probe"] < best["ms_probe"]: best = s
            print(json.dumps({"d": d, "bpk": bpk, "bloom_MiB": round(sb["bloom_bytes"] / 2**20), "bloom2_MiB": round(sb["bloom2_bytes"] / 2**20), "ms_probe": round(best["ms_probe"], 2),
                              "Gprobes_s": round(best["probes"] / best["ms_probe"] / 1e6, 1), "pass_pct": round(100 * best["bloom_pass"] / best["probes"], 3),
                              "matches": best["matches"], "ms_hash_b": round(sb["ms_hash_b"], 2), "ms_build_b": round(sb["ms_build_b"], 2), "ms_dups_b": round(sb["ms_dups_b"], 2)}), flush=True)
