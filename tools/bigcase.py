"""Full-size case (B = 10^8) with the sets cached in /dev/shm so several library builds can be
compared in one gpurun call.  usage: bigcase.py [d1|d2|both] [bpk]"""
import sys, json, os
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from compairr_b200 import Engine, OverlapOptions
import bigcase_sets
a, b = bigcase_sets.sets()
which = sys.argv[1] if len(sys.argv) > 1 else "both"
bpk = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0   # 0 = the engine's default
for d, ind in [(1, True), (2, False)]:
    if which not in ("both", f"d{d}"): continue
    with Engine(OverlapOptions(differences=d, indels=ind, bloom_bits_per_key=bpk), n_reps_a=a.n_reps) as eng:
        db = eng.upload(b); eng.build_b(db); sb = eng.stats(); da = eng.upload(a)
        n = a.n if d == 1 else 200000
        best = None
        for _ in range(2):
            eng.clear_matrix(); eng.run(da, 0, n); s = eng.stats()
            if best is None or s["ms_probe"] < best["ms_probe"]: best = s
        print(json.dumps({"tag": os.environ.get("TAG", ""), "d": d, "ms_probe": round(best["ms_probe"], 2), "Gprobes_s": round(best["probes"] / best["ms_probe"] / 1e6, 1),
                          "pass_pct": round(100 * best["bloom_pass"] / best["probes"], 2), "matches": best["matches"], "f1_MiB": sb["bloom_bytes"] >> 20, "f2_MiB": sb["bloom2_bytes"] >> 20,
                          "ms_build": round(sb["ms_build_b"], 1), "ms_dups": round(sb["ms_dups_b"], 1)}), flush=True)
