"""Shared-memory matrix tile (warp-aggregated atomics) vs plain global atomics on a high-match
workload: self-comparison of a low-complexity set in few repertoires.  usage: tile_ab.py [n_per_rep]"""
import sys, json
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
for reps, d, indels in ((2, 1, True), (8, 1, False), (32, 1, True), (64, 1, True), (100, 1, True)):
    s = synth.small_dense_set(7, reps, n * 2 // reps, max_len=7)
    for flags in (0, 1):
        with Engine(OverlapOptions(differences=d, indels=indels, flags=flags), n_reps_a=s.n_reps) as eng:
            db = eng.upload(s); eng.build_b(db)
            best = None
            for _ in range(3):
                eng.clear_matrix(); eng.run(db); st = eng.stats()
                best = st if best is None or st["ms_probe"] < best["ms_probe"] else best
            print(json.dumps({"reps": reps, "d": d, "indels": indels, "tile": flags == 0, "n": s.n, "matches": best["matches"],
                              "ms_probe": round(best["ms_probe"], 3), "G_matches_s": round(best["matches"] / best["ms_probe"] / 1e6, 2),
                              "matrix_sum": float(eng.matrix().sum())}), flush=True)
