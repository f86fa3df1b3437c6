"""Pair-test rate of the d>=3 path (brute_kernel): C2-like sets, with and without -g."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from compairr_b200 import Engine, OverlapOptions, synth
n_reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s = synth.make_set(1, n_reps, 10000)
L = s.lengths
for g in (True, False):
    if g:
        cnt = np.bincount(L); pairs = float((cnt.astype(np.float64) ** 2).sum())
    else:
        key = (L.astype(np.int64) << 32) | (s.v_gene.astype(np.int64) << 16) | s.j_gene
        _, c = np.unique(key, return_counts=True); pairs = float((c.astype(np.float64) ** 2).sum())
    with Engine(OverlapOptions(differences=3, ignore_genes=g), n_reps_a=s.n_reps) as eng:
        db = eng.upload(s); eng.build_b(db)
        best = None
        for _ in range(3):
            eng.clear_matrix(); eng.run(db); st = eng.stats()
            best = st if best is None or st["ms_total_run"] < best["ms_total_run"] else best
        print(json.dumps({"ignore_genes": g, "n": s.n, "pair_tests": pairs, "ms_probe": round(best["ms_probe"], 2), "ms_total": round(best["ms_total_run"], 2),
                          "T_pairs_s": round(pairs / best["ms_probe"] / 1e9, 3), "matches": best["matches"], "launches": best["kernel_launches"]}), flush=True)
