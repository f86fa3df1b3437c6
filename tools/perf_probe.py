"""Ad-hoc performance probe (not the bench contract): runs a few configurations, prints stats."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth

def run(name, a, b, reps=3, **kw):
    opts = OverlapOptions(**kw)
    with Engine(opts, n_reps_a=max(a.n_reps, 1)) as eng:
        t0 = time.time(); db = eng.upload(b if b is not None else a); t_up = time.time() - t0
        eng.build_b(db); sb = eng.stats()
        da = db if b is None else eng.upload(a)
        best = None
        for _ in range(reps):
            eng.clear_matrix()
            eng.run(da); s = eng.stats()
            if best is None or s["ms_probe"] < best["ms_probe"]: best = s
        gps = best["probes"] / best["ms_probe"] / 1e6
        print(json.dumps({"case": name, "nA": a.n, "nB": (b or a).n, "probes": best["probes"], "ms_probe": round(best["ms_probe"], 3),
              "Gprobes_s": round(gps, 2), "bloom_pass_pct": round(100.0 * best["bloom_pass"] / max(best["probes"], 1), 3),
              "matches": best["matches"], "ms_hash_b": round(sb["ms_hash_b"], 3), "ms_build_b": round(sb["ms_build_b"], 3),
              "ms_dups_b": round(sb["ms_dups_b"], 3), "bloom_MiB": sb["bloom_bytes"] / 2**20, "slots": sb["table_slots"], "upload_s": round(t_up, 3)}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    if which in ("c2", "all"):
        s = synth.make_set(1, 100, 10000)
        for d, ind in [(0, False), (1, False), (1, True)]:
            run(f"C2 self d={d} indels={ind}", s, None, differences=d, indels=ind)
        sub = s.slice(0, 100000)
        run("C2 100k seeds vs 1M d=2", sub, s, differences=2)
    if which in ("mid", "all"):
        pool = synth.make_pool(5, 400000)
        a = synth.make_set(2, 100, 100000, pool=pool)
        b = synth.make_set(3, 100, 100000, pool=pool)
        for bpk in (10, 16, 24):
            run(f"1e7x1e7 d=1 bpk={bpk}", a, b, differences=1, bloom_bits_per_key=bpk)
        run("1e7x1e7 d=1 -i", a, b, differences=1, indels=True)
        run("1e6 of A x1e7 d=2", a.slice(0, 1000000), b, reps=1, differences=2)
