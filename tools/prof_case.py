"""One configuration, few launches — the command ncu wraps.  usage: prof_case.py d1|d1i|d2|d0|mid"""
import sys
sys.path.insert(0, ".")
from compairr_b200 import Engine, OverlapOptions, synth

which = sys.argv[1] if len(sys.argv) > 1 else "d1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if which == "mid":
    pool = synth.make_pool(5, 400000)
    a = synth.make_set(2, 100, 100000, pool=pool); b = synth.make_set(3, 100, 100000, pool=pool)
    opts = OverlapOptions(differences=1)
else:
    b = synth.make_set(1, 100, 10000); a = b
    opts = {"d0": OverlapOptions(differences=0), "d1": OverlapOptions(differences=1),
            "d1i": OverlapOptions(differences=1, indels=True), "d2": OverlapOptions(differences=2)}[which]
with Engine(opts, n_reps_a=a.n_reps) as eng:
    db = eng.upload(b); eng.build_b(db)
    da = db if a is b else eng.upload(a)
    n = 20000 if which == "d2" else a.n
    for _ in range(reps):
        eng.run(da, 0, n)
        s = eng.stats()
        print(which, s["probes"], "probes", round(s["ms_probe"], 3), "ms", round(s["probes"] / s["ms_probe"] / 1e6, 2), "Gprobes/s")
