"""Wall time of the whole command, file in -> file out: our CLI next to the unmodified reference
binary (oracle/_ref/compairr, -t all cores) on the same synthetic AIRR TSV files; outputs compared
byte for byte.  usage: cli_wall.py [reps_a reps_b per_rep [extra compairr options...]]
With -c or -z among the options the command is the one-file cluster / deduplicate run on set A
(-z: set A folded into reps_a // 10 + 1 repertoires so that duplicates exist inside a repertoire)."""
import json, os, subprocess, sys, time, filecmp, re
sys.path.insert(0, ".")
from compairr_b200 import synth
ra = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rb = int(sys.argv[2]) if len(sys.argv) > 2 else 100
per = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
extra = sys.argv[4:] or ["-d", "1", "-i"]
one_file = "-c" in extra or "-z" in extra
tmp = os.environ.get("TMPDIR", "/tmp")
pool = synth.make_pool(5, max(1000, ra * per // 25))
t0 = time.time()
a = synth.make_set(2, ra, per, pool=pool, indel_mutants=True, workers=8)
b = synth.make_set(3, rb if not one_file else 1, per if not one_file else 10, pool=pool, indel_mutants=True, workers=8)
if "-z" in extra:
    import numpy as np
    a.n_reps = ra // 10 + 1
    a.rep = (a.rep % a.n_reps).astype(np.uint32)
fa, fb = os.path.join(tmp, "wall_a.tsv"), os.path.join(tmp, "wall_b.tsv")
a.write_tsv(fa, "a"); b.write_tsv(fb, "b")
print("generated + written in", round(time.time() - t0, 1), "s;", os.path.getsize(fa) + os.path.getsize(fb), "bytes", flush=True)
ncpu = os.cpu_count()
def run(exe, tag, threads):
    out, log = os.path.join(tmp, f"wall_{tag}.out"), os.path.join(tmp, f"wall_{tag}.log")
    t0 = time.perf_counter()
    r = subprocess.run([exe] + ([fa] if one_file else ["-m", fa, fb]) + ["-t", str(threads), "-o", out, "-l", log] + extra, capture_output=True, text=True, env=dict(os.environ, COMPAIRR_B200_TRACE="1"))
    if tag == "ours": print(r.stderr, flush=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr + open(log).read()
    phases = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"^([A-Za-z ]+):\s+100% \(([0-9.]+)s\)", open(log).read(), re.M)}
    return dt, out, phases
res = {}
for tag, exe in (("ours", "compairr_b200/bin/compairr_b200"), ("ref", "oracle/_ref/compairr")):
    dt, out, ph = run(exe, tag, ncpu)
    if tag == "ours":  # second run: the first pays CUDA context creation + page cache effects
        dt2, out, ph = run(exe, tag, ncpu)
        res["ours_first_wall_s"] = round(dt, 2); dt = dt2
    res[tag + "_wall_s"] = round(dt, 2); res[tag + "_phases"] = ph; res[tag + "_out"] = out
res["identical_output"] = filecmp.cmp(res.pop("ours_out"), res.pop("ref_out"), shallow=False)
res["speedup_wall"] = round(res["ref_wall_s"] / res["ours_wall_s"], 1)
res["config"] = {"reps_a": ra, "reps_b": rb, "per_rep": per, "options": extra, "host_threads": ncpu}
print(json.dumps(res, indent=1))
