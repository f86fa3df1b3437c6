"""Where the wall time of the CLI goes: COMPAIRR_B200_TRACE marks of `compairr_b200 -m -d 1 -i`
on files of the bench generator.  usage: cli_trace.py [reps_a reps_b] [--ref]"""
import os, subprocess, sys, time, shutil, tempfile
sys.path.insert(0, ".")
from compairr_b200 import synth
from oracle import oracle as orc
args = [x for x in sys.argv[1:] if not x.startswith("--")]
ra, rb = (int(args[0]), int(args[1])) if len(args) >= 2 else (10, 1000)
pool = synth.make_pool(5, 4_000_000)
tmp = tempfile.mkdtemp(prefix="cli_trace_", dir="/dev/shm")
try:
    t = time.time()
    a = synth.make_set(2, ra, 100000, pool=pool, indel_mutants=True, workers=14)
    b = synth.make_set(3, rb, 100000, pool=pool, indel_mutants=True, workers=14)
    fa, fb = f"{tmp}/a.tsv", f"{tmp}/b.tsv"
    orc.write_tsv(a, fa, "a"); orc.write_tsv(b, fb, "b")
    print(f"generated + wrote {a.n} + {b.n} sequences in {time.time()-t:.1f} s; files {os.path.getsize(fa)>>20} + {os.path.getsize(fb)>>20} MiB", flush=True)
    del a, b
    cli = "compairr_b200/bin/compairr_b200"
    env = dict(os.environ, COMPAIRR_B200_TRACE="1")
    for run in range(2):
        t = time.time()
        r = subprocess.run([cli, "-m", fa, fb, "-d", "1", "-i", "-o", f"{tmp}/o.tsv", "-l", f"{tmp}/l.txt"], env=env, capture_output=True, text=True)
        print(f"--- run {run}: wall {time.time()-t:.3f} s rc={r.returncode}\n{r.stderr}", flush=True)
        print("".join(l for l in open(f"{tmp}/l.txt") if "100%" in l), flush=True)
    if "--ref" in sys.argv:
        t = time.time()
        r = orc.run_reference(["-m", fa, fb, "-d", "1", "-i", "-t", str(os.cpu_count()), "-o", f"{tmp}/r.tsv", "-l", f"{tmp}/rl.txt"], timeout=3000)
        print(f"--- reference -t {os.cpu_count()}: wall {time.time()-t:.3f} s; identical output: {open(f'{tmp}/r.tsv','rb').read() == open(f'{tmp}/o.tsv','rb').read()}")
        print("".join(l for l in open(f"{tmp}/rl.txt") if "100%" in l), flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
