"""compute-sanitizer target: every probe path at small size — d=0, d=1 (+indels), d=2, the any-length kernel (flag 16 and
really long sequences), d=3 on CUDA cores and on the tensor cores, pairs, existence, cluster, dedup.
usage: compute-sanitizer --tool memcheck python tools/memcheck_paths.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from compairr_b200 import OverlapOptions, cluster, dedup, overlap, synth
pool = synth.make_pool(301, 5000)
a = synth.make_set(302, 4, 1500, pool=pool, indel_mutants=True)
b = synth.make_set(303, 5, 1500, pool=pool, indel_mutants=True)
nt_pool = synth.make_pool(304, 3000)
an = synth.make_set(305, 2, 1000, pool=nt_pool, nucleotides=True)
bn = synth.make_set(306, 3, 1000, pool=nt_pool, nucleotides=True)
runs = [
    ("d0", a, b, dict(differences=0)),
    ("d1", a, b, dict(differences=1)),
    ("d1i pairs", a, b, dict(differences=1, indels=True, want_pairs=True)),
    ("d2 -g", a, b, dict(differences=2, ignore_genes=True)),
    ("d1i generic", a, b, dict(differences=1, indels=True, flags=16)),
    ("d2 generic", a.slice(0, 300), b, dict(differences=2, flags=16)),
    ("nt d1i", an, bn, dict(differences=1, indels=True, nucleotides=True)),
    ("nt d2", an.slice(0, 300), bn, dict(differences=2, nucleotides=True)),
    ("d3 cuda cores", a, b, dict(differences=3, flags=4)),
    ("d3 tensor cores -g", a, b, dict(differences=3, ignore_genes=True)),
    ("self d1i", a, None, dict(differences=1, indels=True)),
]
for name, x, y, kw in runs:
    m, p, info = overlap(x, y, OverlapOptions(**kw))
    print(name, "matches", info["run"]["matches"], "probes", info["run"]["probes"], flush=True)
q = a.slice(0, 2000)
q.rep = np.zeros(q.n, np.uint32)
q.n_reps = 1
m, _, info = overlap(q, b, OverlapOptions(differences=1, indels=True, existence=True))
print("existence", info["run"]["matches"], flush=True)
order, no, size, info = cluster(a, OverlapOptions(differences=1, indels=True))
print("cluster", info["clusters"], flush=True)
lead, cnt, merged = dedup(a, OverlapOptions())
print("dedup merged", merged, flush=True)
