"""Wall-clock breakdown of the end-to-end step (host buffers in, matrix out) at full size."""
import sys, time, json, os
sys.path.insert(0, ".")
import numpy as np, torch
from compairr_b200 import Engine, OverlapOptions, NarrowSet, synth
pool = synth.make_pool(5, 4_000_000)
b = synth.make_set(3, 1000, 100000, pool=pool, indel_mutants=True, workers=14)
a = synth.make_set(2, 100, 100000, pool=pool, indel_mutants=True, workers=14)
keep = []
def pin_narrow(s):
    ns = NarrowSet.from_seqset(s)
    for f in ("residues", "lengths", "v_gene", "j_gene", "rep", "count"):
        t = torch.from_numpy(np.ascontiguousarray(getattr(ns, f))).pin_memory(); keep.append(t); setattr(ns, f, t.numpy())
    return ns
bn, an = pin_narrow(b), pin_narrow(a)
print("bytes B", bn.nbytes(), "A", an.nbytes())
# raw H2D bandwidth reference
t = torch.empty(bn.residues.nbytes, dtype=torch.uint8, device="cuda"); src = keep[0]
torch.cuda.synchronize(); t0 = time.perf_counter(); t.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("torch pinned H2D GB/s", round(src.numel() / dt / 1e9, 1))
del t
eng = Engine(OverlapOptions(differences=1, indels=True), n_reps_a=a.n_reps)
for it in range(4):
    t0 = time.perf_counter(); eng.set_b(bn); t1 = time.perf_counter(); sb = eng.stats()
    eng.clear_matrix(); eng.run_a(an); t2 = time.perf_counter(); sa = eng.stats()
    m = eng.matrix(); t3 = time.perf_counter()
    print(json.dumps({"set_b_ms": round(1e3 * (t1 - t0), 1), "pipeline_ms": round(sb["ms_build_b"], 1), "dups_ms": round(sb["ms_dups_b"], 1),
                      "run_a_ms": round(1e3 * (t2 - t1), 1), "a_upload_ms": round(sa["ms_hash_a"], 1), "probe_ms": round(sa["ms_probe"], 1),
                      "matrix_ms": round(1e3 * (t3 - t2), 2)}), flush=True)
