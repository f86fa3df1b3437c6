#!/usr/bin/env python
"""bench.py — measures BASELINE.json's metric (variant probes/s of the `-m` overlap hot path, d = 1
and d = 2, next to the reference's all-core CPU build) on synthetic repertoires of the C3/C4 shape,
one process per GPU.

  python bench.py [--gpus N --steps K --warmup W]                      our arm (N=1 by default)
  torchrun --nproc-per-node N ... bench.py --gpus N ...                 our arm, N ranks
  python bench.py --impl reference ...                                  the reference's CPU build

Headline workload ("C3", WEAK scaling): set B = 1000 repertoires x 100 000 AA CDR3 (10^8 sequences,
on every GPU), set A = 100 repertoires x 100 000 per GPU (rank r holds repertoires [100 r, 100 r +
100) of set A), `-m -d 1 -i` (substitutions + indels), score product.  A step is one whole hot-path
pass: hash B, build table + filters over B, count duplicates, hash the A shard, enumerate + probe +
verify + accumulate, and (N > 1) the all-reduce of the matrix (NCCL inside the library).

  value     probes of all ranks / max-over-ranks device time, inputs resident in HBM
  e2e       the same through the C ABI with pinned HOST buffers: every rank copies 1/N of set B and
            its A shard across PCIe, set B is all-gathered over NVLink (cb_set_b_sharded), the
            matrix is all-reduced and read back — all inside the timed region
  roofline  the enumeration + table kernels: 8 algorithmic bytes per probe (one filter word) / their
            CUDA-event duration, against the measured HBM peak (MEASURED_PEAKS.json)
  d2        C4 geometry: `-m -d 2 -g`, same set B, 2*10^5 seeds per GPU, kernel + whole-step rates
  strong    the FIXED C3 problem (set A = 1000 x 100 000 in total) split over the N GPUs
  c5        C5 geometry at a bounded size: nucleotide queries (one repertoire) against a nucleotide
            set, `-x -n -d 2 -p --no-matrix` (hash path, pairs drained to the host) and `-x -n -g -d 3`
            (length-bucketed one-hot int8 GEMM on tcgen05), probes/s and pair tests/s
  parity_checked  the engine at this N (sharded B, sharded A, all-reduce) against the UNMODIFIED
            reference binary on a sample of the same generator: matrices byte-identical
  cpu_baseline / cli_wall (N=1, rank 0)  oracle/_ref/compairr -t <all cores> on A = 10 repertoires
            vs the FULL set B, files in -> files out; our CLI on the same files; outputs compared
            with cmp; the reference's hot-path phases extrapolated to the arm's set A
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

POOL_SEED = 5
SEED_A, SEED_B = 2, 3
CLI = os.path.join(ROOT, "compairr_b200", "bin", "compairr_b200")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reps-b", type=int, default=1000)
    ap.add_argument("--reps-a-per-gpu", type=int, default=100)
    ap.add_argument("--per-rep", type=int, default=100_000)
    ap.add_argument("--differences", type=int, default=1)
    ap.add_argument("--no-indels", action="store_true")
    ap.add_argument("--bloom-bits", type=float, default=0.0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-d2", action="store_true")
    ap.add_argument("--skip-strong", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--c5-reps-b", type=int, default=100, help="C5 section: nucleotide set-B repertoires (x per-rep sequences)")
    ap.add_argument("--c5-queries", type=int, default=1_000_000, help="C5 section: nucleotide query sequences per GPU")
    ap.add_argument("--d2-seeds", type=int, default=200_000, help="d=2 section: seeds per GPU")
    ap.add_argument("--strong-reps-a", type=int, default=0, help="strong-scaling section: set-A repertoires in total (0 = reps-b)")
    ap.add_argument("--sample-reps-a", type=int, default=10, help="reference run: set-A repertoires (set B is complete)")
    ap.add_argument("--parity-reps", type=str, default="4x40", help="parity sample: AxB repertoires")
    ap.add_argument("--workers", type=int, default=0, help="generator processes (0 = auto)")
    ap.add_argument("--pool-n", type=int, default=4_000_000, help="size of the shared public pool")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic C3: B={a.reps_b}x{a.per_rep} AA, A={a.reps_a_per_gpu}x{a.per_rep} per GPU, public pool {a.pool_n}, "
            f"-m -d {a.differences}{'' if a.no_indels or a.differences != 1 else ' -i'} -s product")


def n_workers(a):
    return a.workers or max(1, min(32, (os.cpu_count() or 8) - 2))


def sum_probes(s, d, indels):
    """closed form; runs of equal residues only matter with indels"""
    L = s.lengths.astype(np.int64)
    sg = s.sigma
    n = np.ones_like(L)
    if d >= 1:
        n += (sg - 1) * L
        if indels:
            res = s.residues
            new_run = np.ones(res.size, dtype=bool)
            new_run[1:] = res[1:] != res[:-1]
            new_run[s.offsets[:-1].astype(np.int64)] = True
            runs = np.add.reduceat(new_run.astype(np.int64), s.offsets[:-1].astype(np.int64))
            n += np.where(L > 1, runs, 0) + sg * (L + 1) - L
    if d >= 2:
        n += (sg - 1) ** 2 * (L * (L - 1) // 2)
    return n.sum()


# ---- the reference's CPU build: files in -> files out --------------------------------------------------

PHASES = ("Computing hashes:", "Check duplicates:", "Hashing sequences:", "Analysing:")


def ref_phases(log_text):
    """Phase times of one reference run (src/util.cc:61-68), in log order: hashes of set 1, duplicate
    check of set 1, hashes of set 2, table build over set 2, analysis."""
    out = []
    for m in re.finditer(r"^(Computing hashes:|Check duplicates:|Hashing sequences:|Analysing:)\s*100% \(([0-9.]+)s\)", log_text, re.M):
        out.append((m.group(1), float(m.group(2))))
    return out


def reference_run(a, tmp, reps_a, reps_b, threads, want_cli=False):
    """oracle/_ref/compairr (the UNMODIFIED reference) on the first reps_a / reps_b repertoires of the
    bench's generator, `-t threads`; optionally our CLI on the same files.  Returns a dict."""
    from compairr_b200 import synth
    from oracle import oracle as orc
    indels = a.differences == 1 and not a.no_indels
    pool = synth.make_pool(POOL_SEED, a.pool_n)
    sa = synth.make_set(SEED_A, reps_a, a.per_rep, pool=pool, indel_mutants=True, workers=n_workers(a))
    sb = synth.make_set(SEED_B, reps_b, a.per_rep, pool=pool, indel_mutants=True, workers=n_workers(a))
    probes = int(sum_probes(sa, a.differences, indels))
    fa, fb = os.path.join(tmp, "a.tsv"), os.path.join(tmp, "b.tsv")
    orc.write_tsv(sa, fa, "a")
    orc.write_tsv(sb, fb, "b")
    n_a, n_b = sa.n, sb.n
    del sa, sb
    opt = ["-m", fa, fb, "-d", str(a.differences)] + (["-i"] if indels else [])
    log, out = os.path.join(tmp, "ref.log"), os.path.join(tmp, "ref.tsv")
    t0 = time.perf_counter()
    r = orc.run_reference(opt + ["-t", str(threads), "-l", log, "-o", out], timeout=3600)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-300:])
    ph = ref_phases(open(log).read())
    names = [p[0] for p in ph]
    if names != ["Computing hashes:", "Check duplicates:", "Computing hashes:", "Hashing sequences:", "Analysing:"]:
        raise RuntimeError(f"unexpected phase list in the reference log: {names}")
    res = {"probes": probes, "wall_s": wall, "t_hash_a": ph[0][1], "t_dups_a": ph[1][1], "t_hash_b": ph[2][1],
           "t_build_b": ph[3][1], "t_analyse": ph[4][1], "n_a": n_a, "n_b": n_b, "threads": threads,
           "cmd": "compairr " + " ".join(opt[:1] + ["a.tsv", "b.tsv"] + opt[3:]) + f" -t {threads}", "out": out}
    if want_cli:
        ours_out, ours_log = os.path.join(tmp, "ours.tsv"), os.path.join(tmp, "ours.log")
        walls = []
        for _ in range(2):   # the second run has the CUDA driver's files in the page cache too
            t0 = time.perf_counter()
            r = subprocess.run([CLI] + opt + ["-l", ours_log, "-o", ours_out], capture_output=True, text=True, timeout=3600)
            walls.append(time.perf_counter() - t0)
            if r.returncode != 0:
                raise RuntimeError("our CLI failed: " + r.stderr[-300:])
        same = open(out, "rb").read() == open(ours_out, "rb").read()
        res["cli"] = {"ours_wall_s": min(walls), "ours_wall_s_runs": walls, "reference_wall_s": wall,
                      "speedup": wall / min(walls), "outputs_identical": same,
                      "what": f"files in -> files out, {res['cmd']} vs compairr_b200 (1 GPU) on the same two TSV files "
                              f"({n_a} + {n_b} sequences), matrix files compared byte for byte"}
    return res


def reference_value(a, res, probes_target):
    """The reference's whole-job rate on the arm's workload: set-B phases as measured (set B is
    complete), set-A phases scaled by probes (analysis time is linear in probes)."""
    scale = probes_target / res["probes"]
    t = res["t_hash_b"] + res["t_build_b"] + scale * (res["t_hash_a"] + res["t_dups_a"] + res["t_analyse"])
    sample = (f"oracle/_ref/{res['cmd']}: A={res['n_a']} vs the FULL B={res['n_b']} sequences of the same generator, one run, "
              f"{res['probes']} probes; log phases hash B {res['t_hash_b']:.2f} s + build B {res['t_build_b']:.2f} s (as measured) + "
              f"[hash A {res['t_hash_a']:.3f} + dups A {res['t_dups_a']:.3f} + analysing {res['t_analyse']:.3f}] s x {scale:.1f} "
              f"(EXTRAPOLATED to the arm's set A by probes) = {t:.2f} s; TSV parsing excluded; wall of the run {res['wall_s']:.1f} s")
    return {"value": probes_target / t, "unit": "probes/s", "cores": res["threads"], "kind": "reference", "sample": sample,
            "hot_path_s": t, "analyse_rate_probes_s": res["probes"] / res["t_analyse"]}


def all_threads():
    return max(1, min(os.cpu_count() or 1, 256))   # -t is capped at 256 (src/compairr.h:109)


def shm_tmp():
    return tempfile.mkdtemp(prefix="compairr_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    if not orc.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/compairr is not built"}))
        return
    indels = a.differences == 1 and not a.no_indels
    tmp = shm_tmp()
    try:
        t0 = time.perf_counter()
        res = reference_run(a, tmp, a.sample_reps_a, a.reps_b, all_threads())
        # the arm's workload at N GPUs: set A = N x reps_a_per_gpu repertoires (weak scaling)
        target = res["probes"] * (a.gpus * a.reps_a_per_gpu / a.sample_reps_a)
        cb = reference_value(a, res, target)
        total = time.perf_counter() - t0
    except Exception as e:  # the reference binary is built in the build container and shipped
        print(json.dumps({"impl": "reference", "unavailable": f"reference run failed: {e}"[:300]}))
        return
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    line = {
        "impl": "reference", "metric": "variant probes/s (-m overlap hot path)", "value": cb["value"],
        "unit": "probes/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "steps_timed": 1,
        "ms_per_step": 1e3 * cb["hot_path_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": cb["sample"],
                   "note": "one timed run whatever --steps says (a run on the full set B takes about a minute); "
                           f"whole arm {total:.0f} s incl. generating and writing the files"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "probes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- our arm -----------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from compairr_b200 import Engine, NarrowSet, OverlapOptions, synth
    from compairr_b200 import dist as cdist
    from compairr_b200.seqset import SeqSet

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    indels = a.differences == 1 and not a.no_indels
    fields = ("residues", "offsets", "v_gene", "j_gene", "rep", "count")
    nfields = ("residues", "lengths", "v_gene", "j_gene", "rep", "count")
    keep = []

    def pin_arr(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        keep.append(t)
        return t.numpy()

    def pin(s):
        out = {f: pin_arr(getattr(s, f)) for f in fields}
        return SeqSet(out["residues"], out["offsets"], out["v_gene"], out["j_gene"], out["rep"], out["count"],
                      s.n_reps, index_base=s.index_base)

    def pin_narrow(s, n_reps=None):
        ns = NarrowSet.from_seqset(s)
        for f in nfields:
            setattr(ns, f, pin_arr(getattr(ns, f)))
        if n_reps is not None:
            ns.n_reps = n_reps
        return ns

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_over_ranks(vals_max, vals_sum):
        if world == 1:
            return list(vals_max), list(vals_sum)
        mx = torch.tensor(vals_max, dtype=torch.float64, device=dev)
        sm = torch.tensor(vals_sum, dtype=torch.float64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return mx.tolist(), sm.tolist()

    # ---- data: set B identical on every rank (rank 0 generates, the others read /dev/shm) ----------
    pool = synth.make_pool(POOL_SEED, a.pool_n)
    t_gen = time.perf_counter()
    shm = f"/dev/shm/compairr_bench_B_{a.reps_b}x{a.per_rep}_{os.getuid()}"
    if world > 1:
        if rank == 0:
            b = synth.make_set(SEED_B, a.reps_b, a.per_rep, pool=pool, indel_mutants=True, workers=n_workers(a))
            os.makedirs(shm, exist_ok=True)
            for f in fields:
                np.save(os.path.join(shm, f + ".npy"), getattr(b, f))
        dist.barrier()
        if rank != 0:
            arrs = {f: np.load(os.path.join(shm, f + ".npy"), mmap_mode="r") for f in fields}
            b = SeqSet(arrs["residues"], arrs["offsets"], arrs["v_gene"], arrs["j_gene"], arrs["rep"], arrs["count"], a.reps_b)
        dist.barrier()
        if rank == 0:
            shutil.rmtree(shm, ignore_errors=True)
    else:
        b = synth.make_set(SEED_B, a.reps_b, a.per_rep, pool=pool, indel_mutants=True, workers=n_workers(a))

    def make_a_shard(reps_per_rank, block_reps=8):
        """repertoires [rank*R, rank*R + R) of a world*R-repertoire set A"""
        s = synth.make_set(SEED_A, reps_per_rank, a.per_rep, pool=pool, indel_mutants=True, block_reps=block_reps,
                           workers=max(1, n_workers(a) // world), first_rep=rank * reps_per_rank)
        s.rep += np.uint32(rank * reps_per_rank)
        s.n_reps = world * reps_per_rank
        s.index_base = rank * reps_per_rank * a.per_rep
        return s

    ra = a.reps_a_per_gpu
    a_sh = make_a_shard(ra)
    t_gen = time.perf_counter() - t_gen
    probes_rank = int(sum_probes(a_sh, a.differences, indels))
    n_b, n_a = b.n, a_sh.n

    b_pin, a_pin = pin(b), pin(a_sh)
    # end-to-end inputs: narrow columns (lengths + smallest lossless dtypes, cb_set_cols) in pinned
    # host memory; of set B only this rank's shard (cb_shard_range)
    bf, bc = cdist.shard_range(b.n, rank, world)
    b_shard_e2e = pin_narrow(b.slice(bf, bc), n_reps=b.n_reps)
    b_shard_e2e.index_base = 0              # of the whole set
    a_e2e = pin_narrow(a_sh)
    h2d_rank = b_shard_e2e.nbytes() + a_e2e.nbytes()
    a_d2 = a_sh.slice(0, min(a.d2_seeds, a_sh.n))
    a_d2_pin = pin(SeqSet(a_d2.residues[int(a_d2.offsets[0]):int(a_d2.offsets[-1])], a_d2.offsets - a_d2.offsets[0], a_d2.v_gene,
                          a_d2.j_gene, a_d2.rep, a_d2.count, a_d2.n_reps, index_base=a_d2.index_base))
    probes_d2_rank = int(sum_probes(a_d2, 2, False))
    del a_sh, a_d2

    # ---- engine + communicator (NCCL inside the library; torch.distributed carries the 128-byte id) ----
    opts = OverlapOptions(differences=a.differences, indels=indels, device=local, bloom_bits_per_key=a.bloom_bits)
    eng = Engine(opts, n_reps_a=world * ra)
    if world > 1:
        eng.comm_init_rank(cdist.exchange_unique_id(Engine.comm_unique_id), rank, world)
    # a dedicated (non-default) torch stream shared with the engine: CUDA events recorded on it
    # bracket the engine's kernels (and its NCCL calls, which run on the same stream)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    db = eng.upload(b_pin)
    da = eng.upload(a_pin)
    eng.build_b(db)
    debug = bool(os.environ.get("BENCH_DEBUG"))

    def resident_step(e, dset_b, dset_a, n_seeds=None):
        """hash B, build, duplicates, hash A, enumerate/probe/accumulate, all-reduce -> (stats, launches)"""
        n = 0
        e.rehash(dset_b); n += 1
        e.build_b(dset_b); n += e.stats()["kernel_launches"]
        sb = e.stats()
        e.rehash(dset_a); n += 1
        e.clear_matrix(); n += 1
        e.run(dset_a, 0, n_seeds)
        st = e.stats()
        n += st["kernel_launches"] + 1          # + the probe-count bookkeeping kernel
        if world > 1:
            e.allreduce_matrix(); n += 1
        if debug and rank == 0:
            print(f"[step] build {sb['ms_build_b']:.1f} + dups {sb['ms_dups_b']:.1f}, probe {st['ms_probe']:.1f}", file=sys.stderr, flush=True)
        return st, n

    def timed_resident(e, dset_b, dset_a, steps, warmup, n_seeds=None, sample_clocks=False):
        for _ in range(warmup):
            resident_step(e, dset_b, dset_a, n_seeds)
        sync_all()
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern, launches = [], 0
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            st, launches = resident_step(e, dset_b, dset_a, n_seeds)
            kern.append(st["ms_probe"])
        ev1.record(stream)
        sync_all()
        t1 = time.perf_counter()
        return {"ms_dev": ev0.elapsed_time(ev1), "kernel_ms": statistics.mean(kern), "launches": launches, "stats": st,
                "clocks": sampler.stop(t0, t1) if sampler else None, "wall_ms": 1e3 * (t1 - t0)}

    # ---- headline: resident ------------------------------------------------------------------------------
    main = timed_resident(eng, db, da, a.steps, a.warmup, sample_clocks=True)
    checksum = float(eng.matrix().sum())
    run_stats = main["stats"]
    build_stats_launch = eng.stats()

    # ---- end to end: host buffers in, matrix out, every step ------------------------------------------
    def step_e2e():
        t0 = time.perf_counter()
        eng.set_b_sharded(b_shard_e2e, n_b)     # 1/world over PCIe, all-gather over NVLink, build
        sb = eng.stats()
        t1 = time.perf_counter()
        eng.clear_matrix()
        eng.run_a(a_e2e)
        t2 = time.perf_counter()
        eng.allreduce_matrix()
        m = eng.matrix()
        t3 = time.perf_counter()
        if debug:
            print(f"[e2e rank {rank}] set_b {1e3*(t1-t0):.1f} (upload+hash {sb['ms_hash_b']:.1f}, gather {sb['ms_gather_b']:.1f}, "
                  f"build {sb['ms_build_b']:.1f}, dups {sb['ms_dups_b']:.1f}) run_a {1e3*(t2-t1):.1f} reduce+read {1e3*(t3-t2):.1f}",
                  file=sys.stderr, flush=True)
        return m
    for _ in range(a.warmup):
        step_e2e()
    sync_all()
    t0e = time.perf_counter()
    for _ in range(a.steps):
        m_e2e = step_e2e()
    e2e_stats = eng.stats()
    sync_all()
    ms_e2e = 1e3 * (time.perf_counter() - t0e)
    d2h = int(m_e2e.nbytes)
    e2e_checksum = float(m_e2e.sum())

    # ---- strong scaling: the fixed C3 problem, set A = strong_reps_a repertoires over all GPUs ------------
    strong = None
    reps_strong = a.strong_reps_a or a.reps_b
    if not a.skip_strong and reps_strong % world == 0 and (reps_strong // world) % 5 == 0:
        da.free()
        t0 = time.perf_counter()
        a_st = make_a_shard(reps_strong // world, block_reps=5)   # blocks of 5: 1000 / 8 ranks = 125 repertoires each
        probes_st = int(sum_probes(a_st, a.differences, indels))
        t_gen_st = time.perf_counter() - t0
        eng_s = eng
        if reps_strong != world * ra:      # the matrix has other dimensions: its own context, same communicator rules
            eng_s = Engine(opts, n_reps_a=reps_strong)
            if world > 1:
                eng_s.comm_init_rank(cdist.exchange_unique_id(Engine.comm_unique_id), rank, world)
            eng_s.set_stream(stream.cuda_stream)
        db_s = db if eng_s is eng else eng_s.upload(b_pin)
        da_s = eng_s.upload(a_st)
        n_a_st = a_st.n
        del a_st
        steps_s, warm_s = max(1, min(a.steps, 5)), max(3, min(a.warmup, 3))
        r = timed_resident(eng_s, db_s, da_s, steps_s, warm_s)
        (ms_dev_s, kern_s), (probes_all_s,) = reduce_over_ranks([r["ms_dev"], r["kernel_ms"]], [float(probes_st)])
        strong = {"value": probes_all_s / (ms_dev_s / steps_s * 1e-3), "unit": "probes/s", "scaling": "strong",
                  "ms_per_step": ms_dev_s / steps_s, "kernel_ms_max_rank": kern_s, "steps": steps_s, "warmup": warm_s,
                  "workload": f"the whole C3 problem: B={a.reps_b}x{a.per_rep}, A={reps_strong}x{a.per_rep} in total "
                              f"({n_a_st} seeds per GPU), -m -d {a.differences}{' -i' if indels else ''}; step = hash B + build + "
                              "dups + hash A shard + kernels + all-reduce; set B's build is replicated (the serial part)",
                  "probes_per_step": int(probes_all_s), "matches_rank0": r["stats"]["matches"], "generate_s": round(t_gen_st, 1)}
        da_s.free()
        if eng_s is not eng:
            db_s.free()
            eng_s.close()
        da = None
    if da is not None:
        da.free()
    db.free()

    # ---- parity at this N: engine (sharded B, sharded A, all-reduce) vs the unmodified reference ------------
    parity = None
    if not a.skip_parity:
        parity = parity_sample(a, eng, rank, world, pool, indels, cdist, dist if world > 1 else None)
    eng.close()

    # ---- d = 2 (C4 geometry: -g), same set B ---------------------------------------------------------------
    d2 = None
    if not a.skip_d2:
        o2 = OverlapOptions(differences=2, ignore_genes=True, device=local, bloom_bits_per_key=a.bloom_bits)
        e2 = Engine(o2, n_reps_a=world * ra)
        if world > 1:
            e2.comm_init_rank(cdist.exchange_unique_id(Engine.comm_unique_id), rank, world)
        e2.set_stream(stream.cuda_stream)
        db2 = e2.upload(b_pin)
        da2 = e2.upload(a_d2_pin)
        e2.build_b(db2)
        r = timed_resident(e2, db2, da2, a.steps, a.warmup)
        (ms_dev2, kern2), (probes_all2,) = reduce_over_ranks([r["ms_dev"], r["kernel_ms"]], [float(probes_d2_rank)])
        d2 = {"value": probes_all2 / (ms_dev2 / a.steps * 1e-3), "unit": "probes/s", "ms_per_step": ms_dev2 / a.steps,
              "kernel_value": probes_d2_rank / (r["kernel_ms"] * 1e-3), "kernel_ms": kern2, "steps": a.steps, "warmup": a.warmup,
              "workload": f"C4 geometry: B={a.reps_b}x{a.per_rep}, A={a.d2_seeds} seeds per GPU, -m -d 2 -g -s product; "
                          "value = whole step (hash B + build + dups + hash A + kernels + all-reduce), kernel_value = enumeration + table kernels",
              "probes_per_step": int(probes_all2), "matches_rank0": r["stats"]["matches"],
              "bloom_pass_frac": r["stats"]["bloom_pass"] / max(r["stats"]["probes"], 1),
              "roofline_frac": 8.0 * probes_d2_rank / (r["kernel_ms"] * 1e-3) / 1e9 / peak_hbm()[0]}
        da2.free()
        db2.free()
        e2.close()

    # ---- C5 geometry, bounded: nucleotide queries in one repertoire vs a nucleotide set ----------------------
    c5 = None
    if not a.skip_c5:
        c5 = c5_section(a, rank, world, local, pool, stream, reduce_over_ranks, sync_all)

    # ---- reduce timings over ranks ------------------------------------------------------------------------
    (ms_dev, ms_e2e, kern), (probes_total, h2d_total) = reduce_over_ranks(
        [main["ms_dev"], ms_e2e, main["kernel_ms"]], [float(probes_rank), float(h2d_rank)])
    probes_total, h2d_total = int(probes_total), int(h2d_total)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_dev / a.steps
    value = probes_total / (ms_step * 1e-3)
    e2e_value = probes_total / (ms_e2e / a.steps * 1e-3)
    peak, peak_src = peak_hbm()
    achieved = 8.0 * probes_rank / (kern * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    bpk_now = build_stats_launch["bloom_bytes"] * 8.0 / max(n_b, 1)
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("probes_per_launch") and abs(tj.get("bits_per_key", 0) - bpk_now) < 1.0:
            # one captured launch covers probes_per_launch probes; a step's launches cover probes_rank
            traffic = tj["dram_bytes_per_launch"] * probes_rank / tj["probes_per_launch"]
        else:
            traffic_note = f"profiles/traffic.json was captured at {tj.get('bits_per_key')} bits/key, this run uses {bpk_now:.1f}: not used"
    line = {
        "metric": "variant probes/s (-m overlap hot path)", "value": value, "unit": "probes/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "set_b_sequences": n_b, "set_a_sequences_per_gpu": n_a,
                   "probes_per_step": probes_total, "l2": "inputs larger than L2 (class filters 4 x %.0f MiB, table %.0f MiB); no flush"
                   % (run_stats["bloom_bytes"] / 2**20 if run_stats["bloom_bytes"] else build_stats_launch["bloom_bytes"] / 2**20,
                      build_stats_launch["table_slots"] * 16 / 2**20),
                   "parallelism": f"set A sharded over {world} GPU(s); set B on every GPU (e2e: 1/{world} uploaded per rank, NVLink all-gather); "
                                  "NCCL all-reduce of the matrix inside the library",
                   "step": "hash B + build table/filters + dups + hash A + enumeration and table kernels (+ all-reduce)",
                   "generate_s": round(t_gen, 1), "matrix_checksum": checksum, "e2e_matrix_checksum": e2e_checksum},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "enum1_kernel<20,indels,32> + table_kernel", "kernel_ms": kern,
                     "bytes_per_probe": 8, "peak_source": peak_src,
                     "note": "8 B/probe is the algorithmic figure (one filter word per variant, SURVEY 8d). The class "
                             "filters let all candidates of a slot share one word, so the kernels move fewer bytes "
                             "than that (traffic = measured DRAM bytes of the enumeration launches of a step) and are bound "
                             "by the ALU pipe; the fraction says how fast the algorithmic work is done relative to an HBM stream",
                     "dram_bytes_per_probe": (traffic / probes_rank) if traffic else None, "traffic_note": traffic_note},
        "e2e": {"value": e2e_value, "unit": "probes/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": ms_e2e / a.steps, "ms_gather_b_rank0": e2e_stats.get("ms_gather_b", 0.0)},
        "gpu_launches": main["launches"] * a.steps,
        "clocks": main["clocks"],
        "wall_ms_per_step": main["wall_ms"] / a.steps,
        "matches_per_step_rank0": run_stats["matches"], "bloom_pass_frac": run_stats["bloom_pass"] / max(run_stats["probes"], 1),
        "d2": d2, "strong": strong, "c5": c5, "parity_checked": parity,
    }
    if world == 1 and not a.skip_cpu_baseline:
        tmp = shm_tmp()
        try:
            res = reference_run(a, tmp, a.sample_reps_a, a.reps_b, all_threads(), want_cli=os.path.exists(CLI))
            cb = reference_value(a, res, probes_total)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if "cli" in res:
                line["cli_wall"] = res["cli"]
                if line["parity_checked"] is not None:
                    line["parity_checked"]["full_size_b"] = {
                        "what": f"CLI (1 GPU) vs the reference binary on A={res['n_a']} x the full B={res['n_b']}: matrix files",
                        "matrix_identical": res["cli"]["outputs_identical"]}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "probes/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e}"[:200]}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def c5_section(a, rank, world, local, pool, stream, reduce_over_ranks, sync_all):
    """BASELINE config 5 at a size that fits the time box: Q nucleotide queries per GPU (one
    repertoire, rank r takes its own block of the generator) against c5_reps_b x per_rep nucleotide
    sequences.  (1) -x -n -d 2 -p --no-matrix: the hash path on ~43-nt sequences (ZP = 96 kernels),
    pairs drained to the host every run; (2) -x -n -g -d 3 -p --no-matrix: the brute-force path, dense
    length buckets on the tcgen05 one-hot int8 GEMM."""
    import torch
    from compairr_b200 import Engine, OverlapOptions, synth
    t0 = time.perf_counter()
    reps_q = (-(-a.c5_queries // a.per_rep) + 7) // 8 * 8      # whole generator blocks of 8 repertoires
    b = synth.make_set(SEED_B, a.c5_reps_b, a.per_rep, pool=pool, nucleotides=True, workers=n_workers(a))
    q = synth.make_set(SEED_A, reps_q, a.per_rep, pool=pool, nucleotides=True, single_repertoire=True,
                       workers=max(1, n_workers(a) // world), first_rep=rank * reps_q)
    q = q.slice(0, min(a.c5_queries, q.n))
    t_gen = time.perf_counter() - t0
    out = {"workload": f"C5 geometry, bounded: {q.n} nucleotide queries per GPU (one repertoire) vs B={a.c5_reps_b}x{a.per_rep} "
                       f"nucleotide sequences (mean length {b.residues.size / max(b.n, 1):.0f}); pairs collected and drained to the host",
           "generate_s": round(t_gen, 1)}
    Lq, Lb = np.diff(q.offsets).astype(np.int64), np.diff(b.offsets).astype(np.int64)
    pair_tests = float((np.bincount(Lq, minlength=512).astype(np.float64) * np.bincount(Lb, minlength=512)).sum())
    for name, kw, steps in (("d2_hash", dict(differences=2), max(1, min(a.steps, 3))),
                            ("d3_gemm", dict(differences=3, ignore_genes=True), max(1, min(a.steps, 3)))):
        eng = Engine(OverlapOptions(existence=True, nucleotides=True, no_matrix=True, want_pairs=True, device=local, **kw), n_reps_a=1)
        eng.set_stream(stream.cuda_stream)
        db, dq = eng.upload(b), eng.upload(q)
        eng.build_b(db)
        ms, kern, pairs = [], [], 0
        for it in range(1 + steps):                       # one warm-up run (sizes the pair buffer and the queue)
            sync_all()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            eng.run(dq)
            st = eng.stats()
            p = eng.drain_pairs()
            ev1.record(stream)
            sync_all()
            if it:
                ms.append(ev0.elapsed_time(ev1))
                kern.append(st["ms_probe"])
                pairs = len(p)
        work = float(sum_probes(q, 2, False)) if name == "d2_hash" else pair_tests
        (ms_max, kern_max), (work_all,) = reduce_over_ranks([statistics.mean(ms), statistics.mean(kern)], [work])
        out[name] = {"value": work_all / (ms_max * 1e-3), "unit": "probes/s" if name == "d2_hash" else "pair tests/s",
                     "kernel_value_rank0": work / (statistics.mean(kern) * 1e-3), "ms_per_step": ms_max, "kernel_ms": kern_max,
                     "steps": steps, "warmup": 1, "pairs_rank0": pairs, "matches_rank0": st["matches"],
                     "args": "-x -n -d 2 -p --no-matrix" if name == "d2_hash" else "-x -n -g -d 3 -p --no-matrix",
                     "note": "value = run + drain of the pairs to host memory (device span); kernel_value = the kernels alone"}
        dq.free()
        db.free()
        eng.close()
    return out


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def parity_sample(a, eng, rank, world, pool, indels, cdist, dist):
    """Every rank runs its part of A_s x B_s (first repertoires of the bench's generator) through the
    N-GPU path — cb_set_b_sharded, A shard by expected probes, cb_allreduce_matrix — and rank 0
    compares the matrix, cell for cell as the CLI would print it, with the output of the UNMODIFIED
    reference binary on the same two sets (written as TSV files)."""
    from compairr_b200 import Engine, OverlapOptions, synth
    from oracle import oracle as orc
    ra, rb = (int(x) for x in a.parity_reps.split("x"))
    sa = synth.make_set(SEED_A, ra, a.per_rep, pool=pool, indel_mutants=True, workers=2)
    sb = synth.make_set(SEED_B, rb, a.per_rep, pool=pool, indel_mutants=True, workers=2)
    e = Engine(OverlapOptions(differences=a.differences, indels=indels, device=eng.opts.device), n_reps_a=ra)
    if world > 1:
        e.comm_init_rank(cdist.exchange_unique_id(Engine.comm_unique_id), rank, world)
    m = cdist.overlap_rank(e, sa, sb, rank, world, a.differences, indels)
    dups = e.dups_b()
    e.close()
    if rank != 0:
        return None
    out = {"n_gpus": world, "sample": f"A={ra}x{a.per_rep} vs B={rb}x{a.per_rep} of the bench generator, -m -d {a.differences}{' -i' if indels else ''}",
           "path": "cb_set_b_sharded (NCCL all-gather) + sharded A + cb_allreduce_matrix" if world > 1 else "cb_set_b_cols + cb_run_a",
           "against": "oracle/_ref/compairr (unmodified reference binary)"}
    if not orc.have_reference():
        out.update({"matrix_identical": None, "note": "oracle/_ref/compairr is not built"})
        return out
    tmp = shm_tmp()
    try:
        fa, fb, fo = os.path.join(tmp, "a.tsv"), os.path.join(tmp, "b.tsv"), os.path.join(tmp, "o.tsv")
        orc.write_tsv(sa, fa, "a")
        orc.write_tsv(sb, fb, "b")
        r = orc.run_reference(["-m", fa, fb, "-d", str(a.differences)] + (["-i"] if indels else []) +
                              ["-t", str(all_threads()), "-l", os.path.join(tmp, "log"), "-o", fo], timeout=1800)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-200:])
        rows = [ln.rstrip("\n").split("\t") for ln in open(fo)]
        want = np.array([[float(x) for x in row[1:]] for row in rows[1:]])
        # the generator's repertoire names sort like their numbers (R0000, R0001, ...)
        # value AND text: every cell formatted as the CLI prints it (%.10lg) equals the reference's field
        same = want.shape == m.shape and bool(np.array_equal(want, m)) and \
            all(format(x, ".10g") == y for x, y in zip(m.ravel(), [c for row in rows[1:] for c in row[1:]]))
        log = open(os.path.join(tmp, "log")).read()
        md = re.search(r"Warning: (\d+) duplicates detected in repertoire set 2", log)
        out.update({"matrix_identical": bool(same), "cells": int(m.size), "matrix_sum": float(m.sum()),
                    "dups_b_identical": (int(md.group(1)) if md else 0) == dups})
    except Exception as ex:
        out.update({"matrix_identical": None, "note": f"reference run failed: {ex}"[:200]})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
