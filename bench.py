#!/usr/bin/env python
"""bench.py — measures BASELINE.json's metric (variant probes/s of the `-m` overlap hot path)
on synthetic repertoires of the C3 shape, one process per GPU.

  python bench.py [--gpus N --steps K --warmup W]                      our arm (N=1 by default)
  torchrun --nproc-per-node N ... bench.py --gpus N ...                 our arm, N ranks (NCCL)
  python bench.py --impl reference ...                                  the reference's CPU build

Workload ("C3"): set B = 1000 repertoires x 100 000 AA CDR3 (10^8 sequences, replicated on every
GPU), set A = 100 repertoires x 100 000 per GPU (10^7 seeds per rank, WEAK scaling: rank r holds
repertoires [100 r, 100 r + 100) of set A), `-m -d 1 -i` (substitutions + indels), score product.
A step is one whole hot-path pass: hash B, build table + Bloom over B, count duplicates, hash the
A shard, enumerate + probe + verify + accumulate, and (N > 1) the NCCL allreduce of the matrix.

  value     probes of all ranks / max-over-ranks device time, inputs resident in HBM
  e2e       same through the C ABI with pinned HOST buffers: H2D of both sets and the D2H read of
            the matrix inside the timed region
  roofline  the enumeration + table kernels: 8 algorithmic bytes per probe (one filter word) / their
            CUDA-event duration, against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the unmodified reference binary (oracle/_ref/compairr, all host threads) on a
            bounded sample of the same workload, hot-path phases from its log
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
    ap.add_argument("--sample-reps-a", type=int, default=10, help="reference sample: set-A repertoires")
    ap.add_argument("--sample-reps-b", type=int, default=50, help="reference sample: set-B repertoires")
    ap.add_argument("--workers", type=int, default=0, help="generator processes (0 = auto)")
    ap.add_argument("--pool-n", type=int, default=4_000_000, help="size of the shared public pool")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic C3: B={a.reps_b}x{a.per_rep} AA, A={a.reps_a_per_gpu}x{a.per_rep} per GPU, public pool {a.pool_n}, "
            f"-m -d {a.differences}{'' if a.no_indels or a.differences != 1 else ' -i'} -s product")


def n_workers(a):
    return a.workers or max(1, min(32, (os.cpu_count() or 8) - 2))


# ---- the reference's CPU build on a bounded sample -------------------------------------------------

def reference_sample(a, steps, warmup):
    """Runs oracle/_ref/compairr (the UNMODIFIED reference, all host threads) on the first
    sample_reps_a / sample_reps_b repertoires of the same synthetic sets.  Returns the dict for
    `cpu_baseline` / the reference arm.  Hot-path time = the log's `Computing hashes` (both sets) +
    `Check duplicates` + `Hashing sequences` + `Analysing` phases (src/util.cc:61-68)."""
    from compairr_b200 import synth
    from oracle import oracle as orc
    if not orc.have_reference():
        return None
    threads = max(1, min(os.cpu_count() or 1, 256))   # -t is capped at 256 (src/compairr.h:109)
    pool = synth.make_pool(POOL_SEED, a.pool_n)
    indels = a.differences == 1 and not a.no_indels
    sa = synth.make_set(SEED_A, a.sample_reps_a, a.per_rep, pool=pool, indel_mutants=True, workers=n_workers(a))
    sb = synth.make_set(SEED_B, a.sample_reps_b, a.per_rep, pool=pool, indel_mutants=True, workers=n_workers(a))
    probes = int(sum_probes(sa, a.differences, indels))
    tmp = tempfile.mkdtemp(prefix="compairr_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fa, fb = os.path.join(tmp, "a.tsv"), os.path.join(tmp, "b.tsv")
        sa.write_tsv(fa, "a")
        sb.write_tsv(fb, "b")
        args = ["-m", fa, fb, "-d", str(a.differences)] + (["-i"] if indels else []) + \
               ["-t", str(threads), "-l", os.path.join(tmp, "log.txt"), "-o", os.path.join(tmp, "out.tsv")]
        hot, wall = [], []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            r = orc.run_reference(args, timeout=3600)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError(r.stderr)
            log = open(os.path.join(tmp, "log.txt")).read()
            phases = {k: sum(float(x) for x in re.findall(re.escape(k) + r"\s*100% \(([0-9.]+)s\)", log))
                      for k in ("Computing hashes:", "Check duplicates:", "Hashing sequences:", "Analysing:")}
            if it >= warmup:
                hot.append(sum(phases.values()))
                wall.append(dt)
        t_hot = statistics.mean(hot)
        return {"value": probes / t_hot, "unit": "probes/s", "cores": threads, "kind": "reference",
                "sample": (f"oracle/_ref/compairr -m -d {a.differences}{' -i' if indels else ''} -t {threads} on "
                           f"A={a.sample_reps_a}x{a.per_rep} vs B={a.sample_reps_b}x{a.per_rep} of the same generator; "
                           f"{probes} probes; hot-path phases {t_hot:.3f} s of {statistics.mean(wall):.3f} s wall "
                           f"(TSV parsing excluded), mean of {steps} run(s)"),
                "ms_per_step": 1e3 * t_hot, "probes": probes, "wall_s": statistics.mean(wall)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


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


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        res = reference_sample(a, a.steps, a.warmup)
    except Exception as e:  # the reference binary is built in the build container and shipped
        print(json.dumps({"impl": "reference", "unavailable": f"reference run failed: {e}"[:300]}))
        return
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/compairr is not built"}))
        return
    line = {
        "impl": "reference", "metric": "variant probes/s (-m overlap hot path)", "value": res["value"],
        "unit": "probes/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "probes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def pinned_copy(torch, arr):
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t, t.numpy()


def run_ours(a):
    import torch
    import torch.distributed as dist
    from compairr_b200 import Engine, OverlapOptions, synth
    from compairr_b200.seqset import SeqSet

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    indels = a.differences == 1 and not a.no_indels

    # ---- data: set B identical on every rank (rank 0 generates, the others read /dev/shm) ----------
    pool = synth.make_pool(POOL_SEED, a.pool_n)
    t_gen = time.perf_counter()
    shm = f"/dev/shm/compairr_bench_B_{a.reps_b}x{a.per_rep}_{os.getuid()}"
    fields = ("residues", "offsets", "v_gene", "j_gene", "rep", "count")
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
    # set A shard of this rank: repertoires [rank*R, rank*R + R) of a world*R-repertoire set
    ra = a.reps_a_per_gpu
    a_sh = synth.make_set(SEED_A, ra, a.per_rep, pool=pool, indel_mutants=True, workers=max(1, n_workers(a) // world),
                          first_rep=rank * ra)
    a_sh.rep += np.uint32(rank * ra)
    a_sh.n_reps = world * ra
    a_sh.index_base = rank * ra * a.per_rep
    t_gen = time.perf_counter() - t_gen
    probes_rank = int(sum_probes(a_sh, a.differences, indels))

    # pinned host copies for the end-to-end leg
    keep = []
    def pin(s):
        out = {}
        for f in fields:
            t, v = pinned_copy(torch, getattr(s, f))
            keep.append(t)
            out[f] = v
        return SeqSet(out["residues"], out["offsets"], out["v_gene"], out["j_gene"], out["rep"], out["count"],
                      s.n_reps, index_base=s.index_base)
    b_pin, a_pin = pin(b), pin(a_sh)
    # end-to-end inputs: the narrow-column form of the same sets (lengths + smallest lossless
    # dtypes, cb_set_cols), in pinned host memory
    from compairr_b200 import NarrowSet
    nfields = ("residues", "lengths", "v_gene", "j_gene", "rep", "count")
    def pin_narrow(s):
        ns = NarrowSet.from_seqset(s)
        for f in nfields:
            t, v = pinned_copy(torch, getattr(ns, f))
            keep.append(t)
            setattr(ns, f, v)
        return ns
    b_e2e, a_e2e = pin_narrow(b), pin_narrow(a_sh)
    h2d = b_e2e.nbytes() + a_e2e.nbytes()   # per rank: every rank uploads all of B and its A shard
    n_b, n_a = b.n, a_sh.n
    del b, a_sh, pool                        # only the pinned copies are used from here on

    opts = OverlapOptions(differences=a.differences, indels=indels, device=local, bloom_bits_per_key=a.bloom_bits)
    eng = Engine(opts, n_reps_a=world * ra)
    # a dedicated (non-default) torch stream shared with the engine: CUDA events recorded on it
    # bracket the engine's kernels, and torch ops / NCCL are ordered with them
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    matrix = torch.zeros((world * ra, a.reps_b), dtype=torch.float64, device=f"cuda:{local}")

    db = eng.upload(b_pin)
    da = eng.upload(a_pin)
    eng.build_b(db)
    eng.bind_matrix(matrix.data_ptr(), matrix.shape[0], matrix.shape[1])

    kernel_ms, launches = [], 0

    debug = bool(os.environ.get("BENCH_DEBUG"))

    def step_resident():
        nonlocal launches
        n = 0
        t0 = time.perf_counter()
        eng.rehash(db); n += 1
        t1 = time.perf_counter()
        eng.build_b(db); n += eng.stats()["kernel_launches"]
        sb = eng.stats()
        t2 = time.perf_counter()
        eng.rehash(da); n += 1
        matrix.zero_(); n += 1
        eng.run(da)
        st = eng.stats()
        t3 = time.perf_counter()
        if debug and rank == 0:
            print(f"[step] rehashB {1e3*(t1-t0):.1f} build {1e3*(t2-t1):.1f} (dev {sb['ms_build_b']:.1f} + dups {sb['ms_dups_b']:.1f}) "
                  f"run {1e3*(t3-t2):.1f} (probe {st['ms_probe']:.1f})", file=sys.stderr, flush=True)
        n += st["kernel_launches"] + 1          # + the probe-count bookkeeping kernel
        if world > 1:
            dist.all_reduce(matrix)
        launches = n
        return st

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_resident()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(a.steps):
        st = step_resident()
        kernel_ms.append(st["ms_probe"])
    ev1.record(stream)
    sync_all()
    t1 = time.perf_counter()
    ms_dev = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t0, t1) if sampler else None
    checksum = float(matrix.sum().item())
    run_stats = eng.stats()

    # ---- end to end: host buffers in, matrix out, every step ------------------------------------------
    eng.bind_matrix(0, 0, 0)
    da.free()
    db.free()
    def step_e2e():
        eng.set_b(b_e2e)
        eng.clear_matrix()
        eng.run_a(a_e2e)
        m = eng.matrix()
        if world > 1:
            mt = torch.from_numpy(m).cuda()
            dist.all_reduce(mt)
            m = mt.cpu().numpy()
        return m
    for _ in range(a.warmup):
        step_e2e()
    sync_all()
    t0e = time.perf_counter()
    for _ in range(a.steps):
        m_e2e = step_e2e()
    sync_all()
    ms_e2e = 1e3 * (time.perf_counter() - t0e)
    d2h = int(m_e2e.nbytes)
    e2e_checksum = float(m_e2e.sum())
    eng.close()

    # ---- reduce timings over ranks ------------------------------------------------------------------------
    vals = torch.tensor([ms_dev, ms_e2e, float(probes_rank), statistics.mean(kernel_ms)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_dev, ms_e2e, kern = mx[0].item(), mx[1].item(), mx[3].item()
        probes_total = int(sm[2].item())
    else:
        kern = statistics.mean(kernel_ms)
        probes_total = probes_rank
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_dev / a.steps
    value = probes_total / (ms_step * 1e-3)
    e2e_value = probes_total / (ms_e2e / a.steps * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = 8.0 * probes_rank / (kern * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("probes_per_launch"):
            # one captured launch covers probes_per_launch probes; a step's launches cover probes_rank
            traffic = tj["dram_bytes_per_launch"] * probes_rank / tj["probes_per_launch"]
    line = {
        "metric": "variant probes/s (-m overlap hot path)", "value": value, "unit": "probes/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "set_b_sequences": n_b, "set_a_sequences_per_gpu": n_a,
                   "probes_per_step": probes_total, "l2": "inputs larger than L2 (Bloom %.0f MiB, table %.0f MiB); no flush"
                   % (run_stats["bloom_bytes"] / 2**20, run_stats["table_slots"] * 16 / 2**20),
                   "parallelism": f"set A sharded over {world} GPU(s), set B replicated, NCCL allreduce of the matrix",
                   "step": "hash B + build table/Bloom + dups + hash A + probe kernel (+ allreduce)",
                   "generate_s": round(t_gen, 1), "matrix_checksum": checksum, "e2e_matrix_checksum": e2e_checksum},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "variant_kernel<20,indels,1>", "kernel_ms": kern,
                     "bytes_per_probe": 8, "peak_source": peak_src,
                     "note": "8 B/probe is the algorithmic figure (one filter word per variant, SURVEY 8d). The parity "
                             "filters let all candidates of a slot share one word, so the kernel moves far fewer "
                             "bytes than that (traffic = measured DRAM bytes per launch) and is instruction-bound; "
                             "the fraction says how fast the algorithmic work is done relative to an HBM stream",
                     "dram_bytes_per_probe": (traffic / probes_rank) if traffic else None},
        "e2e": {"value": e2e_value, "unit": "probes/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches * a.steps,
        "clocks": clocks,
        "wall_ms_per_step": 1e3 * (t1 - t0) / a.steps,
        "matches_per_step_rank0": run_stats["matches"], "bloom_pass_frac": run_stats["bloom_pass"] / max(run_stats["probes"], 1),
    }
    if world == 1 and not a.skip_cpu_baseline:
        try:
            ref = reference_sample(a, 1, 0)
            if ref:
                line["cpu_baseline"] = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "probes/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e}"[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
