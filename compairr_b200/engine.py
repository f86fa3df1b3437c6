"""Host-side mirror of the reference's overlap() seam (src/overlap.cc:607-1079) on numpy arrays.

`Engine` is a thin object wrapper over the C ABI: options -> cb_create, set B -> cb_set_b /
cb_upload + cb_build_b, set A -> cb_run_a / cb_run, results -> cb_get_matrix / cb_drain_pairs.
`overlap()` is the one-call form used by the parity tests: same options as `compairr -m/-x`,
same matrix (rows/cols in the caller's repertoire numbering) and the same pair list."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import cabi
from .seqset import SeqSet

SCORES = {"product": 0, "ratio": 1, "min": 2, "max": 3, "mean": 4, "mh": 5, "jaccard": 6}


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.message = msg


@dataclass
class OverlapOptions:
    """The reference's opt_* globals the hot path reads (src/compairr.h:139-162)."""
    differences: int = 0
    indels: bool = False
    ignore_genes: bool = False
    ignore_counts: bool = False
    score: str = "product"
    existence: bool = False
    nucleotides: bool = False
    no_matrix: bool = False
    want_pairs: bool = False
    device: int = 0
    seed: int = 1
    bloom_bits_per_key: float = 0.0
    table_load_pct: int = 0
    pairs_capacity: int = 0
    flags: int = int(__import__("os").environ.get("CB_FLAGS", "0"))
    queue_capacity: int = 0     # candidate-queue entries (0 = sized from the run); overflow only costs time


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _cb_set(s: SeqSet) -> cabi.cb_set:
    cs = cabi.cb_set()
    cs.n = s.n
    cs.residues = _ptr(s.residues)
    cs.offsets = _ptr(s.offsets)
    cs.v_gene = _ptr(s.v_gene)
    cs.j_gene = _ptr(s.j_gene)
    cs.rep = _ptr(s.rep)
    cs.count = _ptr(s.count)
    cs.n_reps = s.n_reps
    cs.longest = 0
    cs.index_base = s.index_base
    return cs


def _col(a: Optional[np.ndarray]) -> cabi.cb_col:
    c = cabi.cb_col()
    if a is not None:
        c.data = a.ctypes.data
        c.width = a.dtype.itemsize
    return c


def _cb_set_cols(s) -> cabi.cb_set_cols:
    """s: a NarrowSet (seqset.py) — lengths instead of offsets, smallest lossless dtypes."""
    cs = cabi.cb_set_cols()
    cs.n = s.n
    cs.residues = s.residues.ctypes.data
    cs.lengths = _col(s.lengths)
    cs.v_gene, cs.j_gene, cs.rep, cs.count = _col(s.v_gene), _col(s.j_gene), _col(s.rep), _col(s.count)
    cs.n_reps = s.n_reps
    cs.longest = s.longest
    cs.index_base = s.index_base
    return cs


class DeviceSet:
    """A sequence set resident on the GPU (cb_dset)."""

    def __init__(self, eng: "Engine", handle, n: int, n_reps: int):
        self.eng, self.handle, self.n, self.n_reps = eng, handle, n, n_reps

    def free(self):
        if self.handle:
            cabi.lib.cb_free_set(self.eng._ctx, self.handle)
            self.handle = None


class Engine:
    def __init__(self, opts: OverlapOptions, n_reps_a: int = 1):
        cfg = cabi.cb_config()
        cfg.abi_version = cabi.ABI_VERSION
        cfg.device = opts.device
        cfg.alphabet_size = 4 if opts.nucleotides else 20
        cfg.differences = opts.differences
        cfg.indels = int(opts.indels)
        cfg.ignore_genes = int(opts.ignore_genes)
        cfg.ignore_counts = int(opts.ignore_counts)
        cfg.score = SCORES[opts.score.lower()]
        cfg.mode = 1 if opts.existence else 0
        cfg.no_matrix = int(opts.no_matrix)
        cfg.want_pairs = int(opts.want_pairs)
        cfg.n_reps_a = n_reps_a
        cfg.seed = opts.seed
        cfg.bloom_bits_per_key_x16 = int(round(opts.bloom_bits_per_key * 16))
        cfg.table_load_pct = opts.table_load_pct
        cfg.pairs_capacity = opts.pairs_capacity
        cfg.flags = opts.flags
        cfg.queue_capacity = opts.queue_capacity
        self.opts = opts
        self._ctx = C.c_void_p()
        rc = cabi.lib.cb_create(C.byref(cfg), C.byref(self._ctx))
        if rc:
            self._ctx = None
            raise EngineError(rc, cabi.lib.cb_global_error().decode())

    # -- plumbing
    def _check(self, rc):
        if rc:
            raise EngineError(rc, cabi.lib.cb_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            cabi.lib.cb_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        self._check(cabi.lib.cb_set_stream(self._ctx, C.c_void_p(cuda_stream_ptr)))

    # -- sets
    def upload(self, s) -> DeviceSet:
        """s: SeqSet (reference-width columns) or NarrowSet (caller-chosen widths)."""
        h = C.c_void_p()
        if hasattr(s, "lengths") and not isinstance(s, SeqSet):
            cs = _cb_set_cols(s)
            self._check(cabi.lib.cb_upload_cols(self._ctx, C.byref(cs), C.byref(h)))
        else:
            cs = _cb_set(s)
            self._check(cabi.lib.cb_upload(self._ctx, C.byref(cs), C.byref(h)))
        return DeviceSet(self, h, s.n, s.n_reps)

    def rehash(self, d: DeviceSet):
        self._check(cabi.lib.cb_rehash(self._ctx, d.handle))

    def hashes(self, d: DeviceSet) -> np.ndarray:
        out = np.zeros(d.n, dtype=np.uint64)
        self._check(cabi.lib.cb_get_hashes(self._ctx, d.handle, _ptr(out)))
        return out

    def build_b(self, d: DeviceSet):
        self._check(cabi.lib.cb_build_b(self._ctx, d.handle))

    def set_b(self, s):
        if isinstance(s, SeqSet):
            cs = _cb_set(s)
            self._check(cabi.lib.cb_set_b(self._ctx, C.byref(cs)))
        else:
            cs = _cb_set_cols(s)
            self._check(cabi.lib.cb_set_b_cols(self._ctx, C.byref(cs)))

    # -- multi-GPU: NCCL communicator inside the library (csrc/comm.cu)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = cabi.lib.cb_comm_unique_id(buf)
        if rc:
            raise EngineError(rc, cabi.lib.cb_global_error().decode())
        return buf.raw

    def comm_init_rank(self, unique_id: bytes, rank: int, world: int):
        """Collective: every rank's engine calls it with the id made by one of them."""
        self._check(cabi.lib.cb_comm_init_rank(self._ctx, C.c_char_p(unique_id), rank, world))

    def comm_rank(self):
        r, w = C.c_int(), C.c_int()
        self._check(cabi.lib.cb_comm_rank(self._ctx, C.byref(r), C.byref(w)))
        return r.value, w.value

    def set_b_sharded(self, shard, n_total: int):
        """shard: NarrowSet of this rank's cb_shard_range of set B (n_reps of the WHOLE set).
        Collective: upload 1/world over PCIe, all-gather over NVLink, build locally."""
        cs = _cb_set_cols(shard)
        self._check(cabi.lib.cb_set_b_sharded(self._ctx, C.byref(cs), n_total))

    def allreduce_matrix(self):
        """Collective: sum of the partial matrices of all ranks, in place (ncclAllReduce)."""
        self._check(cabi.lib.cb_allreduce_matrix(self._ctx))

    def resident_b(self) -> Optional[DeviceSet]:
        """The context's set B as a DeviceSet (not to be freed), for self-comparison runs."""
        h = cabi.lib.cb_resident_b(self._ctx)
        return DeviceSet(self, None if not h else C.c_void_p(h), 0, 0) if h else None

    def dups_b(self) -> int:
        return int(cabi.lib.cb_dups_b(self._ctx))

    def count_dups(self, d: DeviceSet) -> int:
        out = C.c_uint64()
        self._check(cabi.lib.cb_count_dups(self._ctx, d.handle, C.byref(out)))
        return int(out.value)

    def dedup(self, d: DeviceSet):
        """`compairr -z` on a resident set (src/dedup.cc): (leader index per sequence, group count at
        each leader, members merged away)."""
        lead = np.zeros(d.n, dtype=np.uint32)
        cnt = np.zeros(d.n, dtype=np.uint64)
        merged = C.c_uint64()
        self._check(cabi.lib.cb_dedup(self._ctx, d.handle, _ptr(lead), _ptr(cnt), C.byref(merged)))
        return lead, cnt, int(merged.value)

    def cluster(self, d: DeviceSet):
        """`compairr -c` on a resident set (src/cluster.cc): per output row (sequence index, 1-based
        cluster number, cluster size), plus {"clusters", "edges"}."""
        order = np.zeros(d.n, dtype=np.uint32)
        no = np.zeros(d.n, dtype=np.uint32)
        size = np.zeros(d.n, dtype=np.uint32)
        ncl, ned = C.c_uint64(), C.c_uint64()
        self._check(cabi.lib.cb_cluster(self._ctx, d.handle, _ptr(order), _ptr(no), _ptr(size),
                                        C.byref(ncl), C.byref(ned)))
        return order, no, size, {"clusters": int(ncl.value), "edges": int(ned.value)}

    def run(self, d: DeviceSet, first: int = 0, count: Optional[int] = None):
        self._check(cabi.lib.cb_run(self._ctx, d.handle, first, d.n - first if count is None else count))

    def run_a(self, s):
        if isinstance(s, SeqSet):
            cs = _cb_set(s)
            self._check(cabi.lib.cb_run_a(self._ctx, C.byref(cs)))
        else:
            cs = _cb_set_cols(s)
            self._check(cabi.lib.cb_run_a_cols(self._ctx, C.byref(cs)))

    # -- results
    def matrix(self) -> np.ndarray:
        r, c = C.c_uint64(), C.c_uint64()
        self._check(cabi.lib.cb_matrix_dims(self._ctx, C.byref(r), C.byref(c)))
        out = np.zeros((r.value, c.value), dtype=np.float64)
        self._check(cabi.lib.cb_get_matrix(self._ctx, _ptr(out), out.size))
        return out

    def set_matrix(self, m: np.ndarray):
        m = np.ascontiguousarray(m, dtype=np.float64)
        self._check(cabi.lib.cb_set_matrix(self._ctx, _ptr(m), m.size))

    def bind_matrix(self, device_ptr: int, rows: int, cols: int):
        """Accumulate into a caller-owned device buffer (e.g. a torch tensor's data_ptr())."""
        self._check(cabi.lib.cb_bind_matrix(self._ctx, C.c_void_p(device_ptr), rows, cols))

    def clear_matrix(self):
        self._check(cabi.lib.cb_clear_matrix(self._ctx))

    def matrix_device_ptr(self) -> int:
        return int(cabi.lib.cb_matrix_device(self._ctx) or 0)

    def drain_pairs(self) -> np.ndarray:
        n = C.c_uint64()
        self._check(cabi.lib.cb_pairs_pending(self._ctx, C.byref(n)))
        out = np.zeros((n.value, 2), dtype=np.uint64)
        got = C.c_size_t()
        self._check(cabi.lib.cb_drain_pairs(self._ctx, _ptr(out), n.value, C.byref(got)))
        return out[: got.value]

    def stats(self) -> dict:
        st = cabi.cb_stats()
        self._check(cabi.lib.cb_get_stats(self._ctx, C.byref(st)))
        return st.as_dict()


def comm_init_all(engines):
    """All engines in this process (one per GPU) -> one communicator, engines[i] = rank i."""
    arr = (C.c_void_p * len(engines))(*[e._ctx for e in engines])
    rc = cabi.lib.cb_comm_init_all(arr, len(engines))
    if rc:
        raise EngineError(rc, (cabi.lib.cb_last_error(engines[0]._ctx) or cabi.lib.cb_global_error()).decode())


def shard_range(n_total: int, rank: int, world: int):
    """cb_shard_range: the slice of a set that `rank` uploads in set_b_sharded."""
    f, n = C.c_uint64(), C.c_uint64()
    cabi.lib.cb_shard_range(n_total, rank, world, C.byref(f), C.byref(n))
    return int(f.value), int(n.value)


def probe_count(residues: np.ndarray, sigma: int, differences: int, indels: bool) -> int:
    r = np.ascontiguousarray(residues, dtype=np.uint8)
    return int(cabi.lib.cb_probe_count(_ptr(r), r.size, sigma, differences, int(indels)))


def overlap(a: SeqSet, b: Optional[SeqSet], opts: OverlapOptions):
    """One-call overlap: returns (matrix or None, pairs or None, info dict).
    b=None is the self-comparison of `compairr -m FILE` (overlap.cc:799-825)."""
    self_cmp = b is None
    with Engine(opts, n_reps_a=1 if opts.existence else max(a.n_reps, 1)) as eng:
        db = eng.upload(a if self_cmp else b)
        eng.build_b(db)
        info = {"dups_b": eng.dups_b(), "build": eng.stats()}
        if self_cmp:
            da = db
        else:
            da = eng.upload(a)
            if opts.differences <= 2:
                info["dups_a"] = eng.count_dups(da)
        eng.run(da)
        info["run"] = eng.stats()
        m = None if opts.no_matrix else eng.matrix()
        p = eng.drain_pairs() if opts.want_pairs else None
        if not self_cmp:
            da.free()
        db.free()
        return m, p, info


def dedup(a: SeqSet, opts: OverlapOptions):
    """One-call `compairr -z`: (leader, count, merged) as Engine.dedup."""
    with Engine(opts, n_reps_a=max(a.n_reps, 1)) as eng:
        da = eng.upload(a)
        out = eng.dedup(da)
        da.free()
        return out


def cluster(a: SeqSet, opts: OverlapOptions):
    """One-call `compairr -c`: (order, cluster_no, cluster_size, info) as Engine.cluster."""
    with Engine(opts, n_reps_a=max(a.n_reps, 1)) as eng:
        da = eng.upload(a)
        out = eng.cluster(da)
        da.free()
        return out
