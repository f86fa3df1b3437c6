"""ctypes binding of the C ABI declared in include/compairr_b200.h.

The product path is the CUDA library; there is no fallback.  If libcompairr_b200.so has not been
built (python -c 'import __graft_entry__ as g; g.build()') importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# COMPAIRR_B200_LIB: another build of the same library (A/B runs of kernel variants, tools/)
LIB_PATH = os.environ.get("COMPAIRR_B200_LIB") or os.path.join(_HERE, "libcompairr_b200.so")

ABI_VERSION = 3


class cb_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("alphabet_size", C.c_int32),
        ("differences", C.c_int32), ("indels", C.c_int32), ("ignore_genes", C.c_int32),
        ("ignore_counts", C.c_int32), ("score", C.c_int32), ("mode", C.c_int32),
        ("no_matrix", C.c_int32), ("want_pairs", C.c_int32), ("n_reps_a", C.c_uint32),
        ("seed", C.c_uint64), ("bloom_bits_per_key_x16", C.c_uint32), ("table_load_pct", C.c_uint32),
        ("pairs_capacity", C.c_uint64), ("flags", C.c_uint32), ("queue_capacity", C.c_uint32),
    ]


class cb_set(C.Structure):
    _fields_ = [
        ("n", C.c_uint64), ("residues", C.c_void_p), ("offsets", C.c_void_p),
        ("v_gene", C.c_void_p), ("j_gene", C.c_void_p), ("rep", C.c_void_p), ("count", C.c_void_p),
        ("n_reps", C.c_uint32), ("longest", C.c_uint32), ("index_base", C.c_uint64),
    ]


class cb_col(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("reserved", C.c_uint32)]


class cb_set_cols(C.Structure):
    _fields_ = [
        ("n", C.c_uint64), ("residues", C.c_void_p), ("offsets", cb_col), ("lengths", cb_col),
        ("v_gene", cb_col), ("j_gene", cb_col), ("rep", cb_col), ("count", cb_col),
        ("n_reps", C.c_uint32), ("longest", C.c_uint32), ("index_base", C.c_uint64),
    ]


class cb_stats(C.Structure):
    _fields_ = [
        ("seeds", C.c_uint64), ("probes", C.c_uint64), ("bloom_pass", C.c_uint64),
        ("matches", C.c_uint64), ("pairs", C.c_uint64), ("table_slots", C.c_uint64),
        ("bloom_bytes", C.c_uint64), ("bloom2_bytes", C.c_uint64), ("ms_hash_b", C.c_float), ("ms_build_b", C.c_float),
        ("ms_dups_b", C.c_float), ("ms_hash_a", C.c_float), ("ms_probe", C.c_float),
        ("ms_total_run", C.c_float), ("kernel_launches", C.c_uint32), ("ms_gather_b", C.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/compairr_b200.h declares: (name, restype, argtypes)
P = C.c_void_p
SYMBOLS = [
    ("cb_global_error", C.c_char_p, []),
    ("cb_abi_version", C.c_int, []),
    ("cb_device_count", C.c_int, []),
    ("cb_create", C.c_int, [C.POINTER(cb_config), C.POINTER(P)]),
    ("cb_destroy", None, [P]),
    ("cb_last_error", C.c_char_p, [P]),
    ("cb_set_stream", C.c_int, [P, P]),
    ("cb_upload", C.c_int, [P, C.POINTER(cb_set), C.POINTER(P)]),
    ("cb_upload_cols", C.c_int, [P, C.POINTER(cb_set_cols), C.POINTER(P)]),
    ("cb_free_set", None, [P, P]),
    ("cb_rehash", C.c_int, [P, P]),
    ("cb_get_hashes", C.c_int, [P, P, P]),
    ("cb_build_b", C.c_int, [P, P]),
    ("cb_dups_b", C.c_uint64, [P]),
    ("cb_resident_b", P, [P]),
    ("cb_count_dups", C.c_int, [P, P, C.POINTER(C.c_uint64)]),
    ("cb_dedup", C.c_int, [P, P, P, P, C.POINTER(C.c_uint64)]),
    ("cb_cluster", C.c_int, [P, P, P, P, P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("cb_run", C.c_int, [P, P, C.c_uint64, C.c_uint64]),
    ("cb_set_b", C.c_int, [P, C.POINTER(cb_set)]),
    ("cb_set_b_cols", C.c_int, [P, C.POINTER(cb_set_cols)]),
    ("cb_run_a", C.c_int, [P, C.POINTER(cb_set)]),
    ("cb_run_a_cols", C.c_int, [P, C.POINTER(cb_set_cols)]),
    ("cb_comm_unique_id", C.c_int, [P]),
    ("cb_comm_init_rank", C.c_int, [P, P, C.c_int, C.c_int]),
    ("cb_comm_init_all", C.c_int, [C.POINTER(P), C.c_int]),
    ("cb_comm_rank", C.c_int, [P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("cb_shard_range", None, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("cb_set_b_sharded", C.c_int, [P, C.POINTER(cb_set_cols), C.c_uint64]),
    ("cb_allreduce_matrix", C.c_int, [P]),
    ("cb_matrix_dims", C.c_int, [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("cb_get_matrix", C.c_int, [P, P, C.c_size_t]),
    ("cb_clear_matrix", C.c_int, [P]),
    ("cb_matrix_device", P, [P]),
    ("cb_bind_matrix", C.c_int, [P, P, C.c_uint64, C.c_uint64]),
    ("cb_set_matrix", C.c_int, [P, P, C.c_size_t]),
    ("cb_pairs_pending", C.c_int, [P, C.POINTER(C.c_uint64)]),
    ("cb_drain_pairs", C.c_int, [P, P, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("cb_get_stats", C.c_int, [P, C.POINTER(cb_stats)]),
    ("cb_probe_count", C.c_uint64, [P, C.c_uint32, C.c_int, C.c_int, C.c_int]),
]


def _preload_nccl():
    """libcompairr_b200.so links libnccl.so.2.  In a process that also imports torch the NCCL that
    torch was built against (the wheel's nvidia/nccl/lib) must be the one both use, whichever of the
    two is imported first: map it before our library so the loader resolves the soname to it."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            so = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(so):
                C.CDLL(so, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass   # the system libnccl.so.2 is used


def load():
    _preload_nccl()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. This package has no CPU fallback; build the CUDA library "
            "with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C compairr_b200/csrc`.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.cb_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.cb_abi_version()}, binding {ABI_VERSION}")
    return lib


lib = load()
