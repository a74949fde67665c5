"""ctypes loader for libdbg_b200.so (the C ABI declared in include/dbg_b200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present, the calls raise."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# DBG_B200_LIB: another build of the same library (A/B experiments, tools/build_variant.sh); never a fallback
SO_PATH = os.environ.get("DBG_B200_LIB") or os.path.join(_HERE, "libdbg_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, E_BADARG, E_OOM, E_CUDA, E_INCONSISTENT_EXTS, E_INTERNAL = range(6)
STATUS_NAMES = {0: "DBG_OK", 1: "DBG_E_BADARG", 2: "DBG_E_OOM", 3: "DBG_E_CUDA", 4: "DBG_E_INCONSISTENT_EXTS",
                5: "DBG_E_INTERNAL"}


class DbgError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


class Stats(C.Structure):
    _fields_ = ([(n, C.c_uint64) for n in ("n_seqs", "n_input_kmers", "n_records", "n_buckets", "n_distinct", "n_valid",
                                           "n_nodes", "n_bases", "n_bucket_splits", "rank_rounds", "n_cycle_kmers",
                                           "gpu_launches")] +
                [(n, C.c_float) for n in ("ms_partition", "ms_count", "ms_sort", "ms_table", "ms_links", "ms_rank",
                                          "ms_emit")] +
                [("msp_p", C.c_uint32), ("bucket_bits", C.c_uint32)] +
                [(n, C.c_float) for n in ("ms_k_partition", "ms_k_count", "ms_filter_total", "ms_compress_total")] +
                [("n_records_distinct", C.c_uint64), ("n_passes", C.c_uint64), ("direct_partition", C.c_uint64)])

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MultiInfo(C.Structure):
    _fields_ = ([("n_ranks", C.c_uint32), ("rank", C.c_uint32)] +
                [(n, C.c_uint64) for n in ("n_input_total", "n_valid_total", "n_valid_local", "n_nodes_total", "n_bases_total",
                                           "node0", "base0", "n_queries_sent", "exchange_bytes_sent")] +
                [(n, C.c_uint32) for n in ("replicated", "check_ok", "msp_p", "bucket_bits")] +
                [(n, C.c_float) for n in ("ms_partition", "ms_exchange", "ms_count_sort", "ms_links", "ms_discover", "ms_layout",
                                          "ms_emit", "ms_total")])

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(verbose=False):
    """Compile every CUDA source for sm_100a into libdbg_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j4"], stdout=out)
    return SO_PATH


u64p, u32p, u16p, u8p = (C.POINTER(t) for t in (C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8))
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes): every symbol include/dbg_b200.h declares
SIGNATURES = {
    "dbg_ctx_create": (C.c_int, [C.c_int, vpp]),
    "dbg_ctx_destroy": (None, [vp]),
    "dbg_last_error": (C.c_char_p, [vp]),
    "dbg_stats_get": (C.c_int, [vp, C.POINTER(Stats)]),
    "dbg_ctx_set_param": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "dbg_ctx_synchronize": (C.c_int, [vp]),
    "dbg_ctx_stream": (C.c_void_p, [vp]),
    "dbg_seqset_upload": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vpp]),
    "dbg_seqset_upload_uniform": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, vpp]),
    "dbg_seqset_upload_uniform_async": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, vpp]),
    "dbg_seqset_from_ascii": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), vpp]),
    "dbg_seqset_from_ascii_hashn": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), vpp]),
    "dbg_siphash13": (C.c_uint64, [vp, C.c_uint64]),
    "dbg_graph_serialize": (C.c_int, [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "dbg_graph_deserialize": (C.c_int, [vp, C.c_int, vp, C.c_uint64, vpp]),
    "dbg_seqset_wrap_device": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.c_uint32, vpp]),
    "dbg_seqset_synth": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint32, vpp]),
    "dbg_seqset_len": (C.c_uint64, [vp]),
    "dbg_seqset_n_words": (C.c_uint64, [vp]),
    "dbg_seqset_copy_out": (C.c_int, [vp, vp, vp, vp]),
    "dbg_seqset_free": (None, [vp]),
    "dbg_filter_kmers": (C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_int, C.c_int, C.c_uint64, vpp]),
    "dbg_filter_kmers_host": (C.c_int, [vp, C.c_int, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_int,
                                        C.c_int, C.c_uint64, vpp]),
    "dbg_filter_kmers_colorset": (C.c_int, [vp, C.c_int, vp, vp, C.c_uint32, C.c_int, C.c_uint64, vpp]),
    "dbg_table_colorsets": (C.c_int, [vp, vp]),
    "dbg_table_len": (C.c_uint64, [vp]),
    "dbg_table_all_len": (C.c_uint64, [vp]),
    "dbg_table_n_input": (C.c_uint64, [vp]),
    "dbg_table_k": (C.c_int, [vp]),
    "dbg_table_copy_out": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "dbg_table_from_host": (C.c_int, [vp, C.c_int, C.c_uint64, vp, vp, vp, vp, vpp]),
    "dbg_table_free": (None, [vp]),
    "dbg_compress_kmers_with_hash": (C.c_int, [vp, vp, C.c_int, C.c_int, vpp]),
    "dbg_graph_len": (C.c_uint64, [vp]),
    "dbg_graph_n_bases": (C.c_uint64, [vp]),
    "dbg_graph_n_words": (C.c_uint64, [vp]),
    "dbg_graph_stranded": (C.c_int, [vp]),
    "dbg_graph_k": (C.c_int, [vp]),
    "dbg_graph_copy_out": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "dbg_graph_edges": (C.c_int, [vp, vp, vp, vp]),
    "dbg_graph_fix_exts": (C.c_int, [vp, vp, vp]),
    "dbg_graph_is_compressed": (C.c_int, [vp, vp, C.c_int, C.POINTER(C.c_int64)]),
    "dbg_graph_combine": (C.c_int, [vp, vp, C.c_uint32, vpp]),
    "dbg_compress_graph": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, C.c_uint64, vpp]),
    "dbg_graph_free": (None, [vp]),
    "dbg_reads_to_graph": (C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_int, C.c_int, vpp, vpp]),
    "dbg_reads_to_graph_host": (C.c_int, [vp, C.c_int, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_int,
                                          C.c_int, vpp, vpp]),
    "dbg_reads_to_graph_host_uniform": (C.c_int, [vp, C.c_int, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp, C.c_uint32,
                                                  C.c_int, C.c_int, vpp, vpp]),
    "dbg_plan_filter": (C.c_int, [vp, C.c_int, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dbg_seqset_count_kmers": (C.c_uint64, [vp, C.c_int, vp]),
    "dbg_partition_reads": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vpp]),
    "dbg_partition_n_records": (C.c_uint64, [vp]),
    "dbg_partition_n_input": (C.c_uint64, [vp]),
    "dbg_partition_record_bytes": (C.c_uint32, [vp]),
    "dbg_partition_records_dev": (C.c_void_p, [vp]),
    "dbg_partition_bucket_counts": (C.c_int, [vp, vp]),
    "dbg_partition_free": (None, [vp]),
    "dbg_filter_from_records": (C.c_int, [vp, C.c_int, vp, C.c_uint64, vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                          C.c_int, C.c_int, vpp]),
    "dbg_table_alloc": (C.c_int, [vp, C.c_int, C.c_uint64, vpp]),
    "dbg_remove_censored_exts": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "dbg_table_prefix_hist": (C.c_int, [vp, vp, C.c_int, vp]),
    "dbg_table_device_ptrs": (C.c_int, [vp, vpp, vpp, vpp, vpp]),
    "dbg_table_from_device": (C.c_int, [vp, C.c_int, C.c_uint64, vp, vp, vp, vp, vpp]),
    "dbg_table_from_device_sorted": (C.c_int, [vp, C.c_int, C.c_uint64, vp, vp, vp, vp, vpp]),
    "dbg_graph_from_device": (C.c_int, [vp, C.c_int, C.c_int, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, vpp]),
    "dbg_msp_kmer_buckets": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_uint64]),
    "dbg_msp_sequence": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint64), vp, vp, vp, vp, vp]),
    "dbg_comm_unique_id": (C.c_int, [vp]),
    "dbg_comm_create": (C.c_int, [vp, C.c_int, C.c_int, vp, vpp]),
    "dbg_comm_destroy": (None, [vp]),
    "dbg_comm_rank": (C.c_int, [vp]),
    "dbg_comm_size": (C.c_int, [vp]),
    "dbg_comm_transport": (C.c_char_p, [vp]),
    "dbg_reads_to_graph_multi": (C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_int, C.c_int, C.POINTER(MultiInfo), vpp]),
    "dbg_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, vpp]),
    "dbg_multi_destroy": (None, [vp]),
    "dbg_multi_size": (C.c_int, [vp]),
    "dbg_multi_ctx": (C.c_void_p, [vp, C.c_int]),
    "dbg_multi_transport": (C.c_char_p, [vp]),
    "dbg_multi_reads_to_graph": (C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_int, C.c_int, vp, vp]),
    "dbg_plan_owner_bounds": (C.c_int, [C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "dbg_plan_quantile_cuts": (C.c_int, [C.POINTER(C.c_uint64), C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "dbg_plan_exchange_layout": (C.c_int, [vp, C.c_int, C.c_int, C.c_uint64, vp, vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib
