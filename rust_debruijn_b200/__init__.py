"""debruijn-b200: B200-native read -> unitig path of 10XGenomics/rust-debruijn behind a C ABI.

The package holds only what the path needs: csrc/ (hand-written sm_100a kernels + the C ABI,
built into libdbg_b200.so) and this host-side mirror of the reference interface (api.py)."""
from ._lib import DbgError, SO_PATH, build  # noqa: F401
from .api import (ADD_MOD_65535, MAX, SAT_ADD, WRAP_ADD, BaseGraph, Context, CountFilter, CountFilterSet, Exts, KmerTable,  # noqa: F401
                  PackedDnaStringSet, ScmapCompress, SeqSet, SimpleCompress, compress_graph, compress_kmers, compress_kmers_with_hash,
                  default_context, filter_kmers, msp_kmer_buckets, msp_sequence, reads_to_graph, remove_censored_exts,
                  remove_censored_exts_sharded, table_from_host)
