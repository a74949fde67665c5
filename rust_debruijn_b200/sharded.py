"""Host-side planning of the multi-GPU path, as Python views of the library's own helpers, plus the torch.distributed
convenience wrapper.

Since round 2 the whole multi-GPU data path lives inside libdbg_b200.so (csrc/multi.cu, csrc/shard_compress.cu: NCCL
all-to-alls, CUDA IPC peer windows, bucket-sharded table); see rust_debruijn_b200/multi.py for the entry points.  What is
left here are the pure planning functions — the SAME C functions the library uses (dbg_plan_owner_bounds,
dbg_plan_quantile_cuts), so the world_size-2/3 gloo tests on CPU exercise the code that runs on the GPUs — and a gloo-capable
count exchange used by those tests."""
import ctypes as C
import math

import numpy as np

from . import _lib


def owner_bounds(n_buckets, world):
    """Rank r owns the contiguous bucket range [bounds[r], bounds[r+1]); owner(b) = (b * world) >> bits for 2^bits buckets."""
    out = (C.c_uint64 * (world + 1))()
    st = _lib.lib().dbg_plan_owner_bounds(n_buckets, world, out)
    if st != 0:
        raise ValueError("bad arguments")
    return [int(x) for x in out]


def split_by_owner(counts, world):
    """Per-destination slices of this rank's bucket histogram (list of uint32 arrays)."""
    b = owner_bounds(len(counts), world)
    return [np.ascontiguousarray(counts[b[r]:b[r + 1]]) for r in range(world)]


def min_bucket_bits(world):
    return max(0, math.ceil(math.log2(world))) if world > 1 else 0


def key_range_splitters(hist_total, world):
    """Boundaries (in units of histogram bins) that cut a global histogram into `world` ranges of ~equal mass:
    world + 1 bin indices, first 0, last len(hist) (the library's quantile cuts for the seed-key ranges)."""
    h = np.ascontiguousarray(hist_total, np.uint64)
    out = (C.c_uint64 * (world + 1))()
    st = _lib.lib().dbg_plan_quantile_cuts(h.ctypes.data_as(C.POINTER(C.c_uint64)), len(h), world, out)
    if st != 0:
        raise ValueError("bad arguments")
    return [int(x) for x in out]


def local_bounds(hist_local, cuts):
    """Slice boundaries of a rank's ASCENDING keys for destination ranges given as bin cuts (what the library derives from the
    rank's own histogram: destination r gets the keys whose bin lies in [cuts[r], cuts[r+1]))."""
    csum = np.concatenate([[0], np.cumsum(np.asarray(hist_local, np.int64))])
    return [int(csum[c]) for c in cuts]


def exchange_counts(per_dst, group=None, device="cpu"):
    """All-to-all of the per-bucket record counts: per_dst[r] goes to rank r; returns the list indexed by
    source rank of counts for THIS rank's buckets (gloo: emulated with gathers; the GPUs do this inside the library)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    n_mine = len(per_dst[rank])
    send = [torch.from_numpy(x.astype(np.int64)).to(device) for x in per_dst]
    recv = [torch.empty(n_mine, dtype=torch.int64, device=device) for _ in range(world)]
    for dst in range(world):
        bufs = [torch.empty_like(send[dst]) for _ in range(world)] if rank == dst else None
        dist.gather(send[dst], bufs, dst=dst, group=group)
        if rank == dst:
            for s in range(world):
                recv[s].copy_(bufs[s])
    return [r.cpu().numpy().astype(np.uint32) for r in recv]


def reads_to_graph_sharded(seqs, summarizer, spec, stranded=False, k=31, group=None, comm=None):
    """filter_kmers + compress_kmers_with_hash over all ranks of an initialised torch.distributed group (collective).
    Returns this rank's MultiGraph (see multi.py).  Pass `comm` to reuse a communicator across calls."""
    from .multi import Comm
    own = comm is None
    if own:
        comm = Comm.from_torch(seqs.ctx, group)
    try:
        return comm.reads_to_graph(seqs, summarizer, spec, stranded=stranded, k=k)
    finally:
        if own:
            comm.close()


def exchange_layout(all_counts, me):
    """Where sender `me` stores its records of every bucket inside the owners' windows (dbg_plan_exchange_layout): all_counts is the
    all-gathered [world, n_buckets] uint32 matrix.  Returns (dst_off[n_buckets] in records, recv_total[world])."""
    all_counts = np.ascontiguousarray(all_counts, np.uint32)
    world, nb = all_counts.shape
    dst_off = np.zeros(nb, np.uint64)
    recv_total = np.zeros(world, np.uint64)
    st = _lib.lib().dbg_plan_exchange_layout(all_counts.ctypes.data_as(C.c_void_p), world, me, nb, dst_off.ctypes.data_as(C.c_void_p),
                                             recv_total.ctypes.data_as(C.c_void_p))
    if st != 0:
        raise ValueError("bad arguments")
    return dst_off, recv_total
