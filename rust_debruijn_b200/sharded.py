"""Multi-GPU form of the path: one process per GPU, torch.distributed for the plumbing.

filter_kmers shards by MSP bucket — the reference's own sharded flow (src/test.rs:433-456: msp_sequence
-> per-shard filter_kmers): every rank cuts ITS reads into super-k-mer records with the same plan, one
all-to-all ships each bucket range to the rank that owns it, and each rank counts its buckets.  All
occurrences of a canonical k-mer share a bucket, so the per-rank tables are disjoint and their union is
exactly the unsharded table.

compress_kmers does not shard with a single exchange (unitigs cross buckets): round 1 gathers the valid
k-mers (V ~ 0.03 N) to every rank and runs the single-GPU compression there ("replicas only for S3-S6",
SURVEY §8e) — every rank ends with the same, complete BaseGraph.

The pure planning helpers (owner_bounds, split_by_owner, exchange_counts) use only torch CPU/any-backend
collectives and are covered by world_size-2 gloo tests on CPU."""
import ctypes as C
import math

import numpy as np

from .api import BaseGraph, KmerTable


def owner_bounds(n_buckets, world):
    """Rank r owns the contiguous bucket range [bounds[r], bounds[r+1])."""
    return [(r * n_buckets) // world for r in range(world + 1)]


def split_by_owner(counts, world):
    """Per-destination slices of this rank's bucket histogram (list of uint32 arrays)."""
    b = owner_bounds(len(counts), world)
    return [np.ascontiguousarray(counts[b[r]:b[r + 1]]) for r in range(world)]


def min_bucket_bits(world):
    return max(0, math.ceil(math.log2(world))) if world > 1 else 0


def exchange_counts(per_dst, group=None, device="cpu"):
    """All-to-all of the per-bucket record counts: per_dst[r] goes to rank r; returns the list indexed by
    source rank of counts for THIS rank's buckets."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    n_mine = len(per_dst[rank])
    send = [torch.from_numpy(x.astype(np.int64)).to(device) for x in per_dst]
    recv = [torch.empty(n_mine, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_to_all(recv, send, group=group) if device != "cpu" else _all_to_all_via_gather(recv, send, group)
    return [r.cpu().numpy().astype(np.uint32) for r in recv]


def _all_to_all_via_gather(recv, send, group):
    """gloo has no all_to_all: emulate with one gather-style exchange per destination (CPU tests only)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    for dst in range(world):
        bufs = [torch.empty_like(send[dst]) for _ in range(world)] if rank == dst else None
        dist.gather(send[dst], bufs, dst=dst, group=group)
        if rank == dst:
            for s in range(world):
                recv[s].copy_(bufs[s])


class _DevView:
    """Zero-copy torch view of a device buffer owned by the library (__cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _as_tensor(ptr, nbytes, device):
    import torch
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def filter_kmers_sharded(seqs, summarizer, stranded, k=31, group=None, report_all_kmers=False, timings=None):
    """filter::filter_kmers over ALL ranks' sequences; returns this rank's shard of the table (its buckets)."""
    import torch
    import torch.distributed as dist
    ctx, L = seqs.ctx, seqs.ctx._L
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    n_local = L.dbg_seqset_count_kmers(ctx._h, k, seqs._h)
    tot = torch.tensor([n_local], dtype=torch.int64, device=dev)
    dist.all_reduce(tot, group=group)
    n_total = int(tot.item())
    p, bits = C.c_int(), C.c_int()
    ctx.check(L.dbg_plan_filter(ctx._h, k, n_total, C.byref(p), C.byref(bits)))
    bbits = max(bits.value, min_bucket_bits(world))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
    if ev:
        ev[0].record()
    part = C.c_void_p()
    ctx.check(L.dbg_partition_reads(ctx._h, k, seqs._h, int(bool(stranded)), p.value, bbits, C.byref(part)))
    try:
        nb = 1 << bbits
        counts = np.zeros(nb, np.uint32)
        ctx.check(L.dbg_partition_bucket_counts(part, C.c_void_p(counts.ctypes.data)))
        per_dst = split_by_owner(counts, world)
        recv_counts = exchange_counts(per_dst, group, device=dev)
        rec_bytes = L.dbg_partition_record_bytes(part)
        send_rec = [int(x.sum(dtype=np.uint64)) for x in per_dst]
        recv_rec = [int(x.sum(dtype=np.uint64)) for x in recv_counts]
        send = _as_tensor(L.dbg_partition_records_dev(part), L.dbg_partition_n_records(part) * rec_bytes, dev)
        recv = torch.empty(sum(recv_rec) * rec_bytes, dtype=torch.uint8, device=dev)
        if ev:
            ev[1].record()
        # the single data-path collective: super-k-mer records by owning rank, over NCCL / NVLink
        dist.all_to_all_single(recv, send, [r * rec_bytes for r in recv_rec], [s * rec_bytes for s in send_rec], group=group)
        if ev:
            ev[2].record()
        torch.cuda.synchronize(dev)
    finally:
        L.dbg_partition_free(part)
    h_counts = np.ascontiguousarray(np.stack(recv_counts).astype(np.uint32))
    th = C.c_void_p()
    ctx.check(L.dbg_filter_from_records(ctx._h, k, C.c_void_p(recv.data_ptr()), sum(recv_rec),
                                        C.c_void_p(h_counts.ctypes.data), world, h_counts.shape[1], n_total,
                                        summarizer.min_kmer_obs, int(bool(stranded)), int(bool(report_all_kmers)),
                                        C.byref(th)))
    if ev:
        ev[3].record()
        torch.cuda.synchronize(dev)
        timings.update(ms_partition=ev[0].elapsed_time(ev[1]), ms_exchange=ev[1].elapsed_time(ev[2]),
                       ms_count_sort=ev[2].elapsed_time(ev[3]), exchange_bytes_sent=sum(send_rec) * rec_bytes,
                       n_input_total=n_total, bucket_bits=bbits)
    return KmerTable(ctx, th)


def key_range_splitters(hist_total, world):
    """Boundaries (in units of histogram bins) that cut a global histogram into `world` ranges of ~equal mass.
    Returns world+1 bin indices, first 0, last len(hist)."""
    csum = np.cumsum(hist_total.astype(np.int64))
    total = int(csum[-1]) if len(csum) else 0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, (total * r) // world, side="left")) + 1 if total else 0)
    cuts.append(len(hist_total))
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    cuts[-1] = len(hist_total)
    return [min(c, len(hist_total)) for c in cuts]


def gather_table(table, group=None):
    """Union of every rank's (disjoint, ascending) shard on every rank, ascending.

    A re-sort of the gathered table would cost every rank P times the single-GPU sort.  Instead the shards are
    first redistributed by KEY RANGE (splitters from an all-reduced histogram of the top 16 key bits, one
    all-to-all), each rank sorts only its range (V/P k-mers), and the ranges are all-gathered in rank order —
    already globally ascending, so the table is adopted without sorting (dbg_table_from_device_sorted)."""
    import torch
    import torch.distributed as dist
    ctx, L = table.ctx, table.ctx._L
    world = dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    k, n = table.k, len(table)
    two = k > 32
    lo_p, hi_p, ex_p, cn_p = (C.c_void_p() for _ in range(4))
    ctx.check(L.dbg_table_device_ptrs(table._h, C.byref(lo_p), C.byref(hi_p), C.byref(ex_p), C.byref(cn_p)))
    ctx.synchronize()
    e64 = torch.empty(0, dtype=torch.int64, device=dev)
    lo = _as_tensor(lo_p.value, n * 8, dev).view(torch.int64) if n else e64
    hi = (_as_tensor(hi_p.value, n * 8, dev).view(torch.int64) if n else e64) if two else None
    ex = _as_tensor(ex_p.value, n, dev)
    cn = _as_tensor(cn_p.value, n * 2, dev)
    # ---- splitters: top 16 bits of the 2k-bit key ----
    nbits = 2 * k - 64 if two else 2 * k          # key bits in the most significant word
    top = hi if two else lo
    if nbits >= 16:
        pfx = (top >> (nbits - 16)) & 0xFFFF
    elif two:                                     # K = 33..39: borrow the missing prefix bits from the low word
        miss = 16 - nbits
        pfx = ((top << miss) | ((lo >> (64 - miss)) & ((1 << miss) - 1))) & 0xFFFF
    else:                                         # K < 8
        pfx = (top << (16 - nbits)) & 0xFFFF
    hist = torch.bincount(pfx, minlength=65536)[:65536]
    dist.all_reduce(hist, group=group)
    cuts = key_range_splitters(hist.cpu().numpy(), world)
    # this rank's shard is ascending: destination r gets the contiguous slice with prefix in [cuts[r], cuts[r+1])
    bounds = torch.searchsorted(pfx, torch.tensor(cuts, dtype=torch.int64, device=dev), right=False)
    bounds[-1] = n
    send_n = [int(x) for x in (bounds[1:] - bounds[:-1]).tolist()]
    sn = torch.tensor(send_n, dtype=torch.int64, device=dev)
    rn = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rn, sn, group=group)
    recv_n = [int(x) for x in rn.tolist()]
    m = sum(recv_n)

    def a2a(t, itemsize):
        out = torch.empty(m * itemsize, dtype=torch.uint8, device=dev)
        dist.all_to_all_single(out, t.view(torch.uint8), [x * itemsize for x in recv_n], [x * itemsize for x in send_n],
                               group=group)
        return out

    r_lo, r_ex, r_cn = a2a(lo, 8), a2a(ex, 1), a2a(cn, 2)
    r_hi = a2a(hi, 8) if two else None
    torch.cuda.synchronize(dev)
    # ---- sort this key range (P ascending runs -> one): V/P k-mers ----
    th = C.c_void_p()
    ctx.check(L.dbg_table_from_device(ctx._h, k, m, C.c_void_p(r_lo.data_ptr()),
                                      C.c_void_p(r_hi.data_ptr()) if two else None, C.c_void_p(r_ex.data_ptr()),
                                      C.c_void_p(r_cn.data_ptr()), C.byref(th)))
    piece = KmerTable(ctx, th)
    ctx.check(L.dbg_table_device_ptrs(piece._h, C.byref(lo_p), C.byref(hi_p), C.byref(ex_p), C.byref(cn_p)))
    # ---- all-gather the ordered ranges (padded to the largest) ----
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([m], dtype=torch.int64, device=dev), group=group)
    sizes = [int(x) for x in sizes.tolist()]
    mx, total = max(max(sizes), 1), sum(sizes)

    def gather(ptr, itemsize):
        pad = torch.zeros(mx * itemsize, dtype=torch.uint8, device=dev)
        if m:
            pad[: m * itemsize] = _as_tensor(ptr.value, m * itemsize, dev)
        out = torch.empty(world * mx * itemsize, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, pad, group=group)
        sb = mx * itemsize
        return torch.cat([out[r * sb: r * sb + sizes[r] * itemsize] for r in range(world)])

    g_lo = gather(lo_p, 8)
    g_hi = gather(hi_p, 8) if two else None
    g_ex = gather(ex_p, 1)
    g_cn = gather(cn_p, 2)
    torch.cuda.synchronize(dev)
    piece.free()
    th = C.c_void_p()
    ctx.check(L.dbg_table_from_device_sorted(ctx._h, k, total, C.c_void_p(g_lo.data_ptr()),
                                             C.c_void_p(g_hi.data_ptr()) if two else None, C.c_void_p(g_ex.data_ptr()),
                                             C.c_void_p(g_cn.data_ptr()), C.byref(th)))
    full = KmerTable(ctx, th)
    full.piece_sizes = sizes   # rank r's key range = global indices [sum(sizes[:r]), sum(sizes[:r+1]))
    return full


def _all_gather_uneven(t, sizes, itemsize, group, dev):
    """All-gather 1-D uint8 views of different lengths (sizes in items); returns the concatenation in rank order."""
    import torch
    import torch.distributed as dist
    world = len(sizes)
    mx = max(max(sizes), 1) * itemsize
    pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
    n = t.numel()
    if n:
        pad[:n] = t
    out = torch.empty(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(x == sizes[0] for x in sizes):
        return out if mx == sizes[0] * itemsize else torch.cat([out[r * mx: r * mx + sizes[r] * itemsize] for r in range(world)])
    return torch.cat([out[r * mx: r * mx + sizes[r] * itemsize] for r in range(world)])


def compress_sharded(full, stranded, spec, group=None, lmax=1024, timings=None):
    """compression::compress_kmers_with_hash over a table replicated on every rank, with the WORK split by k-mer index
    range: links for the own range (+ all-gather), unitig discovery for the path ends in the own range, node layout
    from the all-gathered (seed, length) pairs, emission of the own unitigs into zeroed full-size arrays, one
    all-reduce (every word has a single writer, so sum == OR).  Every rank returns the complete BaseGraph.
    Unitigs longer than `lmax` k-mers and cycles are not handled here: the function then returns None and the
    caller runs the replicated single-GPU compression instead."""
    import torch
    import torch.distributed as dist
    ctx, L = full.ctx, full.ctx._L
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    k, V = full.k, len(full)
    sizes = getattr(full, "piece_sizes", None)
    if sizes is None or sum(sizes) != V:
        sizes = [(V * (r + 1)) // world - (V * r) // world for r in range(world)]
    v0 = sum(sizes[:rank])
    v1 = v0 + sizes[rank]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if timings is not None else None

    def mark(i):
        if ev:
            ctx.synchronize()
            ev[i].record()

    mark(0)
    # ---- links of the own range, all-gather ----
    nxt_local = torch.empty(2 * (v1 - v0), dtype=torch.int32, device=dev)
    ctx.check(L.dbg_cs_links(ctx._h, full._h, int(bool(stranded)), v0, v1, C.c_void_p(nxt_local.data_ptr())))
    nxt = _all_gather_uneven(nxt_local.view(torch.uint8), [2 * x for x in sizes], 4, group, dev)
    torch.cuda.synchronize(dev)
    mark(1)
    # ---- unitigs whose winning end is in the own range ----
    paths = torch.empty(max(v1 - v0, 1) * 16, dtype=torch.uint8, device=dev)
    n_paths, n_cov = C.c_uint64(), C.c_uint64()
    ctx.check(L.dbg_cs_paths(ctx._h, C.c_void_p(nxt.data_ptr()), v0, v1, lmax, C.c_void_p(paths.data_ptr()), max(v1 - v0, 1),
                             C.byref(n_paths), C.byref(n_cov)))
    np_local = n_paths.value
    tot = torch.tensor([np_local, n_cov.value], dtype=torch.int64, device=dev)
    allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, tot, group=group)
    allc = allc.view(world, 2).tolist()
    if sum(int(x[1]) for x in allc) != V:
        return None   # long unitigs or cycles present: replicated fallback
    counts = [int(x[0]) for x in allc]
    M = sum(counts)
    mark(2)
    # ---- node layout: all-gather (seed, length), sort + scan (replicated; M entries only) ----
    pairs_local = paths.view(torch.int32).view(-1, 4)[:np_local, :2].contiguous()
    pairs = _all_gather_uneven(pairs_local.view(torch.uint8).view(-1), counts, 8, group, dev)
    seed_sorted = torch.empty(max(M, 1), dtype=torch.int64, device=dev)
    start = torch.empty(max(M, 1), dtype=torch.int64, device=dev)
    length = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    nb = C.c_uint64()
    torch.cuda.synchronize(dev)
    ctx.check(L.dbg_cs_layout(ctx._h, k, M, C.c_void_p(pairs.data_ptr()), C.c_void_p(seed_sorted.data_ptr()),
                              C.c_void_p(start.data_ptr()), C.c_void_p(length.data_ptr()), C.byref(nb)))
    n_bases = nb.value
    n_words = (n_bases + 31) // 32
    mark(3)
    # ---- emission of the own unitigs, all-reduce ----
    words = torch.zeros(n_words + 3, dtype=torch.int64, device=dev)
    extsw = torch.zeros(M // 4 + 1, dtype=torch.int32, device=dev)
    data = torch.zeros(max(M, 1), dtype=torch.int16, device=dev)
    torch.cuda.synchronize(dev)
    ctx.check(L.dbg_cs_emit(ctx._h, full._h, C.c_void_p(nxt.data_ptr()), C.c_void_p(paths.data_ptr()), np_local,
                            C.c_void_p(seed_sorted.data_ptr()), C.c_void_p(start.data_ptr()), M, spec.func,
                            C.c_void_p(words.data_ptr()), C.c_void_p(extsw.data_ptr()), C.c_void_p(data.data_ptr())))
    mark(4)
    dist.all_reduce(words, group=group)
    dist.all_reduce(extsw, group=group)
    dist.all_reduce(data.view(torch.uint8), group=group)   # single writer per node: byte-wise sum has no carries
    torch.cuda.synchronize(dev)
    gh = C.c_void_p()
    ctx.check(L.dbg_graph_from_device(ctx._h, k, int(bool(stranded)), M, n_bases, C.c_void_p(words.data_ptr()),
                                      C.c_void_p(start.data_ptr()), C.c_void_p(length.data_ptr()),
                                      C.c_void_p(extsw.data_ptr()), C.c_void_p(data.data_ptr()), C.byref(gh)))
    mark(5)
    if ev:
        torch.cuda.synchronize(dev)
        timings.update(ms_cs_links=ev[0].elapsed_time(ev[1]), ms_cs_paths=ev[1].elapsed_time(ev[2]),
                       ms_cs_layout=ev[2].elapsed_time(ev[3]), ms_cs_emit=ev[3].elapsed_time(ev[4]),
                       ms_cs_allreduce=ev[4].elapsed_time(ev[5]), compress="sharded")
    return BaseGraph(ctx, gh)


def reads_to_graph_sharded(seqs, summarizer, spec, stranded=False, k=31, group=None, timings=None):
    """filter_kmers (bucket-sharded, one all-to-all) -> gather -> compress_kmers_with_hash (replicated)."""
    import torch
    from .api import compress_kmers_with_hash
    shard = filter_kmers_sharded(seqs, summarizer, stranded, k=k, group=group, timings=timings)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t2 = torch.cuda.Event(enable_timing=True)
    t0.record()
    full = gather_table(shard, group)
    t1.record()
    shard.free()
    g = compress_sharded(full, stranded, spec, group=group, timings=timings)
    if g is None:   # long unitigs / cycles: replicated single-GPU compression on every rank
        g = compress_kmers_with_hash(stranded, spec, full)
        if timings is not None:
            timings["compress"] = "replicated"
    t2.record()
    if timings is not None:
        torch.cuda.synchronize()
        timings.update(ms_gather=t0.elapsed_time(t1), ms_compress=t1.elapsed_time(t2), n_valid_total=len(full))
    full.free()
    return g
