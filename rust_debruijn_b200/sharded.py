"""Multi-GPU form of the path: one process per GPU, torch.distributed for the plumbing.

filter_kmers shards by MSP bucket — the reference's own sharded flow (src/test.rs:433-456: msp_sequence
-> per-shard filter_kmers): every rank cuts ITS reads into super-k-mer records with the same plan, one
all-to-all ships each bucket range to the rank that owns it, and each rank counts its buckets.  All
occurrences of a canonical k-mer share a bucket, so the per-rank tables are disjoint and their union is
exactly the unsharded table.

compress_kmers does not shard with a single exchange (unitigs cross buckets): the valid k-mers (V ~ 0.03 N) are
redistributed by key range, sorted per range and replicated; the compression WORK is then split by k-mer index
range / seed range with the single-GPU fast-path kernels (link pairs all-gathered, path records shipped to the
rank owning their seed range).  Every rank then holds a contiguous run of nodes of the final order: it keeps it
(replicate=False; the runs concatenated in rank order are the single-GPU BaseGraph bit for bit) or the runs are
all-reduced into the complete graph on every rank (replicate=True).  Long unitigs / cycles fall back to the
replicated single-GPU compression.

Every torch op and collective runs under the library's own CUDA stream (_lib_stream), so library kernels, NCCL and
torch are stream-ordered without host synchronisation.  The pure planning helpers (owner_bounds, split_by_owner,
exchange_counts, key_range_splitters, balanced_seed_bounds) use only torch CPU/any-backend collectives and are
covered by world_size-2/3 gloo tests on CPU."""
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
    st = _lib_stream(ctx)
    with torch.cuda.stream(st):
        n_local = L.dbg_seqset_count_kmers(ctx._h, k, seqs._h)
        tot = torch.tensor([n_local], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=group)
        n_total = int(tot.item())
        p, bits = C.c_int(), C.c_int()
        ctx.check(L.dbg_plan_filter(ctx._h, k, n_total, C.byref(p), C.byref(bits)))
        bbits = max(bits.value, min_bucket_bits(world))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
        if ev:
            ev[0].record(st)
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
                ev[1].record(st)
            # the single data-path collective: super-k-mer records by owning rank, over NCCL / NVLink
            dist.all_to_all_single(recv, send, [r * rec_bytes for r in recv_rec], [s * rec_bytes for s in send_rec], group=group)
            if ev:
                ev[2].record(st)
        finally:
            L.dbg_partition_free(part)   # stream-ordered free on the library stream: after the all-to-all
        h_counts = np.ascontiguousarray(np.stack(recv_counts).astype(np.uint32))
        th = C.c_void_p()
        ctx.check(L.dbg_filter_from_records(ctx._h, k, C.c_void_p(recv.data_ptr()), sum(recv_rec),
                                            C.c_void_p(h_counts.ctypes.data), world, h_counts.shape[1], n_total,
                                            summarizer.min_kmer_obs, int(bool(stranded)), int(bool(report_all_kmers)),
                                            C.byref(th)))
        if ev:
            ev[3].record(st)
            ctx.synchronize()
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


def _lib_stream(ctx):
    """torch view of the library's own stream: torch kernels and NCCL collectives issued under
    `torch.cuda.stream(...)` of it are ordered with the library's kernels, so no host synchronisation is needed
    between a library call and the collective that ships its output (or consumes its input)."""
    import torch
    st = getattr(ctx, "_torch_stream", None)
    if st is None:
        st = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", ctx.device))
        ctx._torch_stream = st
    return st


def _table_views(table, dev):
    """(lo, hi | None, exts, counts) torch views of a table's device arrays."""
    import torch
    ctx, L = table.ctx, table.ctx._L
    n, two = len(table), table.k > 32
    lo_p, hi_p, ex_p, cn_p = (C.c_void_p() for _ in range(4))
    ctx.check(L.dbg_table_device_ptrs(table._h, C.byref(lo_p), C.byref(hi_p), C.byref(ex_p), C.byref(cn_p)))
    e64 = torch.empty(0, dtype=torch.int64, device=dev)
    lo = _as_tensor(lo_p.value, n * 8, dev).view(torch.int64) if n else e64
    hi = (_as_tensor(hi_p.value, n * 8, dev).view(torch.int64) if n else e64) if two else None
    return lo, hi, _as_tensor(ex_p.value, n, dev), _as_tensor(cn_p.value, n * 2, dev)


def gather_table(table, group=None, timings=None):
    """Union of every rank's (disjoint, ascending) shard on every rank, ascending.

    A re-sort of the gathered table would cost every rank P times the single-GPU sort.  Instead the shards are
    first redistributed by KEY RANGE (splitters from an all-reduced histogram of the top 16 key bits, one
    all-to-all per array), each rank sorts only its range (V/P k-mers), and the ordered ranges are all-gathered
    (one collective per array) and copied into the arrays of the replicated table (dbg_table_alloc) — no re-sort."""
    import torch
    import torch.distributed as dist
    ctx, L = table.ctx, table.ctx._L
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    k, n = table.k, len(table)
    two = k > 32
    st = _lib_stream(ctx)
    marks = []

    def mark(name):
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            marks.append((name, e))

    with torch.cuda.stream(st):
        mark("begin")
        lo, hi, ex, cn = _table_views(table, dev)
        # ---- splitters: histogram of the top bits of the 2k-bit key (device kernel), all-reduced ----
        hbits = min(16, 2 * k)
        hist_local = torch.empty(1 << hbits, dtype=torch.int32, device=dev)
        ctx.check(L.dbg_table_prefix_hist(ctx._h, table._h, hbits, C.c_void_p(hist_local.data_ptr())))
        hist = hist_local.to(torch.int64)
        mark("hist")
        dist.all_reduce(hist, group=group)
        cuts = key_range_splitters(hist.cpu().numpy(), world)
        # this rank's shard is ascending: destination r gets the contiguous slice with prefix in [cuts[r], cuts[r+1])
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(hist_local, 0, dtype=torch.int64)])
        bounds = csum[torch.tensor(cuts, dtype=torch.int64, device=dev)]
        sn = bounds[1:] - bounds[:-1]
        rn = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(rn, sn, group=group)
        both = torch.stack([sn, rn]).tolist()
        send_n, recv_n = [int(x) for x in both[0]], [int(x) for x in both[1]]
        m = sum(recv_n)

        def a2a(t, itemsize):
            out = torch.empty(m * itemsize, dtype=torch.uint8, device=dev)
            dist.all_to_all_single(out, t.view(torch.uint8), [x * itemsize for x in recv_n], [x * itemsize for x in send_n],
                                   group=group)
            return out

        r_lo, r_ex, r_cn = a2a(lo, 8), a2a(ex, 1), a2a(cn, 2)
        r_hi = a2a(hi, 8) if two else None
        mark("a2a")
        # ---- sort this key range (P ascending runs -> one): V/P k-mers ----
        th = C.c_void_p()
        ctx.check(L.dbg_table_from_device(ctx._h, k, m, C.c_void_p(r_lo.data_ptr()),
                                          C.c_void_p(r_hi.data_ptr()) if two else None, C.c_void_p(r_ex.data_ptr()),
                                          C.c_void_p(r_cn.data_ptr()), C.byref(th)))
        piece = KmerTable(ctx, th)
        del r_lo, r_ex, r_cn, r_hi
        mark("sort")
        # ---- every rank's ordered range straight into the replicated table ----
        sizes = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([m], dtype=torch.int64, device=dev), group=group)
        sizes = [int(x) for x in sizes.tolist()]
        total = sum(sizes)
        th = C.c_void_p()
        ctx.check(L.dbg_table_alloc(ctx._h, k, total, C.byref(th)))
        full = KmerTable(ctx, th)
        dst = _table_views(full, dev)
        src = _table_views(piece, dev)
        # one all-gather per array (pieces padded to the largest: they differ by < 1 bin of the splitter histogram), then
        # P slice copies into the table — concurrent transfers over NVSwitch instead of P sequential broadcasts
        mx = max(max(sizes), 1)
        offs = [sum(sizes[:r]) for r in range(world)]
        for d_arr, s_arr, isz in zip(dst, src, (1, 1, 1, 2)):
            if d_arr is None:
                continue
            if world <= 2:   # two ranks: a broadcast each way moves the same bytes without the padded staging copies
                for r in range(world):
                    if sizes[r]:
                        sl = d_arr[offs[r] * isz:(offs[r] + sizes[r]) * isz]
                        if r == rank:
                            sl.copy_(s_arr)
                        dist.broadcast(sl, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
                continue
            es = d_arr.element_size() * isz          # bytes per k-mer in this array (views of exts / counts are uint8)
            pad = torch.empty(mx * es, dtype=torch.uint8, device=dev)
            pad[:m * es] = s_arr.view(torch.uint8)[:m * es]
            out = torch.empty(world * mx * es, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, pad, group=group)
            d8 = d_arr.view(torch.uint8)
            for r in range(world):
                if sizes[r]:
                    d8[offs[r] * es:(offs[r] + sizes[r]) * es] = out[r * mx * es:(r * mx + sizes[r]) * es]
        mark("bcast")
        ctx.synchronize()
        piece.free()
    if timings is not None:
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            timings["ms_g_" + n1] = e0.elapsed_time(e1)
    full.piece_sizes = sizes   # rank r's key range = global indices [sum(sizes[:r]), sum(sizes[:r+1]))
    return full


def balanced_seed_bounds(seeds_sorted, V, world, group=None, nbin=4096):
    """Slice boundaries (world + 1 indices into this rank's ASCENDING seeds) that send every path record to the rank
    owning its seed's range, with ranges cut at the node-count quantiles over ALL ranks: a seed is the minimum index of
    its unitig, so seeds crowd the low indices and equal index ranges would be badly unbalanced.  Cumulative counts of
    the local seeds at nbin + 1 bin edges are all-reduced; every rank derives the same cuts from the global counts."""
    import torch
    import torch.distributed as dist
    dev = seeds_sorted.device
    n = seeds_sorted.numel()
    edges = torch.arange(nbin + 1, dtype=torch.int64, device=dev) * ((V + nbin - 1) // nbin)
    cl = torch.searchsorted(seeds_sorted, edges, right=False)
    cl[-1] = n
    cg = cl.clone()
    dist.all_reduce(cg, group=group)
    targets = (cg[-1] * torch.arange(1, world, dtype=torch.int64, device=dev)) // world
    cut = torch.searchsorted(cg, targets, right=False).clamp_(max=nbin)
    cut = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), cut, torch.full((1,), nbin, dtype=torch.int64, device=dev)])
    return cl[cut]


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
    if all(x == sizes[0] for x in sizes) and mx == sizes[0] * itemsize:
        return out
    return torch.cat([out[r * mx: r * mx + sizes[r] * itemsize] for r in range(world)])


def compress_sharded(full, stranded, spec, group=None, lmax=1024, timings=None, replicate=True):
    """compression::compress_kmers_with_hash over a table replicated on every rank, with the WORK split: links for
    the own k-mer index range (+ all-gather), 16-byte walk records, path discovery for the unitigs whose left end
    lies in the own range (+ all-gather of the path records), node layout (replicated, M entries), emission of the
    own slice of NODES into zeroed full-size arrays, one all-reduce (every bit has a single writer, so sum == OR).
    replicate=True: every rank returns the complete BaseGraph.  replicate=False: every rank returns ITS run of nodes
    (attributes node0 / base0 = its position in the complete graph, n_nodes_total / n_bases_total); the runs
    concatenated in rank order are the complete BaseGraph, bit for bit — no full-size arrays, no all-reduce.
    Unitigs longer than `lmax` k-mers and cycles are not handled here: the function then returns None and the
    caller runs the replicated single-GPU compression instead."""
    import torch
    import torch.distributed as dist
    ctx, L = full.ctx, full.ctx._L
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    k, V = full.k, len(full)
    if spec.func > 3:
        return None   # ScmapCompress: the sharded link stage has no join_test; the caller compresses replicated
    per = max((V + world - 1) // world, 1)          # equal index ranges: the link all-gather needs no padding logic
    v0, v1 = min(rank * per, V), min((rank + 1) * per, V)
    st = _lib_stream(ctx)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if timings is not None else None

    def mark(i):
        if ev:
            ev[i].record(st)

    with torch.cuda.stream(st):
        mark(0)
        # ---- links of the own range, all-gather, 16-byte walk records of the whole table ----
        nxt = torch.empty(world * per * 2, dtype=torch.int32, device=dev)
        mine = nxt[rank * per * 2:(rank + 1) * per * 2]
        ctx.check(L.dbg_cs_links(ctx._h, full._h, int(bool(stranded)), v0, v1, C.c_void_p(mine.data_ptr())))
        dist.all_gather_into_tensor(nxt, mine, group=group)
        rec16 = torch.empty(max(V, 1) * 16, dtype=torch.uint8, device=dev)
        ctx.check(L.dbg_cs_pack(ctx._h, full._h, C.c_void_p(nxt.data_ptr()), C.c_void_p(rec16.data_ptr())))
        mark(1)
        # ---- unitigs whose left end is in the own range ----
        cap = max(v1 - v0, 1)
        pkey = torch.empty(cap, dtype=torch.int64, device=dev)
        pval = torch.empty(cap, dtype=torch.int32, device=dev)
        n_paths, n_cov = C.c_uint64(), C.c_uint64()
        ctx.check(L.dbg_cs_discover(ctx._h, C.c_void_p(rec16.data_ptr()), V, v0, v1, lmax, C.c_void_p(pkey.data_ptr()),
                                    C.c_void_p(pval.data_ptr()), cap, C.byref(n_paths), C.byref(n_cov)))
        np_local = n_paths.value
        tot = torch.tensor([np_local, n_cov.value], dtype=torch.int64, device=dev)
        allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, tot, group=group)
        allc = allc.view(world, 2).tolist()
        if sum(int(x[1]) for x in allc) != V:
            return None   # long unitigs or cycles present: replicated fallback
        mark(2)
        # ---- path records go to the rank that owns their SEED's index range: that rank ends up with a contiguous run
        # of nodes in the final (ascending seed) order.  Pre-sort by seed so destinations are contiguous slices. ----
        key_shift = 64 - max((V - 1).bit_length(), 1)
        pk_b = torch.empty(max(np_local, 1), dtype=torch.int64, device=dev)
        pv_b = torch.empty(max(np_local, 1), dtype=torch.int32, device=dev)
        which = C.c_int()
        ctx.check(L.dbg_cs_sort_paths(ctx._h, np_local, C.c_void_p(pkey.data_ptr()), C.c_void_p(pval.data_ptr()),
                                      C.c_void_p(pk_b.data_ptr()), C.c_void_p(pv_b.data_ptr()), C.byref(which)))
        pk_s, pv_s = ((pkey, pval) if which.value == 0 else (pk_b, pv_b))
        pk_s, pv_s = pk_s[:np_local], pv_s[:np_local]
        seeds = (pk_s >> key_shift) & ((1 << (64 - key_shift)) - 1)          # arithmetic shift: mask the sign fill
        bnd = balanced_seed_bounds(seeds, V, world, group)
        sn = bnd[1:] - bnd[:-1]
        rn = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(rn, sn, group=group)
        both = torch.stack([sn, rn]).tolist()
        send_n, recv_n = [int(x) for x in both[0]], [int(x) for x in both[1]]
        m_own = sum(recv_n)
        rk = torch.empty(max(m_own, 1), dtype=torch.int64, device=dev)
        rv = torch.empty(max(m_own, 1), dtype=torch.int32, device=dev)
        dist.all_to_all_single(rk[:m_own], pk_s, recv_n, send_n, group=group)
        dist.all_to_all_single(rv[:m_own], pv_s, recv_n, send_n, group=group)
        # ---- own seed range: sort (P ascending runs -> one), lengths, local offsets ----
        rk_b, rv_b = torch.empty_like(rk), torch.empty_like(rv)
        start_l = torch.empty(max(m_own, 1), dtype=torch.int64, device=dev)
        len_l = torch.empty(max(m_own, 1), dtype=torch.int32, device=dev)
        nb = C.c_uint64()
        ctx.check(L.dbg_cs_layout(ctx._h, k, V, m_own, C.c_void_p(rk.data_ptr()), C.c_void_p(rv.data_ptr()),
                                  C.c_void_p(rk_b.data_ptr()), C.c_void_p(rv_b.data_ptr()), C.byref(which),
                                  C.c_void_p(start_l.data_ptr()), C.c_void_p(len_l.data_ptr()), C.byref(nb)))
        ok_s, ov_s = (rk, rv) if which.value == 0 else (rk_b, rv_b)
        tot = torch.tensor([m_own, nb.value], dtype=torch.int64, device=dev)
        allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, tot, group=group)
        allc = allc.view(world, 2).tolist()
        M = sum(int(x[0]) for x in allc)
        n_bases = sum(int(x[1]) for x in allc)
        node0 = sum(int(x[0]) for x in allc[:rank])
        base0 = sum(int(x[1]) for x in allc[:rank])
        n_words = (n_bases + 31) // 32
        mark(3)
        gh = C.c_void_p()
        if replicate:
            # ---- emission of the own nodes at their global positions, all-reduce ----
            words = torch.zeros(n_words + 3, dtype=torch.int64, device=dev)
            exts = torch.zeros(max(M, 1), dtype=torch.uint8, device=dev)
            data = torch.zeros(max(M, 1), dtype=torch.int16, device=dev)
            start = torch.zeros(max(M, 1), dtype=torch.int64, device=dev)
            length = torch.zeros(max(M, 1), dtype=torch.int32, device=dev)
            ctx.check(L.dbg_cs_emit(ctx._h, full._h, C.c_void_p(rec16.data_ptr()), C.c_void_p(ok_s.data_ptr()),
                                    C.c_void_p(ov_s.data_ptr()), C.c_void_p(start_l.data_ptr()), m_own, node0, base0, spec.func,
                                    C.c_void_p(words.data_ptr()), C.c_void_p(exts.data_ptr()), C.c_void_p(data.data_ptr()),
                                    C.c_void_p(start.data_ptr()), C.c_void_p(length.data_ptr())))
            mark(4)
            for t_ in (words, start, length, exts, data.view(torch.uint8)):   # single writer per element: sum == the value
                dist.all_reduce(t_, group=group)
            ctx.check(L.dbg_graph_from_device(ctx._h, k, int(bool(stranded)), M, n_bases, C.c_void_p(words.data_ptr()),
                                              C.c_void_p(start.data_ptr()), C.c_void_p(length.data_ptr()),
                                              C.c_void_p(exts.data_ptr()), C.c_void_p(data.data_ptr()), C.byref(gh)))
        else:
            # ---- emission of the own run of nodes into arrays of its own size (start rebased to the run) ----
            nb_own = nb.value
            words = torch.zeros((nb_own + 31) // 32 + 3, dtype=torch.int64, device=dev)
            exts = torch.empty(max(m_own, 1), dtype=torch.uint8, device=dev)
            data = torch.empty(max(m_own, 1), dtype=torch.int16, device=dev)
            ctx.check(L.dbg_cs_emit(ctx._h, full._h, C.c_void_p(rec16.data_ptr()), C.c_void_p(ok_s.data_ptr()),
                                    C.c_void_p(ov_s.data_ptr()), C.c_void_p(start_l.data_ptr()), m_own, 0, 0, spec.func,
                                    C.c_void_p(words.data_ptr()), C.c_void_p(exts.data_ptr()), C.c_void_p(data.data_ptr()),
                                    None, None))
            mark(4)
            ctx.check(L.dbg_graph_from_device(ctx._h, k, int(bool(stranded)), m_own, nb_own, C.c_void_p(words.data_ptr()),
                                              C.c_void_p(start_l.data_ptr()), C.c_void_p(len_l.data_ptr()),
                                              C.c_void_p(exts.data_ptr()), C.c_void_p(data.data_ptr()), C.byref(gh)))
        mark(5)
    if ev:
        ctx.synchronize()
        timings.update(ms_cs_links=ev[0].elapsed_time(ev[1]), ms_cs_discover=ev[1].elapsed_time(ev[2]),
                       ms_cs_layout=ev[2].elapsed_time(ev[3]), ms_cs_emit=ev[3].elapsed_time(ev[4]),
                       ms_cs_allreduce=ev[4].elapsed_time(ev[5]), compress="sharded")
    g = BaseGraph(ctx, gh)
    g.node0, g.base0 = (0, 0) if replicate else (node0, base0)
    g.n_nodes_total, g.n_bases_total, g.replicated = M, n_bases, bool(replicate)
    return g


def reads_to_graph_sharded(seqs, summarizer, spec, stranded=False, k=31, group=None, timings=None, replicate=True):
    """filter_kmers (bucket-sharded, one all-to-all) -> table gathered by key range -> compress_kmers_with_hash with the
    work split over the ranks (replicated single-GPU compression when long unitigs / cycles are present).
    replicate=False leaves every rank with its own run of nodes (see compress_sharded)."""
    import torch
    from .api import compress_kmers_with_hash
    ctx = seqs.ctx
    st = _lib_stream(ctx)
    shard = filter_kmers_sharded(seqs, summarizer, stranded, k=k, group=group, timings=timings)
    t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0.record(st)
    full = gather_table(shard, group, timings=timings)
    t1.record(st)
    shard.free()
    g = compress_sharded(full, stranded, spec, group=group, timings=timings, replicate=replicate)
    if g is None:   # long unitigs / cycles: replicated single-GPU compression on every rank (complete graph everywhere)
        g = compress_kmers_with_hash(stranded, spec, full)
        g.node0, g.base0, g.n_nodes_total, g.n_bases_total, g.replicated = 0, 0, len(g), None, True
        if timings is not None:
            timings["compress"] = "replicated"
    t2.record(st)
    if timings is not None:
        ctx.synchronize()
        timings.update(ms_gather=t0.elapsed_time(t1), ms_compress=t1.elapsed_time(t2), n_valid_total=len(full))
    full.free()
    return g
