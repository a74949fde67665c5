// compression::compress_kmers_with_hash over a k-mer table that STAYS SHARDED by MSP bucket across the ranks of a
// multi-GPU job (replaces the round-1 design that replicated the whole table on every rank).
//
// The reference's own answer to sharded input is per-shard compression followed by BaseGraph::combine + compress_graph
// (src/test.rs:418-470, src/compression.rs:291-334), which is NOT result-equivalent to the unsharded flow on censored data
// (it runs fix_exts).  Here the stateless link rule of the single-GPU path (try_extend_kmer, src/compression.rs:382-444)
// is evaluated across shards instead, so the union of the per-rank outputs is the single-GPU BaseGraph bit for bit:
//
//   links     per (k-mer, side) of the rank's shard: the neighbour's MSP bucket (a pure function of the k-mer) names the
//             rank that owns it.  Own rank: lookup in the local sorted shard (prefix LUT + binary search).  Other rank:
//             one 16/32-byte query; queries are grouped by owner, exchanged (all-to-all), answered from the owner's shard
//             (index + the Exts nibble on the entered side + data), and the answers patched into the walk records.
//             Consecutive k-mers of a unitig mostly share their minimizer, so only ~1 link in 8 is remote.
//   records   16 bytes per k-mer: both links (index on the owning rank, side, owner rank), count, Exts, first / last
//             base; next to them a copy of the shard's k-mers.  This WINDOW of every rank is mapped into every rank's
//             address space (CUDA IPC over NVLink, or plain peer access inside one process).
//   discover  every path end walks its unitig tracking the smallest K-MER (= the seed, src/compression.rs:574-575: ascending
//             k-mer order is the seed order) and collecting the node; a walk never reads another rank's records: at a rank
//             change the walker itself is shipped (see "WALKER HAND-OFF" below).
//   layout    path records go to the rank owning their seed's key range (quantile cuts of an all-reduced histogram,
//             one CTA-aggregated scatter by destination): that rank sorts them and holds a contiguous run of nodes of
//             the final order.
//   emit      the left-end walker already collected the node (bases, Exts, data) into its message, so the
//             owner of the key range only copies bits into the bit-contiguous PackedDnaStringSet words; nodes too long for
//             a message are re-walked there through the peer windows.
// Only unitigs reachable by end walks (<= lmax k-mers) are handled here; the caller falls back to gathering the table
// and running the single-GPU compression when long unitigs or cycles are present.
#include <algorithm>

#include "common.cuh"
#include "lookup.cuh"

namespace dbg {

template <int W> struct QMsg;
template <> struct __align__(16) QMsg<1> { u64 lo; u32 inc; u32 pad; };
template <> struct __align__(16) QMsg<2> { u64 lo, hi; u32 inc; u32 pad0; u64 pad1; };
// A finished node, shipped to the rank owning its seed's key range: seed k-mer, length, Exts, data and the node's BASES
// (left-aligned 2-bit words), collected by the walker at the node's left end — where most of the unitig is local — so the
// owner of the key range only copies bits.  Nodes longer than 32 * NBW bases carry their left end instead (flag LONG:
// b[0] = port state, aux >> 16 = rank) and are re-walked by the owner through the peer windows.
// meta = length in k-mers | Exts << 16 | flags << 24;  aux = data | rank << 16
static const u32 NODE_LONG = 1u;
template <int W> struct NodeMsg;
template <> struct __align__(16) NodeMsg<1> { u64 lo; u32 meta; u32 aux; u64 b[6]; };          // 64 bytes, <= 192 bases
template <> struct __align__(16) NodeMsg<2> { u64 lo, hi; u32 meta; u32 aux; u64 b[9]; };      // 96 bytes, <= 288 bases
template <int W> struct NodeCfg { static const int NBW = W == 1 ? 6 : 9; };
template <int W> __device__ __forceinline__ Kmer<W> qkey(const QMsg<W>& m) {
    if constexpr (W == 1) return Kmer<1>{m.lo}; else return Kmer<2>{m.lo, m.hi};
}
template <int W> __device__ __forceinline__ Kmer<W> peer_key(const RecPeers& p, u32 rank, u32 idx) {
    if constexpr (W == 1) return Kmer<1>{p.klo[rank][idx]};
    else return Kmer<2>{p.klo[rank][idx], p.khi[rank][idx]};
}

// ---- links of the rank's shard; remote neighbours become queries (flat queue, grouped by owner afterwards) ----
static const int ML_THREADS = 256;
template <int W>
__global__ void __launch_bounds__(ML_THREADS) ms_links_kernel(KP kp, const u64* __restrict__ lo, const u64* __restrict__ hi,
                                                              const u8* __restrict__ exts, const u16* __restrict__ counts, u64 n,
                                                              const u64* __restrict__ lut, int lut_shift, ShardCfg cfg,
                                                              uint4* __restrict__ rec, u64* __restrict__ w_klo, u64* __restrict__ w_khi,
                                                              u64* __restrict__ q_lo, u64* __restrict__ q_hi,
                                                              u32* __restrict__ q_src, u32* __restrict__ q_dst, u64 q_cap,
                                                              u64* __restrict__ ctr /* [0] queries, [1 + r] queries for rank r */,
                                                              u32* __restrict__ err) {
    __shared__ u32 s_n, s_dst[DBG_MAX_RANKS];
    __shared__ u64 s_base;
    if (threadIdx.x == 0) s_n = 0;
    if (threadIdx.x < DBG_MAX_RANKS) s_dst[threadIdx.x] = 0;
    __syncthreads();
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 mask = (1u << cfg.bbits) - 1;
    Kmer<W> qk[2];
    u32 qs[2], qd[2], nq = 0, slot = 0;
    if (i < n) {
        const Kmer<W> key = load_key<W>(lo, hi, i);
        const u32 e = exts[i];
        const bool pal = !cfg.stranded && is_palindrome<W>(kp, key);
        u32 both[2];
        // a neighbour shares all p-mers but one with this k-mer: the shared minima are computed once for both sides
        u32 m_drop_first = 0, m_drop_last = 0;
        if (!pal && (popc4(exts_side(e, 0)) == 1 || popc4(exts_side(e, 1)) == 1))
            kmer_min_scores_shared<W>(kp, key, cfg.p, cfg.stranded != 0, m_drop_first, m_drop_last);
        const u32 pm_mask = cfg.p == 16 ? 0xffffffffu : ((1u << (2 * cfg.p)) - 1);
#pragma unroll
        for (int d = 0; d < 2; d++) {
            u32 succ = NIL;
            const u32 nib = exts_side(e, d);
            if (popc4(nib) == 1 && !pal) {                                     // compression.rs:386
                const u32 base = unique_base(nib);                             // :390
                Kmer<W> nk = d == 0 ? Ops<W>::ext_left(kp, key, base) : Ops<W>::ext_right(kp, key, base);  // :392
                // MSP bucket of the neighbour (strand-symmetric, so taken before canonicalisation): the left neighbour keeps this
                // k-mer's p-mers 0 .. w-2 and gains a new first p-mer, the right neighbour keeps 1 .. w-1 and gains a new last one
                const u32 snew = d == 0 ? pmer_score(kmer_first_pmer<W>(kp, nk, cfg.p), cfg.p, cfg.stranded != 0)
                                        : pmer_score((u32)nk.lo & pm_mask, cfg.p, cfg.stranded != 0);
                const u32 sold = d == 0 ? m_drop_last : m_drop_first;
                const u32 bkt = (snew < sold ? snew : sold) & mask;
                const u32 owner = (u32)(((u64)bkt * (u32)cfg.P) >> cfg.bbits);
                bool flip = false;
                if (!cfg.stranded) {                                           // :396-400
                    const Kmer<W> r = Ops<W>::rc(kp, nk);
                    if (!(nk < r)) { nk = r; flip = true; }
                }
                const int inc = (d ^ 1) ^ (flip ? 1 : 0);                      // :419
                if (owner != (u32)cfg.me) {
                    qk[nq] = nk; qs[nq] = 2u * (u32)i + (u32)d; qd[nq] = owner | ((u32)inc << 8); nq++;
                } else {
                    const u32 j = table_find<W>(lo, hi, lut, lut_shift, nk);   // :410
                    if (j != NIL && j != (u32)i) {                             // not in table / (self => already used) :411-415
                        const bool npal = !cfg.stranded && is_palindrome<W>(kp, nk);   // :403
                        const u32 nnib = exts_side(exts[j], inc);
                        const int cnt = popc4(nnib);                           // :422
                        if (cnt == 0 && !npal) atomicExch(err, 1u);            // :428-434 panic!("unreachable")
                        const bool can_join = !cfg.scmap || counts[i] == counts[j];   // join_test :425, ScmapCompress :92-97
                        if (can_join && cnt == 1 && !npal) {                   // :435
                            Kmer<W> back = inc == 0 ? Ops<W>::ext_left(kp, nk, unique_base(nnib)) : Ops<W>::ext_right(kp, nk, unique_base(nnib));
                            if (!cfg.stranded) { const Kmer<W> r = Ops<W>::rc(kp, back); if (!(back < r)) back = r; }
                            if (back == key) succ = 2u * j + (u32)(inc ^ 1);
                            else atomicExch(err, 2u);
                        }
                    }
                }
            }
            both[d] = succ;
        }
        const u32 meta = (u32)counts[i] | (e << 16) | (Ops<W>::first_base(kp, key) << 24) | (Ops<W>::last_base(kp, key) << 26);
        rec[i] = make_uint4(both[0], both[1], meta, (u32)cfg.me | ((u32)cfg.me << 8));
        w_klo[i] = key.lo;
        if constexpr (W == 2) w_khi[i] = key.hi;
        if (nq) {
            slot = atomicAdd(&s_n, nq);
            for (u32 t = 0; t < nq; t++) atomicAdd(&s_dst[qd[t] & 0xffu], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_n) s_base = atomicAdd(&ctr[0], (u64)s_n);   // one reservation per CTA (same-address atomics serialise)
    if (threadIdx.x < DBG_MAX_RANKS && s_dst[threadIdx.x]) atomicAdd(&ctr[1 + threadIdx.x], (u64)s_dst[threadIdx.x]);
    __syncthreads();
    for (u32 t = 0; t < nq; t++) {
        const u64 pos = s_base + slot + t;
        if (pos < q_cap) {
            q_lo[pos] = qk[t].lo;
            if constexpr (W == 2) q_hi[pos] = qk[t].hi;
            q_src[pos] = qs[t];
            q_dst[pos] = qd[t];
        }
    }
}

// flat queue -> per-owner contiguous segments (message to the owner + the asker's own bookkeeping), one reservation
// per (CTA, owner)
template <int W>
__global__ void __launch_bounds__(256) ms_scatter_queries_kernel(const u64* __restrict__ q_lo, const u64* __restrict__ q_hi,
                                                                  const u32* __restrict__ q_src, const u32* __restrict__ q_dst, u64 nq,
                                                                  const u64* __restrict__ seg_off /* P */, u64* __restrict__ seg_fill /* P, zeroed */,
                                                                  QMsg<W>* __restrict__ msg, u32* __restrict__ src_out) {
    __shared__ u32 s_cnt[DBG_MAX_RANKS];
    __shared__ u64 s_base[DBG_MAX_RANKS];
    if (threadIdx.x < DBG_MAX_RANKS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    constexpr int U = 4;
    u32 dst[U], loc[U];
    const u64 base = (u64)blockIdx.x * (256 * U) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const u64 q = base + (u64)u * 256;
        dst[u] = 0xffffffffu;
        if (q < nq) { dst[u] = q_dst[q]; loc[u] = atomicAdd(&s_cnt[dst[u] & 0xffu], 1u); }
    }
    __syncthreads();
    if (threadIdx.x < DBG_MAX_RANKS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = seg_off[threadIdx.x] + atomicAdd(&seg_fill[threadIdx.x], (u64)s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; u++) {
        const u64 q = base + (u64)u * 256;
        if (dst[u] == 0xffffffffu) continue;
        const u64 pos = s_base[dst[u] & 0xffu] + loc[u];
        QMsg<W> m;
        m.lo = q_lo[q];
        if constexpr (W == 2) { m.hi = q_hi[q]; m.pad0 = 0; m.pad1 = 0; } else { m.pad = 0; }
        m.inc = dst[u] >> 8;
        msg[pos] = m;
        src_out[pos] = q_src[q];
    }
}

// the owner answers: index of the k-mer in its shard (NIL: not there), Exts nibble on the entered side, palindrome flag, data
template <int W>
__global__ void ms_resolve_kernel(KP kp, const u64* __restrict__ lo, const u64* __restrict__ hi, const u8* __restrict__ exts,
                                  const u16* __restrict__ counts, const u64* __restrict__ lut, int lut_shift, int stranded,
                                  const QMsg<W>* __restrict__ msg, u64 nq, uint2* __restrict__ reply) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const QMsg<W> m = msg[q];
    const Kmer<W> nk = qkey<W>(m);
    const u32 j = table_find<W>(lo, hi, lut, lut_shift, nk);
    u32 meta = 0;
    if (j != NIL) {
        const bool npal = !stranded && is_palindrome<W>(kp, nk);
        meta = exts_side(exts[j], (int)(m.inc & 1u)) | (npal ? 16u : 0u) | ((u32)counts[j] << 8);
    }
    reply[q] = make_uint2(j, meta);
}

// the asker patches its walk records with the answers (same rule as the local branch of ms_links_kernel)
template <int W>
__global__ void ms_apply_kernel(KP kp, const u64* __restrict__ lo, const u64* __restrict__ hi, const u16* __restrict__ counts,
                                ShardCfg cfg, const QMsg<W>* __restrict__ msg, const u32* __restrict__ src, const uint2* __restrict__ reply,
                                u64 nq, SegOff seg, uint4* __restrict__ rec, u32* __restrict__ err) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint2 r = reply[q];
    if (r.x == NIL) return;                                                    // not in the table: no link (:411-415)
    int owner = 0;
    while (owner + 1 < cfg.P && q >= seg.off[owner + 1]) owner++;
    const QMsg<W> m = msg[q];
    const Kmer<W> nk = qkey<W>(m);
    const u32 s = src[q], i = s >> 1, d = s & 1u;
    const int inc = (int)(m.inc & 1u);
    const u32 nnib = r.y & 0xfu;
    const bool npal = (r.y & 16u) != 0;
    const int cnt = popc4(nnib);
    if (cnt == 0 && !npal) atomicExch(err, 1u);
    const bool can_join = !cfg.scmap || counts[i] == (u16)(r.y >> 8);
    if (can_join && cnt == 1 && !npal) {
        const Kmer<W> key = load_key<W>(lo, hi, i);
        Kmer<W> back = inc == 0 ? Ops<W>::ext_left(kp, nk, unique_base(nnib)) : Ops<W>::ext_right(kp, nk, unique_base(nnib));
        if (!cfg.stranded) { const Kmer<W> rr = Ops<W>::rc(kp, back); if (!(back < rr)) back = rr; }
        if (back == key) {
            u32* w = reinterpret_cast<u32*>(rec + (u64)i);
            w[d] = 2u * r.x + (u32)(inc ^ 1);
            reinterpret_cast<u8*>(w + 3)[d] = (u8)owner;
        } else {
            atomicExch(err, 2u);
        }
    }
}

// ---- discover + collect by WALKER HAND-OFF ------------------------------------------------------------------------
// Fine-grained loads from another GPU's memory are ruinously slow (measured: a discover kernel that followed links
// through peer-mapped records took 9 ms on 2 GPUs and 412 ms on 8, against 1.7 ms on one), so a walk never leaves its
// rank: when the next k-mer lives elsewhere the WALKER is shipped there — a 32-byte item in a per-destination outbox,
// exchanged in bulk once per round (all-to-all) — and continues on the records of the rank that owns them.
//   Every path end starts a walker that tracks the smallest k-mer (index order = k-mer order inside a shard; k-mers are compared
//   only at a rank change) and the side through which the walk leaves it, and that collects the node as it goes: bases
//   (left-aligned words in registers), end Exts (compression.rs:513-517,534-540), reduced data (:500-511).  The walker that
//   reaches the far end having left the seed "through R" started at the node's LEFT end (compression.rs:574-583, ascending seed
//   order): it holds the finished NodeMsg, which stays on the rank where the walk ended; the walker from the other end is dropped.
// Every walker emits at most ONE thing, so outputs are appended with one reservation per (CTA, destination).
// local part of a walk without collection: from port state `t` (at k-mer t >> 1 of THIS rank, about to be counted) onwards while the links
// stay on this rank.  Returns the state the walk stops at: NIL (path end) or a port on rank `trank`.
struct LocalMin { u32 idx, side; bool any; };
__device__ __forceinline__ u32 walk_local(const uint4* __restrict__ rec, int me, u32 t, u32& trank, u32& cnt, u32 lmax, LocalMin& lm) {
    while (t != NIL && trank == (u32)me && cnt <= lmax) {
        const u32 idx = t >> 1, side = t & 1u;
        const uint4 a = rec[idx];
        cnt++;
        if (!lm.any || idx < lm.idx) { lm.idx = idx; lm.side = side; lm.any = true; }
        t = side ? a.y : a.x;
        trank = (a.w >> (8 * side)) & 0xffu;
    }
    return t;
}

template <int W>
__device__ __forceinline__ Kmer<W> local_key(const u64* __restrict__ klo, const u64* __restrict__ khi, u32 idx) {
    if constexpr (W == 1) return Kmer<1>{klo[idx]}; else return Kmer<2>{klo[idx], khi[idx]};
}

// ---- collecting a node along a walk ----
template <int W>
struct Collector {
    static constexpr int NBW = NodeCfg<W>::NBW;
    u64 bw[NBW];
    u64 acc;
    u32 eb;      // Exts of the node: left nibble set at the first k-mer, right nibble at the last
    u32 done;    // k-mers collected so far
    // k-mer at port state `cur` (leaving side = cur & 1) of this rank; j = its position in the node
    __device__ __forceinline__ void first(const KP& kp, const uint4& r, Kmer<W> key, u32 cur, u32 len) {
        const int K = kp.k;
        const u32 dw = cur & 1u;
        const bool fw = dw == 1u;    // leaving through R while heading right = stored orientation
        if (!fw) key = Ops<W>::rc(kp, key);
        if constexpr (W == 1) {
            bw[0] = key.lo << (64 - 2 * K);
        } else {
            const int sh = 128 - 2 * K;   // 0..62
            bw[0] = sh ? (key.hi << sh) | (key.lo >> (64 - sh)) : key.hi;
            bw[1] = key.lo << sh;
        }
        const u32 e = (r.z >> 16) & 0xffu;
        u32 nib = exts_side(e, (int)(dw ^ 1u));          // left-facing side of the first k-mer (:513-517)
        if (!fw) nib = exts_complement(nib) & 0xfu;
        eb = nib;
        if (len == 1) {
            u32 rn = exts_side(e, (int)dw);
            if (!fw) rn = exts_complement(rn) & 0xfu;
            eb |= rn << 4;
        }
        acc = r.z & 0xffffu;
        done = 1;
    }
    __device__ __forceinline__ void next(const KP& kp, const uint4& r, u32 cur, u32 len, int reduce_op) {
        const u32 dw = cur & 1u;
        const bool fw = dw == 1u;
        const u32 fb = (r.z >> 24) & 3u, lb = (r.z >> 26) & 3u;
        const u64 bb = fw ? lb : 3u - fb;                // last base of the k-mer as it appears in the node
        const u32 pos = (u32)kp.k + done - 1;
        const u64 x = bb << (62 - 2 * (pos & 31));
#pragma unroll
        for (int w = 0; w < NBW; w++) if (w == (int)(pos >> 5)) bw[w] |= x;
        const u64 cnt2 = r.z & 0xffffu;
        if (reduce_op >= DBG_REDUCE_MAX) acc = cnt2 > acc ? cnt2 : acc; else acc += cnt2;   // SCMAP: all equal, max == the value
        if (done == len - 1) {
            u32 rn = exts_side((r.z >> 16) & 0xffu, (int)dw);   // right-facing side of the last k-mer (:534-540)
            if (!fw) rn = exts_complement(rn) & 0xfu;
            eb |= rn << 4;
        }
        done++;
    }
    __device__ __forceinline__ u32 data16(u32 len, int reduce_op) const {
        switch (reduce_op) {
            case DBG_REDUCE_SAT_ADD: return (u32)(acc > 65535 ? 65535 : acc);
            case DBG_REDUCE_WRAP_ADD: return (u32)(acc & 0xffff);
            case DBG_REDUCE_ADD_MOD_65535: return len == 1 ? (u32)acc : (u32)(acc % 65535);   // one k-mer: reduce() never called (:495)
            default: return (u32)acc;
        }
    }
};

// ---- fused walk: ONE pass of walker rounds instead of two (find the seed and the left end, then collect the node).  Both end
// walkers of a path collect the node as they go (bases in their own direction of travel, Exts of their first k-mer, data); the one
// that turns out to have started at the node's left end — it traversed the seed "leaving through R" — arrives at the far end with
// the finished node and appends it to that rank's list, the other one is dropped there.  Half the rounds (each one a kernel, a
// count exchange and an all-to-all), no notices, no emit entries; the price is that every walker carries a node message. ----
// info = origin rank | side through which the walk leaves the current minimum << 8 | bases overflowed the message << 9
template <int W> struct __align__(16) FItem { NodeMsg<W> node; u32 port, done; u64 acc; u32 origin_port, info; u64 pad; };

template <int W>
__device__ __forceinline__ void fwalk_run(const KP& kp, const uint4* __restrict__ rec, const u64* __restrict__ klo, const u64* __restrict__ khi,
                                          int me, FItem<W>& it, bool active, u32 lmax, int reduce_op, WalkOut out) {
    constexpr int NBW = NodeCfg<W>::NBW;
    bool finished = false;
    u32 crank = (u32)me, covered = 0;
    if (active) {
        Collector<W> col;
#pragma unroll
        for (int w = 0; w < NBW; w++) col.bw[w] = it.node.b[w];
        col.acc = it.acc; col.done = it.done; col.eb = (it.node.meta >> 16) & 0xffu;
        bool islong = (it.info >> 9) & 1u;
        u32 cur = it.port;
        LocalMin lm{0, 0, false};
        uint4 r_last = make_uint4(NIL, NIL, 0, 0);
        u32 cur_last = 0;
        while (cur != NIL && crank == (u32)me && col.done <= lmax) {
            const u32 idx = cur >> 1, dw = cur & 1u;
            const uint4 r = rec[idx];
            if (!lm.any || idx < lm.idx) { lm.idx = idx; lm.side = dw; lm.any = true; }
            if (col.done == 0) col.first(kp, r, local_key<W>(klo, khi, idx), cur, 0u);   // (len unknown: the right nibble is set at the end)
            else if (islong || (u32)kp.k + col.done > 32u * NBW) { islong = true; col.done++; }
            else col.next(kp, r, cur, 0u, reduce_op);
            r_last = r; cur_last = cur;
            cur = dw ? r.y : r.x;
            crank = (r.w >> (8 * dw)) & 0xffu;
        }
        // the smallest k-mer seen so far and the side through which the walk left it
        Kmer<W> ckey;
        if constexpr (W == 1) ckey = Kmer<1>{it.node.lo}; else ckey = Kmer<2>{it.node.lo, it.node.hi};
        u32 cside = (it.info >> 8) & 1u;
        const bool had = it.done != 0;
        if (lm.any) {
            const Kmer<W> lk = local_key<W>(klo, khi, lm.idx);
            if (!had || lk < ckey) { ckey = lk; cside = lm.side; }
        }
        it.node.lo = ckey.lo;
        if constexpr (W == 2) it.node.hi = ckey.hi;
        const u32 origin = it.info & 0xffu;
        if (col.done > lmax) active = false;                    // too long for end walks: the caller falls back
        else if (cur == NIL) {
            if (cside != 1u) active = false;                    // started at the right end: the other walker delivers the node
            else {
                finished = true; covered = col.done;
                const u32 dw = cur_last & 1u;
                u32 rn = exts_side((r_last.z >> 16) & 0xffu, (int)dw);   // right-facing side of the last k-mer (:534-540)
                if (dw != 1u) rn = exts_complement(rn) & 0xfu;
                col.eb |= rn << 4;
                if (islong) {   // the owner of the key range re-walks it from its left end through the peer windows
                    it.node.b[0] = it.origin_port;
                    it.node.meta = col.done | ((col.eb & 0xffu) << 16) | (NODE_LONG << 24);
                    it.node.aux = origin << 16;
                } else {
#pragma unroll
                    for (int w = 0; w < NBW; w++) it.node.b[w] = col.bw[w];
                    it.node.meta = col.done | ((col.eb & 0xffu) << 16);
                    it.node.aux = col.data16(col.done, reduce_op) | ((u32)me << 16);
                }
            }
        } else {
#pragma unroll
            for (int w = 0; w < NBW; w++) it.node.b[w] = col.bw[w];
            it.node.meta = (col.eb & 0xffu) << 16;
            it.node.aux = 0;
            it.port = cur; it.done = col.done; it.acc = col.acc;
            it.info = origin | (cside << 8) | ((islong ? 1u : 0u) << 9);
        }
    }
    // finished nodes go to the rank's own list as plain NodeMsg, walkers travel on as FItem
    __shared__ u32 s_cnt[DBG_MAX_RANKS + 1], s_cov;
    __shared__ u64 s_base[DBG_MAX_RANKS + 1];
    if (threadIdx.x <= DBG_MAX_RANKS) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_cov = 0;
    __syncthreads();
    const int dest = finished ? out.P : (int)crank;
    u32 loc = 0;
    if (active) { loc = atomicAdd(&s_cnt[dest], 1u); if (covered) atomicAdd(&s_cov, covered); }
    __syncthreads();
    if (threadIdx.x <= DBG_MAX_RANKS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&out.cursor[threadIdx.x], (u64)s_cnt[threadIdx.x]);
    if (threadIdx.x == 0 && s_cov) atomicAdd(&out.cursor[out.P + 1], (u64)s_cov);
    __syncthreads();
    if (active) {
        const u64 pos = s_base[dest] + loc;
        if (pos >= out.cap[dest]) out.cursor[out.P + 2] = 1;
        else if (finished) reinterpret_cast<NodeMsg<W>*>(out.box[dest])[pos] = it.node;
        else reinterpret_cast<FItem<W>*>(out.box[dest])[pos] = it;
    }
}

// round 0: thread per k-mer of the shard; path ends (and single k-mers) start a walker
template <int W>
__global__ void __launch_bounds__(256) ms_fwalk_start_kernel(KP kp, const uint4* __restrict__ rec, const u64* __restrict__ klo, const u64* __restrict__ khi,
                                                              int me, u64 n, u32 lmax, int reduce_op, WalkOut out) {
    constexpr int NBW = NodeCfg<W>::NBW;
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    FItem<W> it;
    it.node.lo = 0;
    if constexpr (W == 2) it.node.hi = 0;
    it.node.meta = 0; it.node.aux = 0;
#pragma unroll
    for (int w = 0; w < NBW; w++) it.node.b[w] = 0;
    it.port = 0; it.done = 0; it.acc = 0; it.origin_port = 0; it.info = (u32)me; it.pad = 0;
    bool active = false;
    if (v < n) {
        const uint4 a0 = rec[v];
        if (a0.x == NIL || a0.y == NIL) {
            // stored orientation for a single k-mer (heading right = leaving through R), else inwards through the linked side
            const u32 d = (a0.x == NIL && a0.y == NIL) ? 1u : (a0.x == NIL ? 1u : 0u);
            active = true;
            it.port = 2u * (u32)v + d;
            it.origin_port = it.port;
            // a light walk first (no collection): most paths end on this rank, and half of their end walkers started at the right
            // end — those stop here without having collected anything; only left-end walkers and walkers that leave the rank collect
            u32 cnt = 1;
            LocalMin lm{(u32)v, d, true};
            u32 t = d ? a0.y : a0.x;
            u32 trank = (a0.w >> (8 * d)) & 0xffu;
            t = walk_local(rec, me, t, trank, cnt, lmax, lm);
            if (cnt > lmax || (t == NIL && lm.side != 1u)) active = false;
        }
    }
    fwalk_run<W>(kp, rec, klo, khi, me, it, active, lmax, reduce_op, out);
}

// later rounds: thread per received walker
template <int W>
__global__ void __launch_bounds__(256) ms_fwalk_continue_kernel(KP kp, const uint4* __restrict__ rec, const u64* __restrict__ klo, const u64* __restrict__ khi,
                                                                 int me, const FItem<W>* __restrict__ inbox, u64 n_in, u32 lmax, int reduce_op, WalkOut out) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    FItem<W> it;
    const bool active = q < n_in;
    if (active) it = inbox[q];
    else { it.port = NIL; it.done = 0; it.acc = 0; it.info = 0; it.origin_port = 0; it.node.meta = 0; it.node.aux = 0; it.node.lo = 0; }
    fwalk_run<W>(kp, rec, klo, khi, me, it, active, lmax, reduce_op, out);
}

// node messages (any order) -> grouped by the rank owning the seed's key range: destination = range of the seed's top-bits
// bin among the quantile cuts; one reservation per (CTA, destination)
template <int W>
__global__ void __launch_bounds__(256) ms_scatter_nodes_kernel(const NodeMsg<W>* __restrict__ in, u64 m, int P, int bin_shift,
                                                                SegOff cuts /* bins */, SegOff seg /* message offsets */,
                                                                u64* __restrict__ seg_fill, NodeMsg<W>* __restrict__ out) {
    __shared__ u32 s_cnt[DBG_MAX_RANKS];
    __shared__ u64 s_base[DBG_MAX_RANKS];
    if (threadIdx.x < DBG_MAX_RANKS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 dst = 0, loc = 0;
    NodeMsg<W> r;
    if (i < m) {
        r = in[i];
        Kmer<W> key;
        if constexpr (W == 1) key = Kmer<1>{r.lo}; else key = Kmer<2>{r.lo, r.hi};
        const u32 bin = key_prefix<W>(key, bin_shift);
        while ((int)dst + 1 < P && (u64)bin >= cuts.off[dst + 1]) dst++;
        loc = atomicAdd(&s_cnt[dst], 1u);
    }
    __syncthreads();
    if (threadIdx.x < DBG_MAX_RANKS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = seg.off[threadIdx.x] + atomicAdd(&seg_fill[threadIdx.x], (u64)s_cnt[threadIdx.x]);
    __syncthreads();
    if (i < m) out[s_base[dst] + loc] = r;
}
template <int W>
__global__ void ms_unpack_nodes_kernel(const NodeMsg<W>* __restrict__ in, u64 m, u64* __restrict__ k_lo, u64* __restrict__ k_hi, u32* __restrict__ idx) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    k_lo[i] = in[i].lo;
    if constexpr (W == 2) k_hi[i] = in[i].hi;
    idx[i] = (u32)i;
}
template <int W>
__global__ void ms_node_len_kernel(const NodeMsg<W>* __restrict__ in, const u32* __restrict__ idx, u64 m, int K, u64* __restrict__ node_len, u32* __restrict__ out_length) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u64 l = (u64)(in[idx[i]].meta & 0xffffu) + K - 1;
    node_len[i] = l;
    out_length[i] = (u32)l;
}

struct MsNodeWriter {   // same as the single-GPU NodeWriter: interior words are plain stores, the two shared end words atomicOr
    u64* words; u64 first_w, last_w, wi; u64 cur;
    __device__ __forceinline__ void flush() {
        if (wi == first_w || wi == last_w) { if (cur) atomicOr(&words[wi], cur); }
        else words[wi] = cur;
        cur = 0; wi++;
    }
    __device__ __forceinline__ void push(u64 x, int n, u64 pos) {
        const int off = (int)(pos & 31);
        cur |= x >> (2 * off);
        if (off + n >= 32) {
            flush();
            if (off) cur = x << (64 - 2 * off);
        }
    }
};

template <int W>
__global__ void ms_emit_kernel(KP kp, RecPeers peers, const NodeMsg<W>* __restrict__ msgs, const u32* __restrict__ idx, const u64* __restrict__ node_start,
                               u64 m, int reduce_op, u64* __restrict__ words, u8* __restrict__ out_exts, u16* __restrict__ out_data) {
    constexpr int NBW = NodeCfg<W>::NBW;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int K = kp.k;
    const NodeMsg<W> pm = msgs[idx[i]];
    const u32 len = pm.meta & 0xffffu;
    const u64 st = node_start[i];
    const u64 L = (u64)len + K - 1;
    MsNodeWriter nw;
    nw.words = words; nw.first_w = st >> 5; nw.last_w = (st + L - 1) >> 5; nw.wi = nw.first_w; nw.cur = 0;
    u64 pos = st;
    if (!((pm.meta >> 24) & NODE_LONG)) {
        // the node arrived complete: PackedDnaStringSet::add (dna_string.rs:811-821) is a bit copy
#pragma unroll
        for (int w = 0; w < NBW; w++) {
            if ((u64)32 * w < L) {
                const int nb = (int)(L - 32 * w < 32 ? L - 32 * w : 32);
                nw.push(pm.b[w], nb, pos);
                pos += nb;
            }
        }
        if (pos & 31) nw.flush();
        out_exts[i] = (u8)(pm.meta >> 16);
        out_data[i] = (u16)pm.aux;
        return;
    }
    // long node: walk it from its left end through the peer windows
    u32 rank = pm.aux >> 16;
    u32 cur = (u32)pm.b[0];          // at the left end, leaving through its right-facing side
    u64 acc = 0;
    u32 eb = 0;
    {   // first k-mer: all K bases (compression.rs:489-495)
        const u32 dw = cur & 1u;
        const bool fw = dw == 1u;    // leaving through R while heading right = stored orientation
        const uint4 r = peers.rec[rank][cur >> 1];
        Kmer<W> key = peer_key<W>(peers, rank, cur >> 1);
        if (!fw) key = Ops<W>::rc(kp, key);
        if constexpr (W == 1) {
            nw.push(key.lo << (64 - 2 * K), K, pos);
        } else {
            const int sh = 128 - 2 * K;   // 0..62
            const u64 H = sh ? (key.hi << sh) | (key.lo >> (64 - sh)) : key.hi;
            nw.push(H, 32, pos);
            nw.push(key.lo << sh, K - 32, pos + 32);
        }
        pos += K;
        const u32 e = (r.z >> 16) & 0xffu;
        u32 nib = exts_side(e, (int)(dw ^ 1u));          // left-facing side of the first k-mer (:513-517)
        if (!fw) nib = exts_complement(nib) & 0xfu;
        eb = nib;
        if (len == 1) {
            u32 rn = exts_side(e, (int)dw);
            if (!fw) rn = exts_complement(rn) & 0xfu;
            eb |= rn << 4;
        }
        acc = r.z & 0xffffu;
        cur = dw ? r.y : r.x;
        rank = (r.w >> (8 * dw)) & 0xffu;
    }
    for (u32 j = 1; j < len; j++) {
        const u32 dw = cur & 1u;
        const bool fw = dw == 1u;
        const uint4 r = peers.rec[rank][cur >> 1];
        const u32 fb = (r.z >> 24) & 3u, lb = (r.z >> 26) & 3u;
        const u64 b = fw ? lb : 3u - fb;                 // last base of the k-mer as it appears in the node
        nw.push(b << 62, 1, pos);
        pos++;
        const u64 cnt = r.z & 0xffffu;
        if (reduce_op >= DBG_REDUCE_MAX) acc = cnt > acc ? cnt : acc; else acc += cnt;   // SCMAP: all equal, max == the value
        if (j == len - 1) {
            u32 rn = exts_side((r.z >> 16) & 0xffu, (int)dw);   // right-facing side of the last k-mer (:534-540)
            if (!fw) rn = exts_complement(rn) & 0xfu;
            eb |= rn << 4;
        }
        cur = dw ? r.y : r.x;
        rank = (r.w >> (8 * dw)) & 0xffu;
    }
    if (pos & 31) nw.flush();   // partial last word
    out_exts[i] = (u8)eb;
    u16 d;
    switch (reduce_op) {
        case DBG_REDUCE_SAT_ADD: d = (u16)(acc > 65535 ? 65535 : acc); break;
        case DBG_REDUCE_WRAP_ADD: d = (u16)(acc & 0xffff); break;
        case DBG_REDUCE_ADD_MOD_65535: d = len == 1 ? (u16)acc : (u16)(acc % 65535); break;   // one k-mer: reduce() never called (:495)
        default: d = (u16)acc; break;
    }
    out_data[i] = d;
}

// ================================================================================================
// stage entry points (called by multi.cu)
// ================================================================================================
template <int W>
static int ms_links_impl(Ctx* c, const Table* t, ShardCfg cfg, uint4* d_rec, u64* d_klo, u64* d_khi, MsQueries* q) {
    cudaStream_t st = c->stream;
    KP kp = make_kp(t->k);
    const u64 n = t->n;
    q->n_total = 0;
    for (int r = 0; r < DBG_MAX_RANKS; r++) q->n_dst[r] = 0;
    if (n >= (1ull << 31)) DBG_SET_ERR(c, DBG_E_BADARG, "k-mer shard too large for 32-bit port states (%llu)", (unsigned long long)n);
    TRY(q->ctr.alloc_pool(c, 2 + DBG_MAX_RANKS));
    TRY(q->ctr.zero());
    if (n == 0) return DBG_OK;
    TRY(build_prefix_lut<W>(c, t->k, t->lo, t->hi, n, q->lut_cnt, q->lut, &q->lut_shift));
    DBuf<u64> f_lo, f_hi;
    DBuf<u32> f_src, f_dst;
    const u64 cap = 2 * n;
    TRY(f_lo.alloc_pool(c, cap)); TRY(f_src.alloc_pool(c, cap)); TRY(f_dst.alloc_pool(c, cap));
    if (W == 2) TRY(f_hi.alloc_pool(c, cap));
    ms_links_kernel<W><<<grid_for(n, ML_THREADS), ML_THREADS, 0, st>>>(kp, t->lo, t->hi, t->exts, t->counts, n, q->lut.p, q->lut_shift, cfg, d_rec, d_klo, d_khi,
                                                                      f_lo.p, f_hi.p, f_src.p, f_dst.p, cap, q->ctr.p, (u32*)(q->ctr.p + 1 + DBG_MAX_RANKS));
    TRY(check_launch(c, "ms_links"));
    u64 h[2 + DBG_MAX_RANKS];
    TRY(read_u64(c, q->ctr.p, h, 2 + DBG_MAX_RANKS));
    q->n_total = h[0];
    u64 off = 0;
    for (int r = 0; r < cfg.P; r++) { q->n_dst[r] = h[1 + r]; q->off[r] = off; off += h[1 + r]; }
    q->off[cfg.P] = off;
    if (off != q->n_total) DBG_SET_ERR(c, DBG_E_INTERNAL, "query counts do not add up");
    // group by owner
    TRY(q->msg.alloc_pool(c, (q->n_total ? q->n_total : 1) * sizeof(QMsg<W>)));
    TRY(q->src.alloc_pool(c, q->n_total ? q->n_total : 1));
    if (q->n_total) {
        DBuf<u64> d_off, d_fill;
        TRY(d_off.alloc_pool(c, DBG_MAX_RANKS)); TRY(d_fill.alloc_pool(c, DBG_MAX_RANKS));
        TRY(d_fill.zero());
        CU(c, cudaMemcpyAsync(d_off.p, q->off, sizeof(u64) * DBG_MAX_RANKS, cudaMemcpyHostToDevice, st));
        ms_scatter_queries_kernel<W><<<grid_for(q->n_total, 1024), 256, 0, st>>>(f_lo.p, f_hi.p, f_src.p, f_dst.p, q->n_total, d_off.p, d_fill.p,
                                                                                reinterpret_cast<QMsg<W>*>(q->msg.p), q->src.p);
        TRY(check_launch(c, "ms_scatter_queries"));
        TRY(sync(c));   // q->off is host memory read by the asynchronous copy above
    }
    return DBG_OK;
}
int ms_links_dev(Ctx* c, const Table* t, ShardCfg cfg, uint4* d_rec, u64* d_klo, u64* d_khi, MsQueries* q) {
    return t->k <= 32 ? ms_links_impl<1>(c, t, cfg, d_rec, d_klo, d_khi, q) : ms_links_impl<2>(c, t, cfg, d_rec, d_klo, d_khi, q);
}
u32 ms_query_bytes(int k) { return k <= 32 ? (u32)sizeof(QMsg<1>) : (u32)sizeof(QMsg<2>); }
u32 ms_path_bytes(int k) { return k <= 32 ? (u32)sizeof(NodeMsg<1>) : (u32)sizeof(NodeMsg<2>); }

int ms_resolve_dev(Ctx* c, const Table* t, const MsQueries* q, int stranded, const void* d_queries, u64 nq, uint2* d_reply) {
    if (!nq) return DBG_OK;
    KP kp = make_kp(t->k);
    if (t->n == 0) {   // an empty shard owns nothing: every answer is "not there"
        CU(c, cudaMemsetAsync(d_reply, 0xff, nq * sizeof(uint2), c->stream));
        return DBG_OK;
    }
    if (t->k <= 32) ms_resolve_kernel<1><<<grid_for(nq, 256), 256, 0, c->stream>>>(kp, t->lo, t->hi, t->exts, t->counts, q->lut.p, q->lut_shift, stranded,
                                                                                 reinterpret_cast<const QMsg<1>*>(d_queries), nq, d_reply);
    else ms_resolve_kernel<2><<<grid_for(nq, 256), 256, 0, c->stream>>>(kp, t->lo, t->hi, t->exts, t->counts, q->lut.p, q->lut_shift, stranded,
                                                                      reinterpret_cast<const QMsg<2>*>(d_queries), nq, d_reply);
    return check_launch(c, "ms_resolve");
}

int ms_apply_dev(Ctx* c, const Table* t, ShardCfg cfg, const MsQueries* q, const uint2* d_reply, uint4* d_rec) {
    if (!q->n_total) return DBG_OK;
    KP kp = make_kp(t->k);
    SegOff seg;
    for (int r = 0; r <= DBG_MAX_RANKS; r++) seg.off[r] = r <= cfg.P ? q->off[r] : q->off[cfg.P];
    u32* err = (u32*)(q->ctr.p + 1 + DBG_MAX_RANKS);
    if (t->k <= 32) ms_apply_kernel<1><<<grid_for(q->n_total, 256), 256, 0, c->stream>>>(kp, t->lo, t->hi, t->counts, cfg, reinterpret_cast<const QMsg<1>*>(q->msg.p),
                                                                                       q->src.p, d_reply, q->n_total, seg, d_rec, err);
    else ms_apply_kernel<2><<<grid_for(q->n_total, 256), 256, 0, c->stream>>>(kp, t->lo, t->hi, t->counts, cfg, reinterpret_cast<const QMsg<2>*>(q->msg.p),
                                                                            q->src.p, d_reply, q->n_total, seg, d_rec, err);
    return check_launch(c, "ms_apply");
}

// error flag of the link stage (after ms_apply_dev): 0 ok, 1 / 2 = inconsistent Exts (src/compression.rs:428-434)
int ms_link_error(Ctx* c, const MsQueries* q, u32* code) {
    u64 h = 0;
    TRY(read_u64(c, q->ctr.p + 1 + DBG_MAX_RANKS, &h));
    *code = (u32)h;
    return DBG_OK;
}

__global__ void __launch_bounds__(256) ms_count_ends_kernel(const uint4* __restrict__ rec, u64 n, u64* __restrict__ out) {
    __shared__ u32 s_w[8];
    u32 c = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const uint4 a = rec[i];
        c += (a.x == NIL || a.y == NIL) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        if (t) atomicAdd(out, t);
    }
}
// number of path ends (k-mers with a free side) in the shard: bounds the emit entries and the finished nodes of this rank
int ms_count_ends_dev(Ctx* c, const uint4* d_rec, u64 n, u64* d_out) {
    CU(c, cudaMemsetAsync(d_out, 0, 8, c->stream));
    if (!n) return DBG_OK;
    ms_count_ends_kernel<<<(u32)std::min<u64>(grid_for(n, 256), (u64)c->sm_count * 8), 256, 0, c->stream>>>(d_rec, n, d_out);
    return check_launch(c, "ms_count_ends");
}
u32 ms_fitem_bytes(int k) { return k <= 32 ? (u32)sizeof(FItem<1>) : (u32)sizeof(FItem<2>); }
int ms_fwalk_start_dev(Ctx* c, int k, const uint4* rec, const u64* klo, const u64* khi, int me, u64 n, u32 lmax, int reduce_op, const WalkOut& out) {
    if (!n) return DBG_OK;
    KP kp = make_kp(k);
    if (k <= 32) ms_fwalk_start_kernel<1><<<grid_for(n, 256), 256, 0, c->stream>>>(kp, rec, klo, khi, me, n, lmax, reduce_op, out);
    else ms_fwalk_start_kernel<2><<<grid_for(n, 256), 256, 0, c->stream>>>(kp, rec, klo, khi, me, n, lmax, reduce_op, out);
    return check_launch(c, "ms_fwalk_start");
}
int ms_fwalk_continue_dev(Ctx* c, int k, const uint4* rec, const u64* klo, const u64* khi, int me, const void* inbox, u64 n_in, u32 lmax, int reduce_op,
                          const WalkOut& out) {
    if (!n_in) return DBG_OK;
    KP kp = make_kp(k);
    if (k <= 32) ms_fwalk_continue_kernel<1><<<grid_for(n_in, 256), 256, 0, c->stream>>>(kp, rec, klo, khi, me, reinterpret_cast<const FItem<1>*>(inbox), n_in, lmax, reduce_op, out);
    else ms_fwalk_continue_kernel<2><<<grid_for(n_in, 256), 256, 0, c->stream>>>(kp, rec, klo, khi, me, reinterpret_cast<const FItem<2>*>(inbox), n_in, lmax, reduce_op, out);
    return check_launch(c, "ms_fwalk_continue");
}

// cuts: P + 1 bin indices over the top `bits` key bits; seg_off: P + 1 message offsets (prefix sums of the per-destination
// counts the caller derived from the local histogram and the cuts)
int ms_scatter_nodes_dev(Ctx* c, int k, const void* nmsg, u64 m, int P, int bits, const u64* cuts, const u64* seg_off, void* out) {
    if (!m) return DBG_OK;
    SegOff cu, sg;
    for (int r = 0; r <= DBG_MAX_RANKS; r++) { cu.off[r] = cuts[r <= P ? r : P]; sg.off[r] = seg_off[r <= P ? r : P]; }
    DBuf<u64> fill;
    TRY(fill.alloc_pool(c, DBG_MAX_RANKS));
    TRY(fill.zero());
    const int shift = 2 * k - bits;
    if (k <= 32) ms_scatter_nodes_kernel<1><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const NodeMsg<1>*>(nmsg), m, P, shift, cu, sg, fill.p, reinterpret_cast<NodeMsg<1>*>(out));
    else ms_scatter_nodes_kernel<2><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const NodeMsg<2>*>(nmsg), m, P, shift, cu, sg, fill.p, reinterpret_cast<NodeMsg<2>*>(out));
    return check_launch(c, "ms_scatter_nodes");
}
int ms_unpack_nodes_dev(Ctx* c, int k, const void* in, u64 m, u64* k_lo, u64* k_hi, u32* idx) {
    if (!m) return DBG_OK;
    if (k <= 32) ms_unpack_nodes_kernel<1><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const NodeMsg<1>*>(in), m, k_lo, k_hi, idx);
    else ms_unpack_nodes_kernel<2><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const NodeMsg<2>*>(in), m, k_lo, k_hi, idx);
    return check_launch(c, "ms_unpack_nodes");
}
int ms_node_len_dev(Ctx* c, int k, const void* msgs, const u32* idx, u64 m, u64* node_len, u32* out_length) {
    if (!m) return DBG_OK;
    if (k <= 32) ms_node_len_kernel<1><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const NodeMsg<1>*>(msgs), idx, m, k, node_len, out_length);
    else ms_node_len_kernel<2><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const NodeMsg<2>*>(msgs), idx, m, k, node_len, out_length);
    return check_launch(c, "ms_node_len");
}
int ms_emit_dev(Ctx* c, int k, const RecPeers& peers, const void* msgs, const u32* idx, const u64* node_start, u64 m, int reduce_op,
                u64* words, u8* out_exts, u16* out_data) {
    if (!m) return DBG_OK;
    KP kp = make_kp(k);
    if (k <= 32) ms_emit_kernel<1><<<grid_for(m, 128), 128, 0, c->stream>>>(kp, peers, reinterpret_cast<const NodeMsg<1>*>(msgs), idx, node_start, m, reduce_op, words, out_exts, out_data);
    else ms_emit_kernel<2><<<grid_for(m, 128), 128, 0, c->stream>>>(kp, peers, reinterpret_cast<const NodeMsg<2>*>(msgs), idx, node_start, m, reduce_op, words, out_exts, out_data);
    return check_launch(c, "ms_emit");
}

// histogram of the top `bits` bits of m ascending keys (seed-key splitters); d_hist zeroed here
int ms_key_hist_dev(Ctx* c, int k, const u64* k_lo, const u64* k_hi, u64 m, int bits, u32* d_hist) {
    CU(c, cudaMemsetAsync(d_hist, 0, sizeof(u32) << bits, c->stream));
    if (!m) return DBG_OK;
    const int shift = 2 * k - bits;
    if (k <= 32) lut_hist_kernel<1><<<grid_for(m, 256), 256, 0, c->stream>>>(k_lo, k_hi, m, shift, d_hist);
    else lut_hist_kernel<2><<<grid_for(m, 256), 256, 0, c->stream>>>(k_lo, k_hi, m, shift, d_hist);
    return check_launch(c, "ms_key_hist");
}

}  // namespace dbg
