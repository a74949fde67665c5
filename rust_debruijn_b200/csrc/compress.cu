// compression::compress_kmers_with_hash with SimpleCompress, rebuilt for B200
// (replaces CompressFromHash, src/compression.rs:355-594, and BaseGraph::add, src/graph.rs:104-113).
//
// The reference walks greedily from each still-available k-mer in hash-slot order, one dependent
// hash lookup after another, mutating an `available` bit set (src/compression.rs:450-479,574-580).
// That is strictly serial.  The device form is data parallel and produces the same BaseGraph as the
// greedy walk run with seed order = ascending k-mer (the order filter_kmers produces before
// BoomHashMap2::new permutes it):
//
//   S3 lookup        prefix LUT + binary search over the (already sorted) k-mer array (stands in for BoomHashMap2).
//   S4 links         per (k-mer, side): the stateless form of try_extend_kmer (:382-444): unique
//                    extension, neighbour present, neighbour != self, neighbour's extension back is
//                    unique, no palindromes (unstranded, even K).  Links are symmetric, so components
//                    are simple paths or simple cycles of k-mers.  Both links, the count, the Exts and the first /
//                    last base go into ONE 16-byte walk record per k-mer.
//   fast path        (every component is a path of <= 1024 k-mers)  discover: path ends walk to the other end, the
//                    walker that started at the node's left end appends one path record (seed = smallest index =
//                    the k-mer the greedy loop would reach first, :574-575); path records sorted by seed = node
//                    order; emit_walk: one thread per node re-walks its chain and writes whole words.
//   general path     (long unitigs, cycles)  S5 rank: end walks, then list ranking on a contracted graph of
//                    splitters (pointer doubling over port states s = 2i+d carrying window length, min index,
//                    distance, arrival port), cycles in a fixed-round second phase; S6 emit: node id / base offsets
//                    by exclusive scans over seeds, every k-mer ORs its base(s) into the bit-contiguous
//                    PackedDnaStringSet words (src/dna_string.rs:811-821, 383-399), end k-mers supply the node Exts
//                    (:513-517,534-540), counts are reduced per node (SimpleCompress::reduce, :58-60).
#include "common.cuh"
#include "lookup.cuh"

namespace dbg {

// ---- S4 -------------------------------------------------------------------------------------------
template <int W>
__global__ void links_kernel(KP kp, const u64* __restrict__ lo, const u64* __restrict__ hi, const u8* __restrict__ exts,
                             u64 v0, u64 n, const u64* __restrict__ lut, int lut_shift,
                             int stranded, u32* __restrict__ nxt, u32* __restrict__ err, uint4* __restrict__ rec16,
                             const u16* __restrict__ counts, int scmap) {
    // thread t handles k-mer i = v0 + t of the (full) table and writes nxt[2t + side]: v0 = 0, n = V on one GPU;
    // a rank of the sharded compression handles only its own index range.  With rec16 != nullptr the two links go
    // into ONE 16-byte record per k-mer together with everything a unitig walk needs from that k-mer (count, Exts,
    // first and last base), so that every later walk step is a single 16-byte load.
    u64 t_ = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= n) return;
    const u64 i = v0 + t_;
    Kmer<W> key = load_key<W>(lo, hi, i);
    u32 e = exts[i];
    bool pal = !stranded && is_palindrome<W>(kp, key);
    u32 both[2];
#pragma unroll
    for (int d = 0; d < 2; d++) {
        u32 succ = NIL;
        u32 nib = exts_side(e, d);
        if (popc4(nib) == 1 && !pal) {                                     // compression.rs:386
            u32 base = unique_base(nib);                                   // :390
            Kmer<W> nk = d == 0 ? Ops<W>::ext_left(kp, key, base) : Ops<W>::ext_right(kp, key, base);  // :392
            bool flip = false;
            if (!stranded) {                                               // :396-400
                Kmer<W> r = Ops<W>::rc(kp, nk);
                if (!(nk < r)) { nk = r; flip = true; }
            }
            u32 j = table_find<W>(lo, hi, lut, lut_shift, nk);             // :410
            if (j != NIL && j != (u32)i) {                                 // not in table / (self => already used) :411-415
                bool npal = !stranded && is_palindrome<W>(kp, nk);         // :403
                int inc = (d ^ 1) ^ (flip ? 1 : 0);                        // :419
                u32 ne = exts[j];
                u32 nnib = exts_side(ne, inc);
                int cnt = popc4(nnib);                                     // :422
                if (cnt == 0 && !npal) atomicExch(err, 1u);                // :428-434 panic!("unreachable")
                // CompressionSpec::join_test (:425): always true for SimpleCompress, data equality for ScmapCompress (:92-97)
                const bool can_join = !scmap || counts[i] == counts[j];
                if (can_join && cnt == 1 && !npal) {                       // :435
                    // reciprocity (always true for tables built by filter_kmers)
                    Kmer<W> back = inc == 0 ? Ops<W>::ext_left(kp, nk, unique_base(nnib)) : Ops<W>::ext_right(kp, nk, unique_base(nnib));
                    if (!stranded) { Kmer<W> r = Ops<W>::rc(kp, back); if (!(back < r)) back = r; }
                    if (back == key) succ = 2u * j + (u32)(inc ^ 1);
                    else atomicExch(err, 2u);
                }
            }
        }
        both[d] = succ;
    }
    if (rec16) {
        u32 meta = (u32)counts[i] | (e << 16) | (Ops<W>::first_base(kp, key) << 24) | (Ops<W>::last_base(kp, key) << 26);
        rec16[t_] = make_uint4(both[0], both[1], meta, 0u);
    } else {
        nxt[2 * t_] = both[0];
        nxt[2 * t_ + 1] = both[1];
    }
}

// successor of port state s.  sh = 1: compact array nxt[2v + side]; sh = 2: inside the 16-byte records (u32 view)
__device__ __forceinline__ u32 nxt_at(const u32* __restrict__ nxt, int sh, u32 s) { return nxt[((u64)(s >> 1) << sh) | (s & 1u)]; }

// ---- S5a: short unitigs by walking from their ends ------------------------------------------------------
// Every k-mer with exactly one free side is a path end.  Its thread walks to the other end (<= lmax steps),
// tracking the smallest index (the seed, compression.rs:574-575) and the port state in which the seed is
// traversed; the walker that started at the smaller-index end then walks again and writes
// (seed, position, length, orientation) for every k-mer of the unitig.  Paths longer than lmax and cycles
// are left untouched (nlen stays 0) for the pointer-doubling fallback.
// vinfo[v] = (seed, position in unitig, unitig length in k-mers [0 = not ranked yet], flags): one 16-byte store
__global__ void walk_kernel(const u32* __restrict__ nxt, int sh, u64 n, u32 lmax, uint4* __restrict__ vinfo, u64* __restrict__ written) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 wrote = 0;
    if (v < n) {
        u32 a0 = nxt_at(nxt, sh, 2u * (u32)v), a1 = nxt_at(nxt, sh, 2u * (u32)v + 1u);
        if (a0 == NIL && a1 == NIL) {
            vinfo[v] = make_uint4((u32)v, 0u, 1u, 1u);
            wrote = 1;
        } else if (a0 == NIL || a1 == NIL) {
            const u32 d = a0 == NIL ? 1u : 0u;  // the linked side: walk inwards through it
            u32 cur = 2u * (u32)v + d, cnt = 1, minv = (u32)v, minst = cur;
            u32 t = d ? a1 : a0;
            while (t != NIL && cnt <= lmax) {
                cur = t;
                cnt++;
                if ((t >> 1) < minv) { minv = t >> 1; minst = t; }
                t = nxt_at(nxt, sh, cur);
            }
            if (t == NIL && (u32)v < (cur >> 1)) {  // complete, and this is the walker from the smaller-index end
                const bool right = minst & 1u;     // the walk leaves the seed through R: it runs left -> right
                cur = 2u * (u32)v + d;
                for (u32 i = 0; i < cnt; i++) {
                    u32 w = cur >> 1, dw = cur & 1u;
                    u32 lp = right ? dw ^ 1u : dw;
                    vinfo[w] = make_uint4(minv, right ? i : cnt - 1 - i, cnt, (lp == 0) | (lp << 1));
                    cur = nxt_at(nxt, sh, cur);
                }
                wrote = cnt;
            }
        }
    }
    for (int o = 16; o; o >>= 1) wrote += __shfl_xor_sync(0xffffffffu, wrote, o);
    if ((threadIdx.x & 31) == 0 && wrote) atomicAdd(written, (u64)wrote);
}

// ---- S5b: long unitigs and cycles by list ranking on a CONTRACTED graph ------------------------------------
// Among the k-mers the walks left unranked, every path end and ~1/density of the others become SPLITTERS.
// Each (splitter, side) walks to the next splitter (segment_walk_kernel), which contracts the chains to a
// graph of splitters whose edges carry (length, smallest index inside, distance to it).  Pointer doubling then
// runs on the 2 x #splitters reduced port states only; the splitters get (seed, position, orientation) from it
// and hand them to the k-mers inside their segments with one more walk.  Work is O(V), depth ~density + log.
// rec = (ptr, len, minst, mindist): window of `len` consecutive k-mers starting at reduced state s; minst = the
// GLOBAL port state in which the smallest-index k-mer of the window is traversed; mindist = k-mers from s to it.
__global__ void mark_splitters_kernel(const u32* __restrict__ nxt, int sh, const uint4* __restrict__ vinfo, u64 n, u32 density_mask,
                                      u32* __restrict__ is_spl) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    u32 f = 0;
    if (vinfo[v].z == 0) {  // not ranked by the walks
        bool end = nxt_at(nxt, sh, 2u * (u32)v) == NIL || nxt_at(nxt, sh, 2u * (u32)v + 1u) == NIL;
        f = end || (((u32)v * 0x9E3779B1u >> 8) & density_mask) == 0;
    }
    is_spl[v] = f;
}
__global__ void fill_splitters_kernel(const u32* __restrict__ is_spl, const u64* __restrict__ sid, u64 n, u32* __restrict__ spl_vertex) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n && is_spl[v]) spl_vertex[sid[v]] = (u32)v;
}
// thread per reduced state (splitter, side): contract the chain up to the next splitter
__global__ void segment_walk_kernel(const u32* __restrict__ nxt, int sh, const u32* __restrict__ is_spl, const u64* __restrict__ sid,
                                    const u32* __restrict__ spl_vertex, u64 n_red, u32 cap, uint4* __restrict__ rec0,
                                    u32* __restrict__ err) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_red) return;
    u32 v = spl_vertex[r >> 1], d = (u32)r & 1u;
    u32 st = 2u * v + d;
    u32 cur = nxt_at(nxt, sh, st);
    u32 len = 1, minst = st, mind = 0, steps = 0;
    u32 ptr = NIL;
    while (cur != NIL) {
        u32 w = cur >> 1;
        if (is_spl[w]) { ptr = 2u * (u32)sid[w] + (cur & 1u); break; }
        if (w < (minst >> 1)) { minst = cur; mind = len; }
        len++;
        cur = nxt_at(nxt, sh, cur);
        if (++steps > cap) { atomicExch(err, 3u); break; }
    }
    rec0[r] = make_uint4(ptr, len, minst, mind);
}

__global__ void count_ranked_kernel(const uint4* __restrict__ vinfo, u64 n, u64* ranked, u64* unused) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 m = __ballot_sync(0xffffffffu, v < n && vinfo[v].z != 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(ranked, (u64)__popc(m));
}

__global__ void derive_seed_kernel(const uint4* __restrict__ vinfo, u64 n, int K, u32* __restrict__ is_seed,
                                   u64* __restrict__ node_len) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    uint4 vi = vinfo[v];
    bool issd = vi.x == (u32)v;
    is_seed[v] = issd;
    node_len[v] = issd ? (u64)vi.z + K - 1 : 0;
}

__device__ __forceinline__ uint4 pd_combine(uint4 a, uint4 t) {
    // window(a) ++ window(t); ties keep the first occurrence
    if ((t.z >> 1) < (a.z >> 1)) { a.z = t.z; a.w = a.y + t.w; }
    a.y += t.y;
    a.x = t.x >= NIL2 ? NIL : t.x;
    return a;
}

__global__ void pd_round_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, u64 n_states, u64* active) {
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool act = false;
    if (s < n_states) {
        uint4 a = src[s];
        if (a.x == NIL) {          // finished last round: mirror once into the other buffer
            a.x = NIL2;
            dst[s] = a;
        } else if (a.x != NIL2) {
            uint4 t = src[a.x];
            a = pd_combine(a, t);
            dst[s] = a;
            act = a.x != NIL;
        }
    }
    u32 m = __ballot_sync(0xffffffffu, act);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(active, (u64)__popc(m));
}

// cycles: collect still-active reduced states, restart them from their contracted segments, fixed rounds
__global__ void cyc_collect_kernel(const uint4* __restrict__ cur, u64 n_states, const uint4* __restrict__ rec0,
                                   u32* __restrict__ list, u64* cursor, uint4* __restrict__ a, uint4* __restrict__ b,
                                   u8* __restrict__ is_cyc) {
    u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    u32 p = cur[s].x;
    if (p < NIL2) {
        u64 pos = atomicAdd(cursor, 1ull);
        list[pos] = (u32)s;
        uint4 r = rec0[s];
        a[s] = r;
        b[s] = r;
        is_cyc[s >> 1] = 1;
    }
}
__global__ void cyc_round_kernel(const u32* __restrict__ list, u64 n, const uint4* __restrict__ src, uint4* __restrict__ dst) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s = list[i];
    uint4 a = src[s];
    uint4 t = src[a.x];
    if ((t.z >> 1) < (a.z >> 1)) { a.z = t.z; a.w = a.y + t.w; }
    a.y += t.y;
    a.x = t.x;
    dst[s] = a;
}

// ---- S5 result -> per k-mer (seed, position, orientation) ------------------------------------------------
// flags: bit0 fwd (k-mer appears in stored orientation), bit1 left_port (side of the k-mer facing the
// node's left end: 0 = L, 1 = R).  Thread per splitter.
__global__ void assign_splitters_kernel(const uint4* __restrict__ rec, const uint4* __restrict__ rec0,
                                        const u8* __restrict__ is_cyc, const u32* __restrict__ spl_vertex, u64 n_spl,
                                        uint4* __restrict__ vinfo, u64* __restrict__ n_cycle_kmers) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spl) return;
    const u32 v = spl_vertex[i];
    uint4 rl = rec[2 * i], rr = rec[2 * i + 1];
    u32 sd, ps, nn, fw, lp;
    if (!is_cyc[i]) {
        u32 ml = rl.z >> 1, mr = rr.z >> 1;
        nn = rl.y + rr.y - 1;
        if (ml == v && mr == v) {  // v is the seed: stored orientation, extended left then right
            sd = v; fw = 1; lp = 0; ps = rl.y - 1;
        } else {
            int d = ml < mr ? 0 : 1;
            u32 arr = d == 0 ? rl.z : rr.z;     // state in which the chain from v traverses the seed
            sd = arr >> 1;
            lp = (arr & 1u) == 0 ? (u32)d : (u32)(d ^ 1);
            fw = lp == 0;
            ps = (lp == 0 ? rl.y : rr.y) - 1;   // k-mers to the left of v
        }
    } else {
        // cycle: node = [step n-1 .. step 1, seed] of the chain leaving the seed through L (compression.rs:497-511)
        atomicAdd(n_cycle_kmers, (u64)rec0[2 * i].y);  // every cycle k-mer lies in exactly one L-going segment
        sd = rl.z >> 1;
        if (sd == v) {
            uint4 e = rec0[2 * i];               // first contracted segment leaving the seed through L
            nn = e.y + rec[e.x].w;               // ... plus the way from its far splitter back to the seed
        } else {
            nn = rl.w + rr.w;                    // the two ways round from v to the seed
        }
        int p = (rl.z & 1u) ? 0 : 1;             // direction from v whose chain enters the seed through L (leaves through R)
        u32 t = p == 0 ? rl.w : rr.w;
        ps = nn - 1 - t;
        lp = (u32)(p ^ 1);
        fw = lp == 0;
    }
    vinfo[v] = make_uint4(sd, ps, nn, fw | (lp << 1));
}
// thread per splitter: walk rightwards (node coordinates) to the next splitter and rank the k-mers in between
__global__ void segment_assign_kernel(const u32* __restrict__ nxt, int sh, const u32* __restrict__ is_spl,
                                      const u32* __restrict__ spl_vertex, u64 n_spl, u32 cap, uint4* __restrict__ vinfo) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spl) return;
    const u32 v = spl_vertex[i];
    const uint4 me = vinfo[v];
    const u32 d = ((me.w >> 1) & 1u) ^ 1u;  // the side facing right
    u32 cur = nxt_at(nxt, sh, 2u * v + d);
    u32 ps = me.y, steps = 0;
    while (cur != NIL) {
        u32 w = cur >> 1;
        if (is_spl[w]) break;
        ps = ps + 1 == me.z ? 0 : ps + 1;        // wraps on cycles (the seed is the LAST k-mer of its node)
        u32 lp = (cur & 1u) ^ 1u;                // leaving through (cur & 1) = right-facing side
        vinfo[w] = make_uint4(me.x, ps, me.z, (lp == 0) | (lp << 1));
        cur = nxt_at(nxt, sh, cur);
        if (++steps > cap) break;
    }
}

// ---- S6 ---------------------------------------------------------------------------------------------------
struct EmitArgs {
    const u64* lo; const u64* hi; const u8* exts; const u16* counts; u64 n;
    const uint4* vinfo;
    const u64* node_id; const u64* node_start;
    u64* words; u64* out_start; u32* out_length; u32* out_exts_w; u64* acc;
    int reduce_op;
};

template <int W>
__global__ void emit_kernel(KP kp, EmitArgs a) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = v < a.n;
    u64 nid = ~0ull - (threadIdx.x & 31);
    u64 cnt = 0;
    if (act) {
        const int K = kp.k;
        const uint4 vi = a.vinfo[v];
        u32 sd = vi.x, ps = vi.y, nn = vi.z, fl = vi.w;
        bool fw = fl & 1;
        int lp = (fl >> 1) & 1;
        nid = a.node_id[sd];
        u64 st = a.node_start[sd];
        Kmer<W> key = load_key<W>(a.lo, a.hi, v);
        if (!fw) key = Ops<W>::rc(kp, key);
        if (ps == 0) {
            // first k-mer of the node: all K bases, left-aligned then shifted to the node's bit offset
            int off = (int)(st & 31) * 2;
            u64 w = st >> 5;
            if constexpr (W == 1) {
                u64 X = key.lo << (64 - 2 * K);
                atomicOr(&a.words[w], X >> off);
                if (off && off + 2 * K > 64) atomicOr(&a.words[w + 1], X << (64 - off));
            } else {
                int sh = 128 - 2 * K;
                u64 H = sh ? (key.hi << sh) | (key.lo >> (64 - sh)) : key.hi;
                u64 L = key.lo << sh;
                atomicOr(&a.words[w], H >> off);
                u64 m = off ? (H << (64 - off)) | (L >> off) : L;
                if (m) atomicOr(&a.words[w + 1], m);
                if (off) { u64 t = L << (64 - off); if (t) atomicOr(&a.words[w + 2], t); }
            }
        } else {
            u64 g = st + ps + K - 1;
            u64 b = Ops<W>::last_base(kp, key);
            if (b) atomicOr(&a.words[g >> 5], b << (62 - 2 * (g & 31)));
        }
        if (sd == (u32)v) { a.out_start[nid] = st; a.out_length[nid] = (u32)((u64)nn + K - 1); }
        u32 e = a.exts[v];
        u32 eb = 0;
        if (ps == 0) { u32 nib = exts_side(e, lp); if (!fw) nib = exts_complement(nib) & 0xfu; eb |= nib; }            // :513-517
        if (ps == nn - 1) { u32 nib = exts_side(e, lp ^ 1); if (!fw) nib = exts_complement(nib) & 0xfu; eb |= nib << 4; }  // :534-540
        if (eb) atomicOr(&a.out_exts_w[nid >> 2], eb << (8 * (nid & 3)));
        cnt = a.counts[v];
    }
    // per-node data reduction; whole-warp-same-node fast path keeps giant unitigs off a single hot address
    u32 peers = __match_any_sync(0xffffffffu, nid);
    if (peers == 0xffffffffu) {
        if (a.reduce_op >= DBG_REDUCE_MAX) { for (int o = 16; o; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o)); }
        else { for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
        if ((threadIdx.x & 31) == 0 && act) {
            if (a.reduce_op >= DBG_REDUCE_MAX) atomicMax(&a.acc[nid], cnt); else atomicAdd(&a.acc[nid], cnt);
        }
    } else if (act) {
        if (a.reduce_op >= DBG_REDUCE_MAX) atomicMax(&a.acc[nid], cnt); else atomicAdd(&a.acc[nid], cnt);
    }
}

__global__ void finalize_nodes_kernel(const u64* __restrict__ acc, const u32* __restrict__ out_length, int K, u64 m,
                                      int reduce_op, u16* __restrict__ data) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u64 s = acc[i];
    bool single = out_length[i] == (u32)K;  // one k-mer: reduce() never called (compression.rs:495)
    u16 d;
    switch (reduce_op) {
        case DBG_REDUCE_SAT_ADD: d = (u16)(s > 65535 ? 65535 : s); break;
        case DBG_REDUCE_WRAP_ADD: d = (u16)(s & 0xffff); break;
        case DBG_REDUCE_ADD_MOD_65535: d = single ? (u16)s : (u16)(s % 65535); break;
        default: d = (u16)s; break;
    }
    data[i] = d;
}

__global__ void bytes_from_words_kernel(const u32* __restrict__ w, u8* __restrict__ out, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u8)(w[i >> 2] >> (8 * (i & 3)));
}


// ================================================================================================
// Fast path: every component is a path of <= lmax k-mers (sequencing data with errors: all of them).
//   discover   thread per k-mer; a path END walks the chain of 16-byte records to the other end, tracking
//              the smallest index (= the seed, compression.rs:574-575) and the port state in which the seed
//              is traversed.  Exactly one of the two end walkers traverses the seed "leaving through R":
//              that walk runs left -> right in node coordinates, so its start IS the node's left end; it
//              appends one path record (seed, length, left-end state).  Nothing is written per k-mer.
//   sort       path records by seed (node order = ascending seed, compression.rs:574-580); M << V records.
//   emit_walk  thread per NODE in output order: re-walks its chain from the left end, assembles the 2-bit
//              bases in registers and writes the node's words / Exts / data / start / length — coalesced
//              per-node stores instead of per-k-mer atomics into random words.
// One random 16-byte load per k-mer per walk replaces the ~8 random sectors per k-mer of the per-k-mer
// rank + emit formulation (which stays as the general path for long unitigs and cycles).
// ================================================================================================
__global__ void __launch_bounds__(256) discover_kernel(const uint4* __restrict__ rec, u64 v0, u64 n, u32 lmax, int key_shift,
                                                        u64* __restrict__ pkey, u32* __restrict__ pval, u64 cap,
                                                        u64* __restrict__ counters /* [0] paths, [1] k-mers covered */) {
    // thread t handles k-mer v0 + t (v0 = 0, n = V on one GPU; a rank of the sharded compression handles its index range)
    const u64 t_ = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 v = v0 + t_;
    const int lane = threadIdx.x & 31;
    bool emit = false;
    u32 seed = 0, len = 0, left_state = 0;
    if (t_ < n) {
        const uint4 r0 = rec[v];
        if (r0.x == NIL && r0.y == NIL) {
            emit = true; seed = (u32)v; len = 1; left_state = 2u * (u32)v + 1u;   // stored orientation: heading right = leaving through R
        } else if (r0.x == NIL || r0.y == NIL) {
            const u32 d = r0.x == NIL ? 1u : 0u;   // the linked side: walk inwards through it
            u32 cur = 2u * (u32)v + d, cnt = 1, minv = (u32)v, minst = cur;
            u32 t = d ? r0.y : r0.x;
            while (t != NIL && cnt <= lmax) {
                cur = t;
                cnt++;
                if ((t >> 1) < minv) { minv = t >> 1; minst = t; }
                const uint4 r = rec[t >> 1];
                t = (t & 1u) ? r.y : r.x;
            }
            if (t == NIL && (minst & 1u)) { emit = true; seed = minv; len = cnt; left_state = 2u * (u32)v + d; }
        }
    }
    // one reservation per CTA (same-address L2 atomics serialise: per-warp reservations would cost more than the walks)
    __shared__ u32 s_wcnt[8], s_wcov[8];
    __shared__ u64 s_base;
    const int warp = threadIdx.x >> 5;
    const u32 m = __ballot_sync(0xffffffffu, emit);
    u32 cov = emit ? len : 0;
    for (int o = 16; o; o >>= 1) cov += __shfl_xor_sync(0xffffffffu, cov, o);
    if (lane == 0) { s_wcnt[warp] = __popc(m); s_wcov[warp] = cov; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tc = 0, tv = 0;
        for (int w = 0; w < 8; w++) { u32 x = s_wcnt[w]; s_wcnt[w] = tc; tc += x; tv += s_wcov[w]; }
        s_base = tc ? atomicAdd(&counters[0], (u64)tc) : 0;
        if (tv) atomicAdd(&counters[1], (u64)tv);
    }
    __syncthreads();
    if (emit) {
        const u64 pos = s_base + s_wcnt[warp] + __popc(m & ((1u << lane) - 1));
        if (pos < cap) {
            pkey[pos] = ((u64)seed << key_shift) | len;   // seed in the top bits: the sort looks at those only
            pval[pos] = left_state;
        }
    }
}

__global__ void path_len_kernel(const u64* __restrict__ pkey, u64 m, int key_shift, int K, u64* __restrict__ node_len,
                                u32* __restrict__ out_length) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u64 l = (pkey[i] & ((1ull << key_shift) - 1)) + K - 1;
    node_len[i] = l;
    out_length[i] = (u32)l;
}

// Bit-contiguous writer for one node (PackedDnaStringSet::add, dna_string.rs:811-821): words strictly inside the
// node are plain stores, the first / last word may be shared with the neighbouring nodes -> atomicOr.
struct NodeWriter {
    u64* words; u64 first_w, last_w, wi; u64 cur;
    __device__ __forceinline__ void flush() {
        if (wi == first_w || wi == last_w) { if (cur) atomicOr(&words[wi], cur); }
        else words[wi] = cur;
        cur = 0; wi++;
    }
    // append n (1..32) bases given left-aligned in x (bits below the n bases must be zero) at base position pos
    __device__ __forceinline__ void push(u64 x, int n, u64 pos) {
        const int off = (int)(pos & 31);
        cur |= x >> (2 * off);
        if (off + n >= 32) {
            flush();
            if (off) cur = x << (64 - 2 * off);
        }
    }
};

struct EmitWalkArgs {
    const u64* lo; const u64* hi; const uint4* rec;
    const u64* pkey; const u32* pval; const u64* node_start; u64 i0, n_nodes; int key_shift;   // path records [i0, i0 + n_nodes)
    u64 out0, base0;   // node id = record index + out0; base offset = node_start[record] + base0 (sharded emission)
    u64* words; u8* out_exts; u16* out_data; int reduce_op;
    u64* out_start; u32* out_length;   // optional: global start / length of the emitted nodes (sharded emission)
};

template <int W>
__global__ void emit_walk_kernel(KP kp, EmitWalkArgs a) {
    const u64 t_ = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= a.n_nodes) return;
    const u64 i = a.i0 + t_;
    const int K = kp.k;
    const u32 len = (u32)(a.pkey[i] & ((1ull << a.key_shift) - 1));
    u32 cur = a.pval[i];             // at the left end, leaving through its right-facing side
    const u64 st = a.node_start[i] + a.base0;
    const u64 L = (u64)len + K - 1;
    NodeWriter nw;
    nw.words = a.words; nw.first_w = st >> 5; nw.last_w = (st + L - 1) >> 5; nw.wi = nw.first_w; nw.cur = 0;
    u64 pos = st;
    u64 acc = 0;
    u32 eb = 0;
    {   // first k-mer: all K bases (compression.rs:489-495)
        const u32 w = cur >> 1, dw = cur & 1u;
        const bool fw = dw == 1u;    // leaving through R while heading right = stored orientation
        Kmer<W> key = load_key<W>(a.lo, a.hi, w);
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
        const uint4 r = a.rec[w];
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
    }
    for (u32 j = 1; j < len; j++) {
        const u32 w = cur >> 1, dw = cur & 1u;
        const bool fw = dw == 1u;
        const uint4 r = a.rec[w];
        const u32 fb = (r.z >> 24) & 3u, lb = (r.z >> 26) & 3u;
        const u64 b = fw ? lb : 3u - fb;                 // last base of the k-mer as it appears in the node
        nw.push(b << 62, 1, pos);
        pos++;
        const u64 cnt = r.z & 0xffffu;
        if (a.reduce_op >= DBG_REDUCE_MAX) acc = cnt > acc ? cnt : acc; else acc += cnt;   // SCMAP: all equal, max == the value
        if (j == len - 1) {
            u32 rn = exts_side((r.z >> 16) & 0xffu, (int)dw);   // right-facing side of the last k-mer (:534-540)
            if (!fw) rn = exts_complement(rn) & 0xfu;
            eb |= rn << 4;
        }
        cur = dw ? r.y : r.x;
    }
    if (pos & 31) nw.flush();   // partial last word
    a.out_exts[i + a.out0] = (u8)eb;
    if (a.out_start) { a.out_start[i + a.out0] = st; a.out_length[i + a.out0] = (u32)L; }
    u16 d;
    switch (a.reduce_op) {
        case DBG_REDUCE_SAT_ADD: d = (u16)(acc > 65535 ? 65535 : acc); break;
        case DBG_REDUCE_WRAP_ADD: d = (u16)(acc & 0xffff); break;
        case DBG_REDUCE_ADD_MOD_65535: d = len == 1 ? (u16)acc : (u16)(acc % 65535); break;   // one k-mer: reduce() never called (:495)
        default: d = (u16)acc; break;
    }
    a.out_data[i + a.out0] = d;
}

template <int W>
static int compress_impl(Ctx* c, const Table* t, int stranded, int reduce_op, Graph** out) {
    cudaStream_t st = c->stream;
    dbg_stats& S = c->stats;
    KP kp = make_kp(t->k);
    const u64 V = t->n;
    Graph* g = &(new dbg_graph())->g;
    g->ctx = c; g->k = t->k; g->stranded = stranded;
    *out = g;
    S.n_nodes = S.n_bases = 0; S.rank_rounds = 0; S.n_cycle_kmers = 0;
    if (V == 0) return DBG_OK;
    TRY(arena_begin(c));
    if (V >= (1ull << 31)) DBG_SET_ERR(c, DBG_E_BADARG, "k-mer table too large for 32-bit port states (%llu)", (unsigned long long)V);
    CU(c, cudaEventRecord(c->ev[0], st));
    // ---- S3 ----
    int lb = 8;
    while ((1ull << (lb + 4)) <= V && lb < 24) lb++;   // ~8-16 keys per prefix
    if (lb > 2 * t->k) lb = 2 * t->k;
    const int lut_shift = 2 * t->k - lb;
    const u64 n_pfx = 1ull << lb;
    DBuf<u32> lut_cnt;
    DBuf<u64> lut;
    TRY(lut_cnt.alloc(c, n_pfx));
    TRY(lut.alloc(c, n_pfx + 1));
    TRY(lut_cnt.zero());
    lut_hist_kernel<W><<<grid_for(V, 256), 256, 0, st>>>(t->lo, t->hi, V, lut_shift, lut_cnt.p);
    TRY(check_launch(c, "lut_hist"));
    TRY(exclusive_scan_u32_to_u64(c, lut_cnt.p, lut.p, n_pfx, lut.p + n_pfx));
    CU(c, cudaEventRecord(c->ev[1], st));
    // ---- S4: links + per-k-mer walk record (16 bytes: both links, count, Exts, first / last base) ----
    DBuf<uint4> rec16;
    DBuf<u64> ctr;
    TRY(rec16.alloc(c, V));
    TRY(ctr.alloc(c, 4));
    TRY(ctr.zero());
    links_kernel<W><<<grid_for(V, 256), 256, 0, st>>>(kp, t->lo, t->hi, t->exts, 0, V, lut.p, lut_shift, stranded, nullptr,
                                                      (u32*)(ctr.p + 3), rec16.p, t->counts, reduce_op == DBG_REDUCE_SCMAP);
    TRY(check_launch(c, "links"));
    CU(c, cudaEventRecord(c->ev[2], st));
    const u32* nxt_p = reinterpret_cast<const u32*>(rec16.p);   // general path: links read in place (stride 4 words)
    const int nsh = 2;
    // ---- fast path: all components are short paths -> path records, sorted by seed, one walker per node ----
    {
        int bits_v = 1;
        while ((1ull << bits_v) < V) bits_v++;
        const int key_shift = 64 - bits_v;
        const u32 lmax = 1024u;
        DBuf<u64> pk_a, pk_b;
        DBuf<u32> pv_a, pv_b;
        TRY(pk_a.alloc(c, V)); TRY(pv_a.alloc(c, V));
        discover_kernel<<<grid_for(V, 256), 256, 0, st>>>(rec16.p, 0, V, lmax, key_shift, pk_a.p, pv_a.p, V, ctr.p);
        TRY(check_launch(c, "discover"));
        u64 h[4];
        TRY(read_u64(c, ctr.p, h, 4));
        if (h[3] == 1) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "k-mer extension points at a k-mer with no extension back (src/compression.rs:428-434)");
        if (h[3] == 2) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "k-mer extensions are not reciprocal");
        if (h[1] == V && !c->no_fast_compress) {
            const u64 M = h[0], Lb = V + M * (u64)(t->k - 1);   // every k-mer adds one base, every node K-1 more
            TRY(pk_b.alloc(c, M)); TRY(pv_b.alloc(c, M));
            u64 *rk, *rh;
            u32* rv;
            TRY(radix_sort_pairs(c, 1, 64, M, pk_a.p, nullptr, pv_a.p, pk_b.p, nullptr, pv_b.p, &rk, &rh, &rv));
            g->n_nodes = M; g->n_bases = Lb; g->n_words = (Lb + 31) / 32;
            S.n_nodes = M; S.n_bases = Lb;
            DBuf<u64> words, ostart, node_len;
            DBuf<u32> olen;
            DBuf<u8> oexts;
            DBuf<u16> odata;
            TRY(words.alloc_pool(c, g->n_words + 3)); TRY(words.zero());
            TRY(ostart.alloc_pool(c, M)); TRY(olen.alloc_pool(c, M)); TRY(oexts.alloc_pool(c, M)); TRY(odata.alloc_pool(c, M));
            TRY(node_len.alloc(c, M));
            path_len_kernel<<<grid_for(M, 256), 256, 0, st>>>(rk, M, key_shift, t->k, node_len.p, olen.p);
            TRY(check_launch(c, "path_len"));
            TRY(exclusive_scan_u64(c, node_len.p, ostart.p, M, nullptr));
            CU(c, cudaEventRecord(c->ev[3], st));
            EmitWalkArgs ea;
            ea.lo = t->lo; ea.hi = t->hi; ea.rec = rec16.p; ea.pkey = rk; ea.pval = rv; ea.node_start = ostart.p; ea.i0 = 0; ea.n_nodes = M; ea.out0 = 0; ea.base0 = 0; ea.out_start = nullptr; ea.out_length = nullptr;
            ea.key_shift = key_shift; ea.words = words.p; ea.out_exts = oexts.p; ea.out_data = odata.p; ea.reduce_op = reduce_op;
            emit_walk_kernel<W><<<grid_for(M, 128), 128, 0, st>>>(kp, ea);
            TRY(check_launch(c, "emit_walk"));
            CU(c, cudaEventRecord(c->ev[4], st));
            TRY(sync(c));
            g->words = words.take(); g->start = ostart.take(); g->length = olen.take(); g->exts = oexts.take(); g->data = odata.take();
            cudaEventElapsedTime(&S.ms_table, c->ev[0], c->ev[1]);
            cudaEventElapsedTime(&S.ms_links, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&S.ms_rank, c->ev[2], c->ev[3]);
            cudaEventElapsedTime(&S.ms_emit, c->ev[3], c->ev[4]);
            cudaEventElapsedTime(&S.ms_compress_total, c->ev[0], c->ev[4]);
            S.gpu_launches = c->launches;
            return DBG_OK;
        }
        CU(c, cudaMemsetAsync(ctr.p, 0, 24, st));   // general path below reuses counters [0..2]
    }
    // ---- S5a: walks for short unitigs ----
    DBuf<uint4> vinfo;
    DBuf<u32> is_seed;
    DBuf<u64> node_len, node_id, tot;
    TRY(vinfo.alloc(c, V)); TRY(is_seed.alloc(c, V));
    TRY(node_len.alloc(c, V)); TRY(node_id.alloc(c, V)); TRY(tot.alloc(c, 2));
    TRY(vinfo.zero());
    walk_kernel<<<grid_for(V, 256), 256, 0, st>>>(nxt_p, nsh, V, 1024u, vinfo.p, ctr.p + 2);
    TRY(check_launch(c, "walk"));
    {
        u64 h[4];
        TRY(read_u64(c, ctr.p, h, 4));
        if (h[3] == 1) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "k-mer extension points at a k-mer with no extension back (src/compression.rs:428-434)");
        if (h[3] == 2) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "k-mer extensions are not reciprocal");
        S.rank_rounds = 0;
        u64 ranked = h[2];
        // ---- S5b: list ranking on the contracted graph for what the walks left (long unitigs, cycles).
        // First with ~1/64 of the k-mers as splitters; anything still unranked (a cycle that drew no splitter)
        // is redone with every k-mer a splitter, which is plain pointer doubling on the leftovers. ----
        for (int pass = 0; pass < 2 && ranked != V; pass++) {
            const u32 density_mask = pass == 0 ? 63u : 0u;
            const u32 cap = 1u << 20;
            DBuf<u32> is_spl, spl_vertex;
            DBuf<u64> sid, nsp;
            TRY(is_spl.alloc(c, V)); TRY(sid.alloc(c, V)); TRY(nsp.alloc(c, 1));
            mark_splitters_kernel<<<grid_for(V, 256), 256, 0, st>>>(nxt_p, nsh, vinfo.p, V, density_mask, is_spl.p);
            TRY(check_launch(c, "mark_splitters"));
            TRY(exclusive_scan_u32_to_u64(c, is_spl.p, sid.p, V, nsp.p));
            u64 NSPL = 0;
            TRY(read_u64(c, nsp.p, &NSPL));
            if (!NSPL) continue;  // e.g. only short cycles that drew no splitter: next pass makes every k-mer one
            const u64 NR = 2 * NSPL;
            TRY(spl_vertex.alloc(c, NSPL));
            fill_splitters_kernel<<<grid_for(V, 256), 256, 0, st>>>(is_spl.p, sid.p, V, spl_vertex.p);
            TRY(check_launch(c, "fill_splitters"));
            DBuf<uint4> rec0, recA, recB;
            TRY(rec0.alloc(c, NR)); TRY(recA.alloc(c, NR)); TRY(recB.alloc(c, NR));
            segment_walk_kernel<<<grid_for(NR, 256), 256, 0, st>>>(nxt_p, nsh, is_spl.p, sid.p, spl_vertex.p, NR, cap, rec0.p, (u32*)(ctr.p + 3));
            TRY(check_launch(c, "segment_walk"));
            CU(c, cudaMemcpyAsync(recA.p, rec0.p, NR * sizeof(uint4), cudaMemcpyDeviceToDevice, st));
            uint4 *src = recA.p, *dst = recB.p;
            u64 prev_active = ~0ull, active = 0;
            int rounds = 0;
            for (; rounds < 40; rounds++) {
                CU(c, cudaMemsetAsync(ctr.p, 0, 8, st));
                pd_round_kernel<<<grid_for(NR, 256), 256, 0, st>>>(src, dst, NR, ctr.p);
                TRY(check_launch(c, "pd_round"));
                std::swap(src, dst);
                TRY(read_u64(c, ctr.p, &active));
                if (active == 0 || active == prev_active) break;
                prev_active = active;
            }
            // one more pass so that states finished in the last round exist in both buffers
            CU(c, cudaMemsetAsync(ctr.p, 0, 8, st));
            pd_round_kernel<<<grid_for(NR, 256), 256, 0, st>>>(src, dst, NR, ctr.p);
            TRY(check_launch(c, "pd_round"));
            S.rank_rounds += rounds + 1;
            DBuf<u8> is_cyc;
            TRY(is_cyc.alloc(c, NSPL));
            TRY(is_cyc.zero());
            if (active) {
                DBuf<u32> list;
                TRY(list.alloc(c, active));
                CU(c, cudaMemsetAsync(ctr.p + 1, 0, 8, st));
                cyc_collect_kernel<<<grid_for(NR, 256), 256, 0, st>>>(src, NR, rec0.p, list.p, ctr.p + 1, src, dst, is_cyc.p);
                TRY(check_launch(c, "cyc_collect"));
                u64 ncs = 0;
                TRY(read_u64(c, ctr.p + 1, &ncs));
                int cr = 1;
                while ((1ull << cr) < ncs) cr++;
                cr += 1;
                for (int r = 0; r < cr; r++) {
                    cyc_round_kernel<<<grid_for(ncs, 256), 256, 0, st>>>(list.p, ncs, src, dst);
                    TRY(check_launch(c, "cyc_round"));
                    std::swap(src, dst);
                }
                S.rank_rounds += cr;
            }
            CU(c, cudaMemsetAsync(ctr.p + 1, 0, 8, st));
            assign_splitters_kernel<<<grid_for(NSPL, 256), 256, 0, st>>>(src, rec0.p, is_cyc.p, spl_vertex.p, NSPL, vinfo.p, ctr.p + 1);
            TRY(check_launch(c, "assign_splitters"));
            segment_assign_kernel<<<grid_for(NSPL, 256), 256, 0, st>>>(nxt_p, nsh, is_spl.p, spl_vertex.p, NSPL, cap, vinfo.p);
            TRY(check_launch(c, "segment_assign"));
            // how many k-mers are ranked now?
            CU(c, cudaMemsetAsync(ctr.p + 2, 0, 8, st));
            count_ranked_kernel<<<grid_for(V, 256), 256, 0, st>>>(vinfo.p, V, ctr.p + 2, nullptr);
            TRY(check_launch(c, "count_ranked"));
            u64 hh[4];
            TRY(read_u64(c, ctr.p, hh, 4));
            ranked = hh[2];
            S.n_cycle_kmers += hh[1];
            if (hh[3] == 3) DBG_SET_ERR(c, DBG_E_INTERNAL, "segment walk exceeded its step cap");
        }
        if (ranked != V) DBG_SET_ERR(c, DBG_E_INTERNAL, "ranking left %llu k-mers unassigned", (unsigned long long)(V - ranked));
    }
    CU(c, cudaEventRecord(c->ev[3], st));
    // ---- S6 ----
    derive_seed_kernel<<<grid_for(V, 256), 256, 0, st>>>(vinfo.p, V, t->k, is_seed.p, node_len.p);
    TRY(check_launch(c, "derive_seed"));
    TRY(exclusive_scan_u32_to_u64(c, is_seed.p, node_id.p, V, tot.p));
    TRY(exclusive_scan_u64(c, node_len.p, node_len.p, V, tot.p + 1));
    u64 h[2];
    TRY(read_u64(c, tot.p, h, 2));
    const u64 M = h[0], Lb = h[1];
    g->n_nodes = M; g->n_bases = Lb; g->n_words = (Lb + 31) / 32;
    S.n_nodes = M; S.n_bases = Lb;
    DBuf<u64> words, ostart, acc;
    DBuf<u32> olen, oextw;
    DBuf<u8> oexts;
    DBuf<u16> odata;
    TRY(words.alloc_pool(c, g->n_words + 3)); TRY(words.zero());
    TRY(ostart.alloc_pool(c, M)); TRY(olen.alloc_pool(c, M)); TRY(oextw.alloc(c, M / 4 + 1)); TRY(oextw.zero());
    TRY(acc.alloc(c, M)); TRY(acc.zero()); TRY(oexts.alloc_pool(c, M)); TRY(odata.alloc_pool(c, M));
    EmitArgs ea;
    ea.lo = t->lo; ea.hi = t->hi; ea.exts = t->exts; ea.counts = t->counts; ea.n = V;
    ea.vinfo = vinfo.p;
    ea.node_id = node_id.p; ea.node_start = node_len.p;
    ea.words = words.p; ea.out_start = ostart.p; ea.out_length = olen.p; ea.out_exts_w = oextw.p; ea.acc = acc.p;
    ea.reduce_op = reduce_op;
    emit_kernel<W><<<grid_for(V, 256), 256, 0, st>>>(kp, ea);
    TRY(check_launch(c, "emit"));
    finalize_nodes_kernel<<<grid_for(M, 256), 256, 0, st>>>(acc.p, olen.p, t->k, M, reduce_op, odata.p);
    TRY(check_launch(c, "finalize_nodes"));
    bytes_from_words_kernel<<<grid_for(M, 256), 256, 0, st>>>(oextw.p, oexts.p, M);
    TRY(check_launch(c, "bytes_from_words"));
    CU(c, cudaEventRecord(c->ev[4], st));
    TRY(sync(c));
    g->words = words.take(); g->start = ostart.take(); g->length = olen.take(); g->exts = oexts.take(); g->data = odata.take();
    cudaEventElapsedTime(&S.ms_table, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&S.ms_links, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&S.ms_rank, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&S.ms_emit, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&S.ms_compress_total, c->ev[0], c->ev[4]);
    S.gpu_launches = c->launches;
    return DBG_OK;
}

// ================================================================================================
// filter::remove_censored_exts / remove_censored_exts_sharded (src/filter.rs:238-306), SURVEY §8f N2: drop the
// extension bits of every valid k-mer that point at a k-mer which is not valid (plain variant), or which is not
// valid but was seen in this shard, i.e. is in all_kmers (sharded variant: extensions into other shards are
// kept).  One thread per k-mer, up to 8 (16) lookups through the same prefix LUT + binary search that stands in
// for the reference's binary_search_by_key.  Only keys are read, every thread rewrites its own Exts byte.
// ================================================================================================
template <int W>
__global__ void censor_exts_kernel(KP kp, const u64* __restrict__ lo, const u64* __restrict__ hi, u8* __restrict__ exts, u64 n,
                                   const u64* __restrict__ lut, int lut_shift, int stranded, const u64* __restrict__ all_lo,
                                   const u64* __restrict__ all_hi, const u64* __restrict__ all_lut, int all_shift, int sharded) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Kmer<W> key = load_key<W>(lo, hi, i);
    const u32 e = exts[i];
    u32 ne = 0;
#pragma unroll
    for (int d = 0; d < 2; d++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const u32 bit = 1u << (4 * d + b);
            if (!(e & bit)) continue;                                                    // Exts::has_ext
            Kmer<W> nk = d == 0 ? Ops<W>::ext_left(kp, key, b) : Ops<W>::ext_right(kp, key, b);   // kmer.extend(i, dir)
            if (!stranded) { Kmer<W> r = Ops<W>::rc(kp, nk); if (r < nk) nk = r; }        // min_rc
            bool keep = table_find<W>(lo, hi, lut, lut_shift, nk) != NIL;                 // valid: not censored
            if (!keep && sharded) keep = table_find<W>(all_lo, all_hi, all_lut, all_shift, nk) == NIL;   // not seen here: other shard
            if (keep) ne |= bit;
        }
    }
    exts[i] = (u8)ne;
}

template <int W>
static int censor_impl(Ctx* c, Table* t, int stranded, int sharded) {
    if (t->n == 0) return DBG_OK;
    if (t->n >= (1ull << 32) - 1 || t->n_all >= (1ull << 32) - 1) DBG_SET_ERR(c, DBG_E_BADARG, "table too large for 32-bit lookups");
    TRY(arena_begin(c));
    KP kp = make_kp(t->k);
    DBuf<u32> cnt, acnt;
    DBuf<u64> lut, alut;
    int shift = 0, ashift = 0;
    TRY(build_prefix_lut<W>(c, t->k, t->lo, t->hi, t->n, cnt, lut, &shift));
    if (sharded) TRY(build_prefix_lut<W>(c, t->k, t->all_lo, t->all_hi, t->n_all, acnt, alut, &ashift));
    censor_exts_kernel<W><<<grid_for(t->n, 256), 256, 0, c->stream>>>(kp, t->lo, t->hi, t->exts, t->n, lut.p, shift, stranded,
                                                                      t->all_lo, t->all_hi, alut.p, ashift, sharded);
    TRY(check_launch(c, "censor_exts"));
    return sync(c);
}

int remove_censored_exts_dev(Ctx* c, Table* t, int stranded, int sharded) {
    if (!t) DBG_SET_ERR(c, DBG_E_BADARG, "null table");
    if (sharded && t->n_all == 0 && t->n != 0)
        DBG_SET_ERR(c, DBG_E_BADARG, "remove_censored_exts_sharded needs all_kmers: run filter_kmers with report_all_kmers");
    return t->k <= 32 ? censor_impl<1>(c, t, stranded, sharded) : censor_impl<2>(c, t, stranded, sharded);
}

// histogram of the top `bits` bits of the (ascending) keys: 2^bits u32 counters, zeroed here
int table_prefix_hist_dev(Ctx* c, const Table* t, int bits, u32* d_hist) {
    if (!t || bits < 1 || bits > 24 || bits > 2 * t->k) DBG_SET_ERR(c, DBG_E_BADARG, "bad prefix width %d", bits);
    CU(c, cudaMemsetAsync(d_hist, 0, sizeof(u32) << bits, c->stream));
    if (!t->n) return DBG_OK;
    const int shift = 2 * t->k - bits;
    if (t->k <= 32) lut_hist_kernel<1><<<grid_for(t->n, 256), 256, 0, c->stream>>>(t->lo, t->hi, t->n, shift, d_hist);
    else lut_hist_kernel<2><<<grid_for(t->n, 256), 256, 0, c->stream>>>(t->lo, t->hi, t->n, shift, d_hist);
    return check_launch(c, "lut_hist");
}

// Adopt caller-owned device arrays (copied) as a BaseGraph handle.
int graph_from_device_dev(Ctx* c, int k, int stranded, u64 n_nodes, u64 n_bases, const u64* d_words, const u64* d_start,
                          const u32* d_length, const u8* d_exts, const u16* d_data, Graph** out) {
    *out = nullptr;
    cudaStream_t st = c->stream;
    Graph* g = &(new dbg_graph())->g;
    g->ctx = c; g->k = k; g->stranded = stranded; g->n_nodes = n_nodes; g->n_bases = n_bases; g->n_words = (n_bases + 31) / 32;
    *out = g;
    if (n_nodes == 0) return DBG_OK;
    DBuf<u64> words, ostart;
    DBuf<u32> olen;
    DBuf<u8> oexts;
    DBuf<u16> odata;
    int rc = words.alloc_pool(c, g->n_words + 3);
    if (rc == DBG_OK) rc = ostart.alloc_pool(c, n_nodes);
    if (rc == DBG_OK) rc = olen.alloc_pool(c, n_nodes);
    if (rc == DBG_OK) rc = oexts.alloc_pool(c, n_nodes);
    if (rc == DBG_OK) rc = odata.alloc_pool(c, n_nodes);
    if (rc != DBG_OK) { free_graph(g); *out = nullptr; return rc; }
    cudaMemcpyAsync(words.p, d_words, g->n_words * 8, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(ostart.p, d_start, n_nodes * 8, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(olen.p, d_length, n_nodes * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(odata.p, d_data, n_nodes * 2, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(oexts.p, d_exts, n_nodes, cudaMemcpyDeviceToDevice, st);
    if (spin_sync(st) != cudaSuccess) {
        free_graph(g); *out = nullptr;
        DBG_SET_ERR(c, DBG_E_CUDA, "graph_from_device: %s", cudaGetErrorString(cudaGetLastError()));
    }
    g->words = words.take(); g->start = ostart.take(); g->length = olen.take(); g->exts = oexts.take(); g->data = odata.take();
    c->stats.n_nodes = n_nodes; c->stats.n_bases = n_bases;
    return DBG_OK;
}

int compress_dev(Ctx* c, const Table* t, int stranded, int reduce_op, Graph** out) {
    *out = nullptr;
    if (!t) DBG_SET_ERR(c, DBG_E_BADARG, "null table");
    if (reduce_op < 0 || reduce_op > DBG_REDUCE_SCMAP) DBG_SET_ERR(c, DBG_E_BADARG, "unknown reduce_op %d", reduce_op);
    int rc = t->k <= 32 ? compress_impl<1>(c, t, stranded, reduce_op, out) : compress_impl<2>(c, t, stranded, reduce_op, out);
    if (rc != DBG_OK && *out) { free_graph(*out); *out = nullptr; }
    return rc;
}

}  // namespace dbg
