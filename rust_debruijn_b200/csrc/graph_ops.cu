// Node-level graph operations on a device-resident BaseGraph:
//   BaseGraph::finish + DebruijnGraph::find_edges / find_link        src/graph.rs:116-142, 223-291
//   DebruijnGraph::fix_exts / get_valid_exts                          src/graph.rs:337-377
//   DebruijnGraph::is_compressed                                      src/graph.rs:296-334
//   BaseGraph::combine                                                src/graph.rs:71-100
//   compression::compress_graph (CompressFromGraph)                   src/compression.rs:100-349
#include "common.cuh"
#include "lookup.cuh"

namespace dbg {

// ================================================================================================
// BaseGraph::finish + DebruijnGraph::find_edges / find_link for every (node, side) — src/graph.rs:116-142, 223-291
// (SURVEY §8f N1).  left_order / right_order (BoomHashMap: first / last k-mer of a node -> node id) become two
// sorted (k-mer, node) arrays searched through a prefix LUT; one thread per (node, side) tries its <= 4 extensions.
// Output slot (node * 2 + side) * 4 + base: target node (0xffffffff = no such extension / link not in this graph),
// flags bit 0 = incoming side (0 Left, 1 Right), bit 1 = rc flip.
// ================================================================================================
template <int W>
__device__ __forceinline__ Kmer<W> kmer_at(const KP& kp, const u64* __restrict__ words, u64 b) {   // Vmer::get_kmer
    const int K = kp.k;
    Kmer<W> r;
    if constexpr (W == 1) {
        r.lo = bases64(words, b) >> (64 - 2 * K);
    } else {
        const u64 H = bases64(words, b), L = bases64(words, b + 32);
        const int sh = 128 - 2 * K;   // 0..62
        r.hi = sh ? H >> sh : H;
        r.lo = sh ? (L >> sh) | (H << (64 - sh)) : L;
    }
    return r;
}

template <int W>
__global__ void node_term_kmers_kernel(KP kp, const u64* __restrict__ words, const u64* __restrict__ start, const u32* __restrict__ length,
                                       u64 m, u64* __restrict__ f_lo, u64* __restrict__ f_hi, u64* __restrict__ l_lo, u64* __restrict__ l_hi,
                                       u32* __restrict__ id_a, u32* __restrict__ id_b) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const Kmer<W> f = kmer_at<W>(kp, words, start[i]);                         // first_kmer, lib.rs:369-371
    const Kmer<W> l = kmer_at<W>(kp, words, start[i] + length[i] - kp.k);      // last_kmer, lib.rs:374-376
    f_lo[i] = f.lo; l_lo[i] = l.lo;
    if constexpr (W == 2) { f_hi[i] = f.hi; l_hi[i] = l.hi; }
    id_a[i] = (u32)i; id_b[i] = (u32)i;
}

template <int W>
struct EndMap { const u64* lo; const u64* hi; const u32* node; const u64* lut; int shift; };

template <int W>
__device__ __forceinline__ u32 endmap_find(const EndMap<W>& mp, Kmer<W> key) {
    const u32 j = table_find<W>(mp.lo, mp.hi, mp.lut, mp.shift, key);
    return j == NIL ? NIL : mp.node[j];
}

template <int W>
__global__ void graph_edges_kernel(KP kp, const u64* __restrict__ words, const u64* __restrict__ start, const u32* __restrict__ length,
                                   const u8* __restrict__ exts, u64 m, int stranded, EndMap<W> left, EndMap<W> right,
                                   u32* __restrict__ target, u8* __restrict__ flags) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * m) return;
    const u64 n = t >> 1;
    const int dir = (int)(t & 1);
    const Kmer<W> kmer = kmer_at<W>(kp, words, dir ? start[n] + length[n] - kp.k : start[n]);   // term_kmer
    const u32 e = exts[n];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 tg = NIL, fl = 0;
        if ((e >> (4 * dir + i)) & 1u) {                                            // find_edges, graph.rs:229-239
            const Kmer<W> x = dir ? Ops<W>::ext_right(kp, kmer, i) : Ops<W>::ext_left(kp, kmer, i);
            // find_link, graph.rs:252-291: same strand through the opposite side, else (unstranded) the rc through the same side
            tg = endmap_find<W>(dir ? left : right, x);
            fl = dir ? 0u : 1u;
            if (tg == NIL && !stranded) {
                tg = endmap_find<W>(dir ? right : left, Ops<W>::rc(kp, x));
                fl = (dir ? 1u : 0u) | 2u;
            }
            if (tg == NIL) fl = 0;
        }
        target[t * 4 + i] = tg;
        flags[t * 4 + i] = (u8)fl;
    }
}

// valid_nodes (optional, host, one byte per node): fix_exts' BitSet — links into nodes marked 0 do not count
__global__ void fix_exts_kernel(const u32* __restrict__ target, const u8* __restrict__ valid_nodes, u64 m, u8* __restrict__ exts) {
    const u64 n = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= m) return;
    u32 e = 0;
#pragma unroll
    for (int s = 0; s < 8; s++) {   // slot (n * 2 + dir) * 4 + i  <->  Exts bit 4 * dir + i
        const u32 tg = target[n * 8 + s];
        if (tg != NIL && (!valid_nodes || valid_nodes[tg])) e |= 1u << s;
    }
    exts[n] = (u8)e;
}

// DebruijnGraph::is_compressed (graph.rs:296-334): thread per (node, dir); the FIRST collapsible pair in the reference's
// iteration order (node ascending, Left before Right) wins through an atomicMin on ((node * 2 + dir) << 32 | next).
template <int W>
__global__ void is_compressed_kernel(KP kp, const u64* __restrict__ words, const u64* __restrict__ start, const u32* __restrict__ length,
                                     const u16* __restrict__ data, const u32* __restrict__ target, const u8* __restrict__ flags, u64 m,
                                     int stranded, int scmap, u64* __restrict__ result) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * m) return;
    const u64 i = t >> 1;
    const int dir = (int)(t & 1);
    auto single = [&](u64 n, int d, u32& nxt, int& ret) {
        int cnt = 0;
        for (int b = 0; b < 4; b++) {
            const u64 s = (n * 2 + d) * 4 + b;
            if (target[s] != NIL) { cnt++; nxt = target[s]; ret = flags[s] & 1; }
        }
        return cnt == 1;
    };
    u32 nxt = 0, back = 0;
    int ret = 0, r2 = 0;
    if (!single(i, dir, nxt, ret)) return;                       // dir_edges.len() == 1
    if (!single(nxt, ret, back, r2)) return;                     // ret_edges.len() == 1
    if (length[i] == (u32)kp.k && is_palindrome<W>(kp, kmer_at<W>(kp, words, start[i]))) return;      // we are a palindrome
    if (length[nxt] == (u32)kp.k && is_palindrome<W>(kp, kmer_at<W>(kp, words, start[nxt]))) return;  // the neighbour is
    if (i == nxt) return;                                        // smooth circle biting its own tail
    if (scmap && data[i] != data[nxt]) return;                   // spec.join_test (ScmapCompress: data equality)
    atomicMin(result, (t << 32) | nxt);
}

// The adjacency of every (node, side) under the Exts `d_exts` (the graph's own, or a pruned copy): first / last k-mers, the two
// sorted end maps, graph_edges_kernel.  All scratch lives in the ctx arena until the EdgeSet goes out of scope.
template <int W>
struct EdgeSet {
    DBuf<u64> fa_lo, fa_hi, fb_lo, fb_hi, la_lo, la_hi, lb_lo, lb_hi, lut_l, lut_r;
    DBuf<u32> ia, ib, ja, jb, cnt_l, cnt_r, target;
    DBuf<u8> flags;
    int build(Ctx* c, const Graph* g, const u8* d_exts) {
        const u64 m = g->n_nodes;
        cudaStream_t st = c->stream;
        KP kp = make_kp(g->k);
        TRY(fa_lo.alloc(c, m)); TRY(fb_lo.alloc(c, m)); TRY(la_lo.alloc(c, m)); TRY(lb_lo.alloc(c, m));
        if (W == 2) { TRY(fa_hi.alloc(c, m)); TRY(fb_hi.alloc(c, m)); TRY(la_hi.alloc(c, m)); TRY(lb_hi.alloc(c, m)); }
        TRY(ia.alloc(c, m)); TRY(ib.alloc(c, m)); TRY(ja.alloc(c, m)); TRY(jb.alloc(c, m));
        TRY(target.alloc(c, 8 * m)); TRY(flags.alloc(c, 8 * m));
        node_term_kmers_kernel<W><<<grid_for(m, 256), 256, 0, st>>>(kp, g->words, g->start, g->length, m, fa_lo.p, fa_hi.p, la_lo.p, la_hi.p, ia.p, ja.p);
        TRY(check_launch(c, "node_term_kmers"));
        u64 *flo, *fhi, *llo, *lhi;
        u32 *fid, *lid;
        TRY(radix_sort_pairs(c, W, 2 * g->k, m, fa_lo.p, fa_hi.p, ia.p, fb_lo.p, fb_hi.p, ib.p, &flo, &fhi, &fid));   // left_order
        TRY(radix_sort_pairs(c, W, 2 * g->k, m, la_lo.p, la_hi.p, ja.p, lb_lo.p, lb_hi.p, jb.p, &llo, &lhi, &lid));   // right_order
        EndMap<W> L, R;
        L.lo = flo; L.hi = fhi; L.node = fid;
        R.lo = llo; R.hi = lhi; R.node = lid;
        TRY(build_prefix_lut<W>(c, g->k, flo, fhi, m, cnt_l, lut_l, &L.shift));
        TRY(build_prefix_lut<W>(c, g->k, llo, lhi, m, cnt_r, lut_r, &R.shift));
        L.lut = lut_l.p; R.lut = lut_r.p;
        graph_edges_kernel<W><<<grid_for(2 * m, 256), 256, 0, st>>>(kp, g->words, g->start, g->length, d_exts, m, g->stranded, L, R,
                                                                    target.p, flags.p);
        return check_launch(c, "graph_edges");
    }
};

// h_target / h_flags != nullptr: copy the adjacency out (dbg_graph_edges); fix == 1: rewrite the graph's Exts from it
// (DebruijnGraph::fix_exts / get_valid_exts, graph.rs:337-377); fix == 2: is_compressed, *pair_out = -1 or (node << 32 | next)
template <int W>
static int graph_edges_impl(Ctx* c, Graph* g, u32* h_target, u8* h_flags, int fix, const u8* h_valid_nodes, int scmap = 0, long long* pair_out = nullptr) {
    const u64 m = g->n_nodes;
    if (pair_out) *pair_out = -1;
    if (m == 0) return DBG_OK;
    if (m >= (1ull << 31)) DBG_SET_ERR(c, DBG_E_BADARG, "too many nodes for 32-bit ids");
    TRY(arena_begin(c));
    cudaStream_t st = c->stream;
    KP kp = make_kp(g->k);
    EdgeSet<W> es;
    TRY(es.build(c, g, g->exts));
    if (h_target) {
        CU(c, cudaMemcpyAsync(h_target, es.target.p, 8 * m * sizeof(u32), cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(h_flags, es.flags.p, 8 * m, cudaMemcpyDeviceToHost, st));
    }
    if (fix == 2) {
        DBuf<u64> res;
        TRY(res.alloc(c, 1));
        TRY(res.fill_ff());
        is_compressed_kernel<W><<<grid_for(2 * m, 256), 256, 0, st>>>(kp, g->words, g->start, g->length, g->data, es.target.p, es.flags.p, m,
                                                                     g->stranded, scmap, res.p);
        TRY(check_launch(c, "is_compressed"));
        u64 h = 0;
        TRY(read_u64(c, res.p, &h));
        if (h != ~0ull) *pair_out = (long long)(((h >> 33) << 32) | (h & 0xffffffffull));
        return DBG_OK;
    }
    if (fix) {
        DBuf<u8> d_valid;
        if (h_valid_nodes) {
            TRY(d_valid.alloc(c, m));
            CU(c, cudaMemcpyAsync(d_valid.p, h_valid_nodes, m, cudaMemcpyHostToDevice, st));
        }
        fix_exts_kernel<<<grid_for(m, 256), 256, 0, st>>>(es.target.p, h_valid_nodes ? d_valid.p : nullptr, m, g->exts);
        TRY(check_launch(c, "fix_exts"));
        return sync(c);
    }
    return sync(c);
}

int graph_edges_dev(Ctx* c, const Graph* g, u32* h_target, u8* h_flags) {
    if (!g || !h_target || !h_flags) DBG_SET_ERR(c, DBG_E_BADARG, "null argument");
    Graph* gm = const_cast<Graph*>(g);   // not modified when fix == 0
    return g->k <= 32 ? graph_edges_impl<1>(c, gm, h_target, h_flags, 0, nullptr) : graph_edges_impl<2>(c, gm, h_target, h_flags, 0, nullptr);
}
int graph_is_compressed_dev(Ctx* c, const Graph* g, int scmap, long long* pair_out) {
    if (!g || !pair_out) DBG_SET_ERR(c, DBG_E_BADARG, "null argument");
    Graph* gm = const_cast<Graph*>(g);   // not modified
    return g->k <= 32 ? graph_edges_impl<1>(c, gm, nullptr, nullptr, 2, nullptr, scmap, pair_out) : graph_edges_impl<2>(c, gm, nullptr, nullptr, 2, nullptr, scmap, pair_out);
}
int graph_fix_exts_dev(Ctx* c, Graph* g, const u8* h_valid_nodes) {
    if (!g) DBG_SET_ERR(c, DBG_E_BADARG, "null graph");
    return g->k <= 32 ? graph_edges_impl<1>(c, g, nullptr, nullptr, 1, h_valid_nodes) : graph_edges_impl<2>(c, g, nullptr, nullptr, 1, h_valid_nodes);
}


// ================================================================================================
// BaseGraph::combine — src/graph.rs:71-100: every node of every graph re-added in order (PackedDnaStringSet::add,
// src/dna_string.rs:811-821: bases appended bit-contiguously), exts / data concatenated.  One thread per OUTPUT word gathers
// the <= 32 bases it holds from the node(s) that cover it.
// ================================================================================================
__global__ void combine_words_kernel(const u64* const* __restrict__ gwords, const u64* __restrict__ node_off, u32 n_graphs,
                                     const u64* __restrict__ src_start, const u64* __restrict__ dst_start, const u32* __restrict__ length,
                                     u64 m, u64 n_bases, u64* __restrict__ out, u64 n_words) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    u64 b = w * 32;
    const u64 end = b + 32 < n_bases ? b + 32 : n_bases;
    u64 lo = 0, hi = m;                                  // largest node with dst_start <= b
    while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (dst_start[mid] <= b) lo = mid; else hi = mid; }
    u64 i = lo, acc = 0;
    while (b < end) {
        while (dst_start[i] + length[i] <= b) i++;
        const u64 j = b - dst_start[i];
        const u64 left = dst_start[i] + length[i] - b;
        const u32 n = (u32)(end - b < left ? end - b : left);
        u32 ga = 0, gb = n_graphs;                         // graph that owns node i
        while (gb - ga > 1) { const u32 mid = (ga + gb) >> 1; if (node_off[mid] <= i) ga = mid; else gb = mid; }
        u64 x = bases64(gwords[ga], src_start[i] + j);
        if (n < 32) x &= ~0ull << (64 - 2 * n);
        acc |= x >> (2 * (b & 31));
        b += n;
    }
    out[w] = acc;
}

static Graph* new_graph(Ctx* c, int k, int stranded) {
    Graph* g = &(new dbg_graph())->g;
    g->ctx = c; g->k = k; g->stranded = stranded;
    return g;
}

int graph_combine_dev(Ctx* c, const Graph* const* gs, u32 n, Graph** out) {
    *out = nullptr;
    if (n == 0 || !gs) DBG_SET_ERR(c, DBG_E_BADARG, "combine needs at least one graph (K is a property of the graph handle)");
    u64 m = 0;
    int all_str = 1, any_str = 0;
    for (u32 i = 0; i < n; i++) {
        if (!gs[i]) DBG_SET_ERR(c, DBG_E_BADARG, "null graph");
        if (gs[i]->k != gs[0]->k) DBG_SET_ERR(c, DBG_E_BADARG, "graphs of different K");
        all_str &= gs[i]->stranded != 0; any_str |= gs[i]->stranded != 0;
        m += gs[i]->n_nodes;
    }
    if (any_str && !all_str) DBG_SET_ERR(c, DBG_E_BADARG, "attempted to combine stranded and unstranded graphs");   // graph.rs:90-92
    if (m >= (1ull << 31)) DBG_SET_ERR(c, DBG_E_BADARG, "too many nodes for 32-bit ids");
    Graph* og = new_graph(c, gs[0]->k, all_str);
    *out = og;
    if (m == 0) return DBG_OK;
    TRY(arena_begin(c));
    cudaStream_t st = c->stream;
    DBuf<u64> words, ostart, src_start, d_off, d_tot;
    DBuf<const u64*> d_gw;
    DBuf<u32> olen;
    DBuf<u8> oexts;
    DBuf<u16> odata;
    int rc = ostart.alloc_pool(c, m);
    if (rc == DBG_OK) rc = olen.alloc_pool(c, m);
    if (rc == DBG_OK) rc = oexts.alloc_pool(c, m);
    if (rc == DBG_OK) rc = odata.alloc_pool(c, m);
    if (rc == DBG_OK) rc = src_start.alloc(c, m);
    if (rc == DBG_OK) rc = d_off.alloc(c, n + 1);
    if (rc == DBG_OK) rc = d_gw.alloc(c, n);
    if (rc == DBG_OK) rc = d_tot.alloc(c, 1);
    if (rc != DBG_OK) { free_graph(og); *out = nullptr; return rc; }
    std::vector<u64> h_off(n + 1);
    std::vector<const u64*> h_gw(n);
    u64 pos = 0;
    for (u32 i = 0; i < n; i++) {
        const Graph* g = gs[i];
        h_off[i] = pos; h_gw[i] = g->words;
        if (g->n_nodes) {
            cudaMemcpyAsync(olen.p + pos, g->length, g->n_nodes * 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(oexts.p + pos, g->exts, g->n_nodes, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(odata.p + pos, g->data, g->n_nodes * 2, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(src_start.p + pos, g->start, g->n_nodes * 8, cudaMemcpyDeviceToDevice, st);
        }
        pos += g->n_nodes;
    }
    h_off[n] = pos;
    cudaMemcpyAsync(d_off.p, h_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_gw.p, h_gw.data(), n * sizeof(u64*), cudaMemcpyHostToDevice, st);
    u64 nb = 0;
    rc = exclusive_scan_u32_to_u64(c, olen.p, ostart.p, m, d_tot.p);
    if (rc == DBG_OK) rc = read_u64(c, d_tot.p, &nb);     // also orders the pageable host vectors before they go out of scope
    if (rc == DBG_OK) {
        og->n_nodes = m; og->n_bases = nb; og->n_words = (nb + 31) / 32;
        rc = words.alloc_pool(c, og->n_words + 3);
    }
    if (rc == DBG_OK) {
        cudaMemsetAsync(words.p + og->n_words, 0, 3 * 8, st);
        combine_words_kernel<<<grid_for(og->n_words, 256), 256, 0, st>>>(d_gw.p, d_off.p, n, src_start.p, ostart.p, olen.p, m, nb, words.p, og->n_words);
        rc = check_launch(c, "combine_words");
    }
    if (rc == DBG_OK) rc = sync(c);
    if (rc != DBG_OK) { og->n_nodes = 0; free_graph(og); *out = nullptr; return rc; }
    og->words = words.take(); og->start = ostart.take(); og->length = olen.take(); og->exts = oexts.take(); og->data = odata.take();
    return DBG_OK;
}

// ================================================================================================
// compression::compress_graph — src/compression.rs:100-349 (CompressFromGraph).  The reference walks greedily from every
// still-available node in node order (Left, then Right), consulting and clearing an `available_nodes` bit set.  Device form:
//   1. fix_exts(Some(available))  (:309)   adjacency of every (node, side) + pruned Exts copy (EdgeSet, fix_exts_kernel)
//   2. links      the stateless form of try_extend_node (:115-205) per PORT s = 2 * node + d (d = direction of travel):
//                 unique extension, node not a K-long palindrome, target available and not the node itself, the k-mer we
//                 arrive on not a palindrome, join_test, exactly one extension back.  For graphs whose links are symmetric
//                 (checked: anything produced by compress_kmers / combine) the components are simple paths / cycles of nodes
//                 and the greedy result is: one output node per component, seed = its smallest node id, seed in forward
//                 orientation, a cycle opened so that the seed ends up at its right end (the Left walk runs first, :241-242).
//   3. rank       Wyllie pointer doubling over the 2M ports, 16-byte records (next, smallest port id, bases to the chain end);
//                 ports still active after ceil(log2 2M) + 1 rounds lie on cycles: cut at the seed, rank again.
//   4. place      per node: seed, orientation, base offset inside the output node; seeds -> node order / base offsets by scans
//   5. emit       work items of <= 1024 bases of one source node: (reverse-complemented) bases written at their destination
//                 bit offset (sequence_of_path, graph.rs:471-491: every node but the first drops its first K-1 bases);
//                 end nodes supply the Exts (:266-276), data reduced per output node (spec.reduce)
//   6. finish + fix_exts(None)  (:330-331)
// ================================================================================================
static const u32 TERM = 0xffffffffu;
struct __align__(16) PRec { u32 nxt, mn; u64 ds; };

template <int W>
__global__ void cg_links_kernel(KP kp, const u64* __restrict__ words, const u64* __restrict__ start, const u32* __restrict__ length,
                                const u8* __restrict__ fexts, const u16* __restrict__ data, const u8* __restrict__ valid,
                                const u32* __restrict__ target, const u8* __restrict__ flags, u64 m, int stranded, int scmap,
                                u32* __restrict__ link, u32* __restrict__ err) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * m) return;
    const u64 i = t >> 1;
    const int d = (int)(t & 1);
    const u32 K = (u32)kp.k;
    u32 out = TERM;
    do {
        if (valid && !valid[i]) break;                                         // censored nodes are never walked
        const u32 nib = exts_side(fexts[i], d);
        if (popc4(nib) != 1) break;                                            // :120
        if (!stranded && length[i] == K && is_palindrome<W>(kp, kmer_at<W>(kp, words, start[i]))) break;   // :121
        const u64 slot = t * 4 + unique_base(nib);
        const u32 nid = target[slot];
        if (nid == NIL) { atomicMax(err, 3u); break; }                         // panic!("No kmer") :138 — not after fix_exts
        const int inc = flags[slot] & 1;
        if (valid && !valid[nid]) break;                                       // :173
        if (nid == i) break;                                                   // extend_node removed the start node first (:214)
        // next_kmer (:129) is the neighbour's end k-mer on the side we arrive at, or its rc: a palindrome either way (:174)
        if (!stranded && is_palindrome<W>(kp, kmer_at<W>(kp, words, inc ? start[nid] + length[nid] - K : start[nid]))) break;
        if (scmap && data[i] != data[nid]) break;                              // spec.join_test :175
        const int cnt = popc4(exts_side(fexts[nid], inc));                     // incoming_count :187
        if (cnt == 0) { atomicMax(err, 1u); break; }                           // panic!("unreachable") :195
        if (cnt == 1) out = nid * 2 + (u32)(inc ^ 1);                          // Unique(next, next_side_outgoing) :198
    } while (0);
    link[t] = out;
}

// s = (A, d) -> l = (B, out): the port that travels back out of B through the side we came in, l ^ 1, must lead to (A, d ^ 1)
__global__ void cg_symmetry_kernel(const u32* __restrict__ link, u64 n_ports, u32* __restrict__ err) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_ports) return;
    const u32 l = link[s];
    if (l != TERM && link[l ^ 1u] != (u32)(s ^ 1)) atomicMax(err, 5u);
}

__global__ void cg_init_kernel(const u32* __restrict__ link, const u32* __restrict__ length, int K, u64 n_ports, PRec* __restrict__ rec) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_ports) return;
    PRec r;
    r.nxt = link[s]; r.mn = (u32)s; r.ds = (u64)(length[s >> 1] - (u32)(K - 1));
    rec[s] = r;
}

__global__ void cg_round_kernel(const PRec* __restrict__ src, PRec* __restrict__ dst, u64 n_ports, u32* __restrict__ active) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool on = false;
    if (s < n_ports) {
        PRec r = src[s];
        if (r.nxt != TERM) {
            const PRec q = src[r.nxt];
            r.mn = q.mn < r.mn ? q.mn : r.mn;
            r.ds += q.ds;
            r.nxt = q.nxt;
            on = r.nxt != TERM;
        }
        dst[s] = r;
    }
    if (__any_sync(0xffffffffu, on) && (threadIdx.x & 31) == 0) *active = 1u;
}

// after the last round only ports on cycles are still active; the thread of port (seed, Right) opens its cycle
__global__ void cg_cut_cycles_kernel(const PRec* __restrict__ rec, u64 n_ports, u32* __restrict__ link) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_ports) return;
    const PRec r = rec[s];
    if (r.nxt == TERM || r.mn != (u32)s || !(s & 1)) return;
    const u32 a = link[s];
    link[s] = TERM;            // the Right walk from the seed finds its neighbour already used
    link[a ^ 1u] = TERM;       // ... and the Left walk, having gone all the way round, finds the seed used
}

__global__ void cg_place_kernel(const PRec* __restrict__ rec, const u32* __restrict__ link, const u32* __restrict__ length,
                                const u8* __restrict__ valid, int K, u64 m, u32* __restrict__ seed_of, u64* __restrict__ off,
                                u8* __restrict__ nflags, u32* __restrict__ is_seed, u64* __restrict__ seed_len, u32* __restrict__ chunks,
                                u32* __restrict__ err) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (valid && !valid[i]) { seed_of[i] = TERM; off[i] = 0; nflags[i] = 0; is_seed[i] = 0; seed_len[i] = 0; chunks[i] = 0; return; }
    const PRec ra = rec[2 * i], rb = rec[2 * i + 1];
    const u32 sa = ra.mn >> 1, sb = rb.mn >> 1;
    const u32 seed = sa < sb ? sa : sb;
    bool fwd;
    if (seed == (u32)i) fwd = true;                       // the seed keeps its stored orientation (node_path starts (seed, Dir::Left), :248)
    else if (sa == seed) fwd = (ra.mn & 1u) == 0u;        // travelling Left out of i we pass the seed travelling Left: same orientation
    else fwd = (rb.mn & 1u) == 1u;
    const u64 w = (u64)(length[i] - (u32)(K - 1));
    const u64 dl = fwd ? ra.ds : rb.ds, dr = fwd ? rb.ds : ra.ds;   // bases from i to the output's left / right end (overlaps removed)
    const u32 pl = (u32)(2 * i) + (fwd ? 0u : 1u);
    const bool leftmost = link[pl] == TERM, rightmost = link[pl ^ 1u] == TERM;
    seed_of[i] = seed;
    off[i] = dl - w;
    nflags[i] = (u8)((fwd ? 1 : 0) | (leftmost ? 2 : 0) | (rightmost ? 4 : 0));
    const u32 cnt = length[i] - (leftmost ? 0u : (u32)(K - 1));
    chunks[i] = (cnt + 1023u) / 1024u;
    if (seed == (u32)i) {
        const u64 total = dl + dr - w + (u64)(K - 1);
        if (total > 0xffffffffull) atomicMax(err, 6u);    // BaseGraph lengths are u32 (dna_string.rs:766)
        is_seed[i] = 1; seed_len[i] = total;
    } else { is_seed[i] = 0; seed_len[i] = 0; }
}

__global__ void cg_emit_kernel(const u64* __restrict__ words, const u64* __restrict__ start, const u32* __restrict__ length,
                               const u32* __restrict__ seed_of, const u64* __restrict__ off, const u8* __restrict__ nflags,
                               const u64* __restrict__ base_start, const u64* __restrict__ chunk_off, u64 m, u64 n_items, int K,
                               u64* __restrict__ out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    u64 lo = 0, hi = m;                                   // the node whose chunk range holds item t (chunk_off has m + 1 entries)
    while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (chunk_off[mid] <= t) lo = mid; else hi = mid; }
    while (chunk_off[lo + 1] <= t) lo++;                   // (ties: nodes without chunks)
    const u64 i = lo;
    const u32 f = nflags[i], len = length[i];
    const bool fwd = f & 1u;
    const u32 skip = (f & 2u) ? 0u : (u32)(K - 1);
    const u64 D = base_start[seed_of[i]] + off[i];        // destination of the node's oriented base 0
    u32 j = skip + (u32)(t - chunk_off[i]) * 1024u;
    const u32 j1 = j + 1024u < len ? j + 1024u : len;
    const u64 s0 = start[i];
    while (j < j1) {
        const u64 db = D + j;
        const u32 room = 32u - (u32)(db & 31);
        const u32 n = j1 - j < room ? j1 - j : room;
        u64 x;
        if (fwd) x = bases64(words, s0 + j);
        else {                                             // oriented bases j .. j+n-1 = rc of source bases len-j-n .. len-j-1
            x = rev2_64(~bases64(words, s0 + len - j - n));
            if (n < 32) x <<= 64 - 2 * n;
        }
        if (n < 32) x &= ~0ull << (64 - 2 * n);
        x >>= 2 * (db & 31);
        if (n == 32) out[db >> 5] = x; else atomicOr(&out[db >> 5], x);
        j += n;
    }
}

__global__ void cg_node_meta_kernel(const u32* __restrict__ seed_of, const u8* __restrict__ nflags, const u8* __restrict__ fexts,
                                    const u16* __restrict__ data, const u64* __restrict__ node_ord, const u64* __restrict__ base_start,
                                    const u64* __restrict__ seed_len, u64 m, int reduce_op, u8* __restrict__ lext, u8* __restrict__ rext,
                                    u8* __restrict__ single, u64* __restrict__ acc, u64* __restrict__ out_start, u32* __restrict__ out_len) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u32 seed = seed_of[i];
    if (seed == TERM) return;
    const u64 o = node_ord[seed];
    const u32 f = nflags[i], e = fexts[i];
    const bool fwd = f & 1u;
    if (f & 2u) lext[o] = (u8)(fwd ? (e & 0xfu) : (exts_complement(e) >> 4));          // :266-270
    if (f & 4u) rext[o] = (u8)(fwd ? (e >> 4) : (exts_complement(e) & 0xfu));          // :272-276
    if ((f & 6u) == 6u) single[o] = 1;
    if (reduce_op == DBG_REDUCE_MAX || reduce_op == DBG_REDUCE_SCMAP) atomicMax(&acc[o], (u64)data[i]);
    else atomicAdd(&acc[o], (u64)data[i]);
    if (seed == (u32)i) { out_start[o] = base_start[i]; out_len[o] = (u32)seed_len[i]; }
}

__global__ void cg_finalize_kernel(const u8* __restrict__ lext, const u8* __restrict__ rext, const u8* __restrict__ single,
                                   const u64* __restrict__ acc, u64 mo, int reduce_op, u8* __restrict__ exts, u16* __restrict__ data) {
    const u64 o = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= mo) return;
    exts[o] = (u8)((rext[o] << 4) | (lext[o] & 0xfu));                                  // Exts::from_single_dirs, lib.rs:591-595
    const u64 a = acc[o];
    u16 d;
    switch (reduce_op) {                                                                // the closures of SimpleCompress folded over the node
        case DBG_REDUCE_SAT_ADD: d = (u16)(a > 65535 ? 65535 : a); break;
        case DBG_REDUCE_WRAP_ADD: d = (u16)a; break;
        case DBG_REDUCE_ADD_MOD_65535: d = single[o] ? (u16)a : (u16)(a % 65535); break;
        default: d = (u16)a; break;
    }
    data[o] = d;
}

template <int W>
static int compress_graph_impl(Ctx* c, const Graph* g, int stranded, int reduce_op, const u64* h_censor, u64 n_censor, Graph* og) {
    const u64 m = g->n_nodes, np = 2 * m;
    const int K = g->k;
    cudaStream_t st = c->stream;
    KP kp = make_kp(K);
    DBuf<u8> valid, fexts;
    if (n_censor) {
        std::vector<u8> hv(m, 1);
        for (u64 i = 0; i < n_censor; i++) {
            if (h_censor[i] >= m) DBG_SET_ERR(c, DBG_E_BADARG, "censor_nodes[%llu] = %llu: no such node", i, h_censor[i]);
            hv[h_censor[i]] = 0;
        }
        TRY(valid.alloc(c, m));
        CU(c, cudaMemcpy(valid.p, hv.data(), m, cudaMemcpyHostToDevice));
    }
    TRY(fexts.alloc(c, m));
    EdgeSet<W> es;
    TRY(es.build(c, g, g->exts));
    fix_exts_kernel<<<grid_for(m, 256), 256, 0, st>>>(es.target.p, valid.p, m, fexts.p);                       // :309
    TRY(check_launch(c, "fix_exts"));
    DBuf<u32> link, flag;
    TRY(link.alloc(c, np)); TRY(flag.alloc(c, 2)); TRY(flag.zero());
    cg_links_kernel<W><<<grid_for(np, 256), 256, 0, st>>>(kp, g->words, g->start, g->length, fexts.p, g->data, valid.p, es.target.p,
                                                         es.flags.p, m, stranded, reduce_op == DBG_REDUCE_SCMAP, link.p, flag.p);
    TRY(check_launch(c, "cg_links"));
    cg_symmetry_kernel<<<grid_for(np, 256), 256, 0, st>>>(link.p, np, flag.p);
    TRY(check_launch(c, "cg_symmetry"));
    u64 hf = 0;
    TRY(read_u64(c, flag.p, &hf));
    const u32 e0 = (u32)hf;
    if (e0 == 1) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "compress_graph: a unique extension leads into a node with no extension back (\"unreachable\", compression.rs:195)");
    if (e0 == 3) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "compress_graph: extension without a link after fix_exts (\"No kmer\", compression.rs:138)");
    if (e0 == 5) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "compress_graph: non-reciprocal node links (Exts of neighbouring nodes disagree, or a node longer than K "
                             "ends in a palindromic k-mer): the greedy walk's result would depend on its order");
    // ---- rank
    DBuf<PRec> ra, rb;
    TRY(ra.alloc(c, np)); TRY(rb.alloc(c, np));
    int max_rounds = 2;
    while ((1ull << (max_rounds - 1)) < np) max_rounds++;
    PRec* cur = nullptr;
    for (int attempt = 0; attempt < 2; attempt++) {
        cg_init_kernel<<<grid_for(np, 256), 256, 0, st>>>(link.p, g->length, K, np, ra.p);
        TRY(check_launch(c, "cg_init"));
        cur = ra.p;
        PRec* nxt = rb.p;
        u32 active = 1;
        for (int r = 0; r < max_rounds && active; r++) {
            CU(c, cudaMemsetAsync(flag.p + 1, 0, 4, st));
            cg_round_kernel<<<grid_for(np, 256), 256, 0, st>>>(cur, nxt, np, flag.p + 1);
            TRY(check_launch(c, "cg_round"));
            TRY(read_u64(c, flag.p, &hf));
            active = (u32)(hf >> 32);
            PRec* t = cur; cur = nxt; nxt = t;
            c->stats.rank_rounds++;
        }
        if (!active) break;
        if (attempt == 1) DBG_SET_ERR(c, DBG_E_INTERNAL, "compress_graph: ranking did not converge after the cycles were opened");
        cg_cut_cycles_kernel<<<grid_for(np, 256), 256, 0, st>>>(cur, np, link.p);
        TRY(check_launch(c, "cg_cut_cycles"));
    }
    // ---- place
    DBuf<u32> seed_of, is_seed, chunks;
    DBuf<u64> off, seed_len, node_ord, base_start, chunk_off, tot;
    DBuf<u8> nflags;
    TRY(seed_of.alloc(c, m)); TRY(is_seed.alloc(c, m)); TRY(chunks.alloc(c, m)); TRY(off.alloc(c, m)); TRY(seed_len.alloc(c, m));
    TRY(node_ord.alloc(c, m)); TRY(base_start.alloc(c, m)); TRY(chunk_off.alloc(c, m + 1)); TRY(tot.alloc(c, 3)); TRY(nflags.alloc(c, m));
    cg_place_kernel<<<grid_for(m, 256), 256, 0, st>>>(cur, link.p, g->length, valid.p, K, m, seed_of.p, off.p, nflags.p, is_seed.p, seed_len.p,
                                                     chunks.p, flag.p);
    TRY(check_launch(c, "cg_place"));
    TRY(exclusive_scan_u32_to_u64(c, is_seed.p, node_ord.p, m, tot.p));
    TRY(exclusive_scan_u64(c, seed_len.p, base_start.p, m, tot.p + 1));
    TRY(exclusive_scan_u32_to_u64(c, chunks.p, chunk_off.p, m, tot.p + 2));
    CU(c, cudaMemcpyAsync(chunk_off.p + m, tot.p + 2, 8, cudaMemcpyDeviceToDevice, st));
    u64 h[3];
    TRY(read_u64(c, tot.p, h, 3));
    TRY(read_u64(c, flag.p, &hf));
    if ((u32)hf == 6) DBG_SET_ERR(c, DBG_E_BADARG, "compress_graph: a merged node is longer than 2^32 - 1 bases");
    const u64 mo = h[0], nb = h[1], n_items = h[2];
    og->n_nodes = mo; og->n_bases = nb; og->n_words = (nb + 31) / 32;
    if (mo == 0) return DBG_OK;
    DBuf<u64> words, ostart, acc;
    DBuf<u32> olen;
    DBuf<u8> oexts, lext, rext, single;
    DBuf<u16> odata;
    TRY(words.alloc_pool(c, og->n_words + 3)); TRY(words.zero());
    TRY(ostart.alloc_pool(c, mo)); TRY(olen.alloc_pool(c, mo)); TRY(oexts.alloc_pool(c, mo)); TRY(odata.alloc_pool(c, mo));
    TRY(acc.alloc(c, mo)); TRY(acc.zero()); TRY(lext.alloc(c, mo)); TRY(rext.alloc(c, mo)); TRY(single.alloc(c, mo)); TRY(single.zero());
    cg_emit_kernel<<<grid_for(n_items, 256), 256, 0, st>>>(g->words, g->start, g->length, seed_of.p, off.p, nflags.p, base_start.p, chunk_off.p,
                                                          m, n_items, K, words.p);
    TRY(check_launch(c, "cg_emit"));
    cg_node_meta_kernel<<<grid_for(m, 256), 256, 0, st>>>(seed_of.p, nflags.p, fexts.p, g->data, node_ord.p, base_start.p, seed_len.p, m, reduce_op,
                                                         lext.p, rext.p, single.p, acc.p, ostart.p, olen.p);
    TRY(check_launch(c, "cg_node_meta"));
    cg_finalize_kernel<<<grid_for(mo, 256), 256, 0, st>>>(lext.p, rext.p, single.p, acc.p, mo, reduce_op, oexts.p, odata.p);
    TRY(check_launch(c, "cg_finalize"));
    TRY(sync(c));
    og->words = words.take(); og->start = ostart.take(); og->length = olen.take(); og->exts = oexts.take(); og->data = odata.take();
    return DBG_OK;
}

int compress_graph_dev(Ctx* c, const Graph* g, int stranded, int reduce_op, const u64* h_censor, u64 n_censor, Graph** out) {
    *out = nullptr;
    if (!g) DBG_SET_ERR(c, DBG_E_BADARG, "null graph");
    if (reduce_op < 0 || reduce_op > DBG_REDUCE_SCMAP) DBG_SET_ERR(c, DBG_E_BADARG, "unknown reduce_op %d", reduce_op);
    if (n_censor && !h_censor) DBG_SET_ERR(c, DBG_E_BADARG, "null censor_nodes");
    if (g->n_nodes >= (1ull << 31)) DBG_SET_ERR(c, DBG_E_BADARG, "too many nodes for 32-bit ids");
    Graph* og = new_graph(c, g->k, stranded != 0);       // BaseGraph::new(stranded) :320
    if (g->n_nodes == 0) { *out = og; return DBG_OK; }
    int rc = arena_begin(c);
    if (rc == DBG_OK) rc = g->k <= 32 ? compress_graph_impl<1>(c, g, stranded, reduce_op, h_censor, n_censor, og)
                                      : compress_graph_impl<2>(c, g, stranded, reduce_op, h_censor, n_censor, og);
    // graph.finish(); dbg.fix_exts(None)  :330-331  (all scratch of the stage above has been released)
    if (rc == DBG_OK && og->n_nodes) rc = g->k <= 32 ? graph_edges_impl<1>(c, og, nullptr, nullptr, 1, nullptr) : graph_edges_impl<2>(c, og, nullptr, nullptr, 1, nullptr);
    if (rc != DBG_OK) { if (!og->words) og->n_nodes = 0; free_graph(og); return rc; }
    c->stats.n_nodes = og->n_nodes; c->stats.n_bases = og->n_bases;
    *out = og;
    return DBG_OK;
}

}  // namespace dbg
