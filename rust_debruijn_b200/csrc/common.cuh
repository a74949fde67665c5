// Host-side plumbing shared by the stage files: context, error propagation, stream-ordered device
// buffers, scan/sort entry points.  No exceptions cross the C ABI: every stage returns a dbg status.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <time.h>
#include <string>
#include <vector>

#include "../../include/dbg_b200.h"
#include "kmer.cuh"

namespace dbg {

struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host->device uploads that overlap the partition kernel
    cudaMemPool_t pool = nullptr;
    std::string err;
    dbg_stats stats;
    // tunables (dbg_ctx_set_param)
    int msp_p = 0;             // 0 = auto
    int target_bucket_occ = 0; // 0 = auto (k-mer occurrences per MSP bucket)
    int dedup = 1;             // deduplicate super-k-mer records per bucket before counting (K <= 32)
    int direct_partition = 1;  // records straight into per-bucket regions sized by a sampling pass (0 = always stage + scatter)
    u64 direct_min_tiles = 2048;  // ... only for inputs of at least this many 4096-base tiles
    int pipelined_direct_failed = 0;  // a direct partition sized from the first upload chunk overflowed once: keep pipelined inputs on staging
    int no_fast_compress = 0;  // testing: 1 = always take the general (per-k-mer rank + emit) compression path
    u64 valid_est_div = 0;     // testing: size the valid-k-mer buffer as N / div (0 = default estimate N/8 + 2^20)
    u64 mem_budget_bytes = 0;  // scratch budget for the pass planner (0 = 60% of free device memory, capped by memory_size)
    u64 launches = 0;          // kernels launched by this ctx (gpu_launches in bench.py)
    // pinned scratch for small D2H reads
    u64* h_scratch = nullptr;
    // scratch arena: one device slab, bump/LIFO allocated per top-level call, regrown between calls when a
    // call needed more than it holds (the overflow of that call is served by the pool)
    char* arena = nullptr;
    u64 arena_size = 0, arena_off = 0, arena_want = 0;
    // timing events
    cudaEvent_t ev[12];
};

#define DBG_SET_ERR(ctx, code, ...)                            \
    do {                                                       \
        char _b[512];                                          \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                 \
        (ctx)->err = _b;                                       \
        return (code);                                         \
    } while (0)

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            int _c = (_e == cudaErrorMemoryAllocation) ? DBG_E_OOM : DBG_E_CUDA;                        \
            DBG_SET_ERR(ctx, _c, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
        }                                                                                               \
    } while (0)

#define TRY(call)                 \
    do {                          \
        int _s = (call);          \
        if (_s != DBG_OK) return _s; \
    } while (0)

// Device buffer.  alloc(): scratch from the ctx arena (no driver call in steady state; LIFO release).
// alloc_pool(): result storage from the ctx memory pool (cudaMallocAsync), may be handed out with take().
template <typename T>
struct DBuf {
    T* p = nullptr;
    u64 n = 0;
    Ctx* ctx = nullptr;
    u64 arena_prev = ~0ull;  // arena offset to restore on release; ~0 = not an arena buffer
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    int alloc(Ctx* c, u64 count) {
        release();
        ctx = c;
        n = count;
        u64 bytes = ((count ? count : 1) * sizeof(T) + 255) & ~255ull;
        if (c->arena_off + bytes <= c->arena_size) {
            arena_prev = c->arena_off;
            p = reinterpret_cast<T*>(c->arena + c->arena_off);
            c->arena_off += bytes;
            if (c->arena_off > c->arena_want) c->arena_want = c->arena_off;
            return DBG_OK;
        }
        // does not fit: remember the demand (as if it had been bump-allocated) and serve from the pool
        u64 would = c->arena_off + bytes;
        if (would > c->arena_want) c->arena_want = would;
        arena_prev = ~0ull;
        CU(c, cudaMallocAsync((void**)&p, bytes, c->pool, c->stream));
        return DBG_OK;
    }
    int alloc_pool(Ctx* c, u64 count) {
        release();
        ctx = c;
        n = count;
        arena_prev = ~0ull;
        u64 bytes = (count ? count : 1) * sizeof(T);
        CU(c, cudaMallocAsync((void**)&p, bytes, c->pool, c->stream));
        return DBG_OK;
    }
    int zero() {
        CU(ctx, cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx->stream));
        return DBG_OK;
    }
    int fill_ff() {
        CU(ctx, cudaMemsetAsync(p, 0xff, (n ? n : 1) * sizeof(T), ctx->stream));
        return DBG_OK;
    }
    void release() {
        if (p && ctx) {
            if (arena_prev != ~0ull) {
                // LIFO: give the space back only when this is the top of the arena (stream order keeps reuse safe)
                u64 bytes = ((n ? n : 1) * sizeof(T) + 255) & ~255ull;
                if (arena_prev + bytes == ctx->arena_off) ctx->arena_off = arena_prev;
            } else {
                cudaFreeAsync(p, ctx->stream);
            }
        }
        p = nullptr;
        n = 0;
        arena_prev = ~0ull;
    }
    T* take() {  // pool buffers only
        if (arena_prev != ~0ull) { fprintf(stderr, "dbg: take() on an arena buffer\n"); abort(); }
        T* q = p; p = nullptr; n = 0; return q;
    }
};

// Called at the start of every top-level stage call: empty the arena, regrow it if the last call overflowed.
inline int arena_begin(Ctx* c) {
    c->arena_off = 0;
    if (c->arena_want > c->arena_size) {
        CU(c, cudaStreamSynchronize(c->stream));
        if (c->arena) CU(c, cudaFree(c->arena));
        c->arena = nullptr;
        c->arena_size = 0;
        u64 want = c->arena_want + c->arena_want / 8 + (64ull << 20);
        cudaError_t e = cudaMalloc((void**)&c->arena, want);
        if (e == cudaSuccess) c->arena_size = want;
        else { cudaGetLastError(); c->arena_want = 0; }  // stay on the pool
    }
    return DBG_OK;
}

// Host waits poll cudaStreamQuery instead of blocking in cudaStreamSynchronize: the path has a few short waits per call
// (sizes read back between stages), and on a shared host a thread that went to sleep in the driver — or that gave up
// its time slice with sched_yield(), measured in round 2: isolated 35-95 ms steps — can take tens of milliseconds to be
// scheduled again, far longer than the kernels it waits for.  Every wait of the steady-state path ends within a few
// milliseconds, so the poll is a pure spin for the first 50 ms; only a wait that outlives that (a cold start, a huge
// input) starts yielding between polls so that it does not pin a core for seconds.
inline cudaError_t spin_sync(cudaStream_t st) {
    cudaError_t e;
    timespec t0, t1;
    bool timed = false, polite = false;
    unsigned spins = 0;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
        if (polite) { sched_yield(); continue; }
        if ((++spins & 1023u) == 0) {
            if (!timed) { clock_gettime(CLOCK_MONOTONIC, &t0); timed = true; }
            else {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                polite = (t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec) > 50000000ll;
            }
        }
    }
    return e;
}
inline int sync(Ctx* c) {
    CU(c, spin_sync(c->stream));
    return DBG_OK;
}
// Read a few u64 from the device (blocking).
inline int read_u64(Ctx* c, const void* dptr, u64* out, int count = 1) {
    CU(c, cudaMemcpyAsync(c->h_scratch, dptr, 8 * count, cudaMemcpyDeviceToHost, c->stream));
    CU(c, spin_sync(c->stream));
    for (int i = 0; i < count; i++) out[i] = c->h_scratch[i];
    return DBG_OK;
}
inline int check_launch(Ctx* c, const char* what) {
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) DBG_SET_ERR(c, DBG_E_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
    return DBG_OK;
}
inline u32 grid_for(u64 n, u32 block) { return (u32)((n + block - 1) / block); }

// ---- device-resident objects behind the opaque C handles -------------------------------------------
struct SeqSet {
    Ctx* ctx;
    u64* words = nullptr;   // DnaString words (+2 words of zero padding)
    u64 n_words = 0;
    u64* start = nullptr;   // may be null when uniform_len != 0 (start = i * uniform_len)
    u32* length = nullptr;  // may be null when uniform_len != 0
    u8* seq_exts = nullptr; // may be null (=> Exts::empty())
    u64 n_seqs = 0;
    u32 uniform_len = 0;
    u32 max_len = 0;
    bool contiguous = false;  // start[i+1] == start[i] + length[i] (PackedDnaStringSet::add layout)
    u64 base0 = 0, total_end = 0;  // global base range [start[0], start[n-1] + length[n-1]) when contiguous
    bool owned = true;
    // pipelined upload (fused host entry points): chunk i of `words` is on the device once pend_ev[i] fired
    int n_pending = 0;
    cudaEvent_t pend_ev[8];
    u64 pend_words_end[8];
};

struct Table {
    Ctx* ctx;
    int k = 0;
    u64 n = 0, n_all = 0;
    u64 n_input = 0;      // input k-mer occurrences (filter.rs:152-155)
    u64* lo = nullptr;    // ascending canonical k-mers
    u64* hi = nullptr;    // k > 32 only
    u8* exts = nullptr;
    u16* counts = nullptr;
    u64* all_lo = nullptr;
    u64* all_hi = nullptr;
    u64* colors = nullptr;  // CountFilterSet only: bit c = label c was observed for the k-mer
};

struct Graph {
    Ctx* ctx;
    int k = 0;
    int stranded = 0;
    u64 n_nodes = 0, n_bases = 0, n_words = 0;
    u64* words = nullptr;
    u64* start = nullptr;
    u32* length = nullptr;
    u8* exts = nullptr;
    u16* data = nullptr;
};

// Bucket-contiguous super-k-mer records of one rank's reads (multi-GPU building block).
struct Partition {
    Ctx* ctx;
    int k = 0, p = 0, bbits = 0, rec_words = 2;
    u64 n_input = 0, n_rec = 0;
    u64* rec = nullptr;          // n_rec * rec_words u64
    u32* bucket_count = nullptr; // 2^bbits
    u64* bucket_off = nullptr;   // 2^bbits + 1
};

// Make every pending upload chunk of a sequence set visible to the ctx stream.
inline int seqset_ready(Ctx* c, SeqSet* s) {
    for (int i = 0; i < s->n_pending; i++) CU(c, cudaStreamWaitEvent(c->stream, s->pend_ev[i], 0));
    s->n_pending = 0;
    return DBG_OK;
}

}  // namespace dbg
// the opaque C handles are thin wrappers (first member) so stage code can allocate them directly
struct dbg_ctx { dbg::Ctx c; };
struct dbg_seqset { dbg::SeqSet s; };
struct dbg_kmer_table { dbg::Table t; };
struct dbg_graph { dbg::Graph g; };
struct dbg_partition { dbg::Partition p; };
namespace dbg {

// ---- primitives (scan_sort.cu) ------------------------------------------------------------------------
// out[i] = sum_{j<i} in[j]; total written to *d_total (device) if non-null.  In-place allowed for u64.
int exclusive_scan_u32_to_u64(Ctx* c, const u32* in, u64* out, u64 n, u64* d_total);
int exclusive_scan_u64(Ctx* c, const u64* in, u64* out, u64 n, u64* d_total);
// Stable LSD radix sort of (key words, 32-bit payload) by ascending key.  key_bits = significant bits.
// Buffers are ping-ponged; on return *res_* point at the buffers holding the result.
int radix_sort_pairs(Ctx* c, int W, int key_bits, u64 n, u64* lo_a, u64* hi_a, u32* val_a, u64* lo_b, u64* hi_b,
                     u32* val_b, u64** res_lo, u64** res_hi, u32** res_val);

// ---- stages ---------------------------------------------------------------------------------------------
int filter_kmers_dev(Ctx* c, int k, const SeqSet* s, u32 min_obs, int stranded, int report_all, u64 mem_gb,
                     Table** out);
int compress_dev(Ctx* c, const Table* t, int stranded, int reduce_op, Graph** out);
int filter_kmers_colorset_dev(Ctx* c, int k, const SeqSet* s, const u8* h_labels, u32 min_obs, int stranded, u64 mem_gb, Table** out);
void plan_filter(const Ctx* c, int k, u64 N, int* p_out, int* bbits_out);
int partition_reads_dev(Ctx* c, int k, const SeqSet* s, int stranded, int p, int bbits, Partition** out);
void free_partition(Partition* P);
// partition of one rank's reads as bucket REGIONS (possibly with gaps) for the fused compaction + exchange of multi.cu
struct PartRegions {
    Ctx* ctx = nullptr;
    int k = 0, p = 0, bbits = 0, rec_words = 2, direct = 0;
    u64 n_input = 0, n_rec = 0;
    const u64* rec = nullptr;     // records, bucket b = [start[b], start[b] + cnt[b])
    const u64* start = nullptr;   // 2^bbits (+1) record offsets
    const u32* cnt = nullptr;     // 2^bbits
    void* holder = nullptr;       // owner of the device buffers
};
#ifndef DBG_MAX_RANKS
#define DBG_MAX_RANKS 8
#endif
struct ScatterDst { u64* base[DBG_MAX_RANKS]; u64 bound[DBG_MAX_RANKS + 1]; int P, me; };
int partition_regions_dev(Ctx* c, int k, const SeqSet* s, int stranded, int p, int bbits, PartRegions** out);
void free_part_regions(PartRegions* R);
int bucket_totals_dev(Ctx* c, const u32* d_all_cnt, int P, int me, u32 nb, u32* d_tot, u32* d_pre);
int local_buckets_dev(Ctx* c, const u64* d_goff, const u32* d_tot, u32 b0, u32 n, u64* d_off, u32* d_cnt);
int scatter_buckets_dev(Ctx* c, const PartRegions* R, const u64* d_goff, const u32* d_pre, const ScatterDst& D, u64* d_sums);
int filter_from_bucketed_dev(Ctx* c, int k, u64* d_records, u64 n_records, const u64* d_off, const u32* d_cnt, u32 n_local,
                             u64 n_kmers_local, u64 n_input_total, u32 min_obs, int stranded, int report_all, Table** out);
int graph_edges_dev(Ctx* c, const Graph* g, u32* h_target, u8* h_flags);
int graph_fix_exts_dev(Ctx* c, Graph* g, const u8* h_valid_nodes);
int graph_is_compressed_dev(Ctx* c, const Graph* g, int scmap, long long* pair_out);
int graph_combine_dev(Ctx* c, const Graph* const* gs, u32 n, Graph** out);
int compress_graph_dev(Ctx* c, const Graph* g, int stranded, int reduce_op, const u64* h_censor, u64 n_censor, Graph** out);
int remove_censored_exts_dev(Ctx* c, Table* t, int stranded, int sharded);
int table_prefix_hist_dev(Ctx* c, const Table* t, int bits, u32* d_hist);
int graph_from_device_dev(Ctx* c, int k, int stranded, u64 n_nodes, u64 n_bases, const u64* d_words, const u64* d_start,
                          const u32* d_length, const u8* d_exts, const u16* d_data, Graph** out);
int msp_sequence_dev(Ctx* c, int k, int p, const SeqSet* s, int rc, const u32* h_perm, u64 cap, u64* n_out, u32* h_seq, u32* h_start,
                     u32* h_len, u32* h_bucket, u8* h_exts);
int msp_kmer_buckets_dev(Ctx* c, int k, int p, const SeqSet* s, int stranded, u32* h_out, u64 n_out);
int filter_from_records_dev(Ctx* c, int k, const u64* d_records, u64 n_records, const u32* h_counts, u32 n_src, u32 n_local,
                            u64 n_input_total, u32 min_obs, int stranded, int report_all, Table** out);
int synth_reads_dev(Ctx* c, u64 R, u64 seed, u32 err_thr, SeqSet** out);
void free_seqset(SeqSet* s);
void free_table(Table* t);
void free_graph(Graph* g);
int count_input_kmers_dev(Ctx* c, int k, const SeqSet* s, u64* N_out);

// ---- sharded compression over a bucket-sharded table (shard_compress.cu; orchestrated by multi.cu) ----------------
struct ShardCfg { int P, me, p, bbits, stranded, scmap; };
// every rank's peer-mapped window: walk records (one uint4 per k-mer) and a copy of the shard's k-mers
struct RecPeers { const uint4* rec[DBG_MAX_RANKS]; const u64* klo[DBG_MAX_RANKS]; const u64* khi[DBG_MAX_RANKS]; };
struct SegOff { u64 off[DBG_MAX_RANKS + 1]; };
// neighbour queries of one rank, grouped by owning rank: segment r = [off[r], off[r + 1]) of msg / src
struct MsQueries {
    DBuf<u64> ctr, lut;
    DBuf<u32> lut_cnt, src;
    DBuf<unsigned char> msg;
    int lut_shift = 0;
    u64 n_total = 0, n_dst[DBG_MAX_RANKS] = {0}, off[DBG_MAX_RANKS + 1] = {0};
};
int ms_links_dev(Ctx* c, const Table* t, ShardCfg cfg, uint4* d_rec, u64* d_klo, u64* d_khi, MsQueries* q);
u32 ms_query_bytes(int k);
u32 ms_path_bytes(int k);
int ms_resolve_dev(Ctx* c, const Table* t, const MsQueries* q, int stranded, const void* d_queries, u64 nq, uint2* d_reply);
int ms_apply_dev(Ctx* c, const Table* t, ShardCfg cfg, const MsQueries* q, const uint2* d_reply, uint4* d_rec);
int ms_link_error(Ctx* c, const MsQueries* q, u32* code);
// outputs of one walker round: destinations 0 .. P-1 = per-rank outboxes, P = the rank's own list (emit entries / finished nodes)
struct WalkOut {
    void* box[DBG_MAX_RANKS + 1];
    u64 cap[DBG_MAX_RANKS + 1];
    u64* cursor;   // [0 .. P] items appended per destination, [P + 1] k-mers covered by the emit entries, [P + 2] overflow flag
    int P;
};
int ms_count_ends_dev(Ctx* c, const uint4* d_rec, u64 n, u64* d_out);
u32 ms_fitem_bytes(int k);
int ms_fwalk_start_dev(Ctx* c, int k, const uint4* rec, const u64* klo, const u64* khi, int me, u64 n, u32 lmax, int reduce_op, const WalkOut& out);
int ms_fwalk_continue_dev(Ctx* c, int k, const uint4* rec, const u64* klo, const u64* khi, int me, const void* inbox, u64 n_in, u32 lmax, int reduce_op,
                          const WalkOut& out);
int ms_scatter_nodes_dev(Ctx* c, int k, const void* nmsg, u64 m, int P, int bits, const u64* cuts, const u64* seg_off, void* out);
int ms_unpack_nodes_dev(Ctx* c, int k, const void* in, u64 m, u64* k_lo, u64* k_hi, u32* idx);
int ms_node_len_dev(Ctx* c, int k, const void* msgs, const u32* idx, u64 m, u64* node_len, u32* out_length);
int ms_emit_dev(Ctx* c, int k, const RecPeers& peers, const void* msgs, const u32* idx, const u64* node_start, u64 m, int reduce_op,
                u64* words, u8* out_exts, u16* out_data);
int ms_key_hist_dev(Ctx* c, int k, const u64* k_lo, const u64* k_hi, u64 m, int bits, u32* d_hist);

}  // namespace dbg
