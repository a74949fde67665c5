// Multi-GPU form of the path INSIDE the library (SURVEY §8b "Threading", §8e): one rank per GPU, either one process
// per GPU (dbg_comm_create: the caller hands every rank the same NCCL unique id, e.g. over its own rendezvous) or one
// process driving several GPUs with one host thread per rank (dbg_multi_create).  No torch, no Python in the data path.
//
//   filter_kmers   shards by MSP bucket — the reference's own sharded flow (src/test.rs:433-456: msp_sequence -> per-shard
//                  filter_kmers): every rank cuts ITS reads into per-bucket regions of super-k-mer records with the same plan;
//                  the per-bucket counts are all-gathered, so every sender knows the final position of each of its buckets in
//                  the owner's bucket-contiguous receive window, and ONE kernel stores the records there over peer memory
//                  (NVLink; compaction + exchange fused, no send buffer, no merge on the receiver); the owner counts its buckets.
//   compress_kmers the table STAYS sharded (shard_compress.cu): remote neighbour queries by a small all-to-all; unitig walkers
//                  never read another rank's records — a walker whose chain continues elsewhere is shipped there (bulk all-to-all
//                  per round) with the node it has collected so far; finished nodes go to the rank that owns their seed's key
//                  range; every rank ends with a contiguous run of nodes of the complete BaseGraph (runs concatenated in rank
//                  order = the single-GPU graph, bit for bit).
//
// Transports: NCCL (+ CUDA IPC for the peer windows) is the product path; a host-staged "local" transport connects ranks
// living in one process without NCCL — it lets several ranks share ONE device, which is how the multi-rank logic is
// tested on a single-GPU box (NCCL refuses two ranks on one device).
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

#include "common.cuh"

namespace dbg {

// ---- NCCL, loaded at run time (the single-GPU library must load on machines without it) --------------------------
struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi* nccl_api() {
    static std::mutex mu;
    static NcclApi api;
    static bool tried = false;
    std::lock_guard<std::mutex> g(mu);
    if (!tried) {
        tried = true;
        // a process that already mapped an NCCL (e.g. the one bundled with torch) gets that copy: same soname
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (h) {
            api.h = h;
#define LOADSYM(n) api.n = reinterpret_cast<decltype(api.n)>(dlsym(h, "nccl" #n))
            LOADSYM(GetUniqueId); LOADSYM(CommInitRank); LOADSYM(CommDestroy); LOADSYM(AllReduce); LOADSYM(AllGather);
            LOADSYM(Send); LOADSYM(Recv); LOADSYM(GroupStart); LOADSYM(GroupEnd); LOADSYM(GetErrorString);
#undef LOADSYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather || !api.Send || !api.Recv ||
                !api.GroupStart || !api.GroupEnd)
                api.h = nullptr;
        }
    }
    return api.h ? &api : nullptr;
}

#define NC(c, call)                                                                                                      \
    do {                                                                                                                 \
        ncclResult_t _r = (call);                                                                                        \
        if (_r != ncclSuccess)                                                                                           \
            DBG_SET_ERR(c, DBG_E_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, api->GetErrorString ? api->GetErrorString(_r) : "nccl error"); \
    } while (0)

// ---- transport interface -----------------------------------------------------------------------------------------
struct Transport {
    int rank = 0, size = 1;
    Ctx* ctx = nullptr;
    // peer window: every rank's walk-record array, visible from every rank
    char* win = nullptr;
    u64 win_cap = 0, win_gen = 0;
    void* peer_ptr[DBG_MAX_RANKS] = {nullptr};
    virtual ~Transport() {}
    virtual int all_reduce_sum(void* d, u64 n, bool is64) = 0;                                  // in place, device, stream-ordered
    virtual int all_gather(const void* d_send, void* d_recv, u64 bytes) = 0;                    // device buffers
    virtual int all_to_all_v(const void* d_send, const u64* soff, const u64* scnt, void* d_recv, const u64* roff, const u64* rcnt) = 0;   // bytes
    virtual int exchange_windows() = 0;                                                         // fills peer_ptr[] (collective)
    virtual const char* name() const = 0;

    // host-value conveniences (blocking)
    int all_reduce_host(u64* v, int n) {
        DBuf<u64> d;
        TRY(d.alloc_pool(ctx, n));
        for (int i = 0; i < n; i++) ctx->h_scratch[i] = v[i];
        CU(ctx, cudaMemcpyAsync(d.p, ctx->h_scratch, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
        TRY(all_reduce_sum(d.p, n, true));
        return read_u64(ctx, d.p, v, n);
    }
    int all_gather_host(const u64* v, int n, u64* out /* size * n */) {
        DBuf<u64> s, r;
        TRY(s.alloc_pool(ctx, n)); TRY(r.alloc_pool(ctx, (u64)n * size));
        for (int i = 0; i < n; i++) ctx->h_scratch[i] = v[i];
        CU(ctx, cudaMemcpyAsync(s.p, ctx->h_scratch, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
        TRY(all_gather(s.p, r.p, 8ull * n));
        return read_u64(ctx, r.p, out, n * size);
    }
    int barrier() { u64 one = 1; return all_reduce_host(&one, 1); }
    // (re)allocate the own window (plain cudaMalloc: exportable through CUDA IPC) and learn everybody's
    int ensure_window(u64 bytes) {
        if (bytes > win_cap) {
            CU(ctx, spin_sync(ctx->stream));
            if (win) CU(ctx, cudaFree(win));
            win = nullptr; win_cap = 0;
            const u64 want = bytes + bytes / 4 + (1ull << 20);
            CU(ctx, cudaMalloc((void**)&win, want));
            win_cap = want;
            win_gen++;
        }
        return exchange_windows();
    }
    void free_window() { if (win) cudaFree(win); win = nullptr; win_cap = 0; }
};

// ---- NCCL transport: one communicator per rank; peer windows through CUDA IPC (or raw pointers inside one process) ----
struct WinInfo { u64 pid, device, ptr, gen; cudaIpcMemHandle_t handle; };
struct NcclTransport : Transport {
    NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    WinInfo seen[DBG_MAX_RANKS];
    void* opened[DBG_MAX_RANKS] = {nullptr};
    const char* name() const override { return "nccl"; }
    ~NcclTransport() override {
        for (int r = 0; r < DBG_MAX_RANKS; r++) if (opened[r]) cudaIpcCloseMemHandle(opened[r]);
        if (comm && api) api->CommDestroy(comm);
        free_window();
    }
    int all_reduce_sum(void* d, u64 n, bool is64) override {
        NC(ctx, api->AllReduce(d, d, n, is64 ? ncclUint64 : ncclUint32, ncclSum, comm, ctx->stream));
        return DBG_OK;
    }
    int all_gather(const void* d_send, void* d_recv, u64 bytes) override {
        NC(ctx, api->AllGather(d_send, d_recv, bytes, ncclUint8, comm, ctx->stream));
        return DBG_OK;
    }
    int all_to_all_v(const void* d_send, const u64* soff, const u64* scnt, void* d_recv, const u64* roff, const u64* rcnt) override {
        if (scnt[rank] != rcnt[rank]) DBG_SET_ERR(ctx, DBG_E_INTERNAL, "all_to_all_v: own segment sizes differ");
        if (scnt[rank]) CU(ctx, cudaMemcpyAsync((char*)d_recv + roff[rank], (const char*)d_send + soff[rank], scnt[rank], cudaMemcpyDeviceToDevice, ctx->stream));
        NC(ctx, api->GroupStart());
        for (int r = 0; r < size; r++) {
            if (r == rank) continue;
            if (scnt[r]) NC(ctx, api->Send((const char*)d_send + soff[r], scnt[r], ncclUint8, r, comm, ctx->stream));
            if (rcnt[r]) NC(ctx, api->Recv((char*)d_recv + roff[r], rcnt[r], ncclUint8, r, comm, ctx->stream));
        }
        NC(ctx, api->GroupEnd());
        return DBG_OK;
    }
    int exchange_windows() override {
        WinInfo mine;
        memset(&mine, 0, sizeof(mine));
        mine.pid = (u64)getpid(); mine.device = (u64)ctx->device; mine.ptr = (u64)(uintptr_t)win; mine.gen = win_gen;
        if (win) CU(ctx, cudaIpcGetMemHandle(&mine.handle, win));
        DBuf<unsigned char> s, r;
        TRY(s.alloc_pool(ctx, sizeof(WinInfo))); TRY(r.alloc_pool(ctx, sizeof(WinInfo) * size));
        std::vector<WinInfo> all(size);
        CU(ctx, cudaMemcpyAsync(s.p, &mine, sizeof(WinInfo), cudaMemcpyHostToDevice, ctx->stream));
        TRY(all_gather(s.p, r.p, sizeof(WinInfo)));
        CU(ctx, cudaMemcpyAsync(all.data(), r.p, sizeof(WinInfo) * size, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, spin_sync(ctx->stream));
        for (int p = 0; p < size; p++) {
            if (p == rank) { peer_ptr[p] = win; continue; }
            const WinInfo& w = all[p];
            if (w.pid == mine.pid) {   // same process: plain peer access
                if ((int)w.device != ctx->device) {
                    cudaError_t e = cudaDeviceEnablePeerAccess((int)w.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) DBG_SET_ERR(ctx, DBG_E_CUDA, "peer access %d -> %d: %s", ctx->device, (int)w.device, cudaGetErrorString(e));
                    cudaGetLastError();
                }
                peer_ptr[p] = (void*)(uintptr_t)w.ptr;
                continue;
            }
            if (opened[p] && seen[p].gen == w.gen && seen[p].ptr == w.ptr && seen[p].pid == w.pid) { peer_ptr[p] = opened[p]; continue; }
            if (opened[p]) { cudaIpcCloseMemHandle(opened[p]); opened[p] = nullptr; }
            if (!w.ptr) { peer_ptr[p] = nullptr; seen[p] = w; continue; }
            void* q = nullptr;
            CU(ctx, cudaIpcOpenMemHandle(&q, w.handle, cudaIpcMemLazyEnablePeerAccess));
            opened[p] = q; seen[p] = w; peer_ptr[p] = q;
        }
        return DBG_OK;
    }
};

// ---- local transport: ranks are host threads of one process; data moves by cudaMemcpy between their buffers ----
struct LocalHub {
    int size = 0;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    u64 gen = 0;
    const void* ptr[DBG_MAX_RANKS];
    const u64* soff[DBG_MAX_RANKS];
    const u64* scnt[DBG_MAX_RANKS];
    int device[DBG_MAX_RANKS];
    std::vector<u64> host[DBG_MAX_RANKS];
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const u64 g = gen;
        if (++waiting == size) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
struct LocalTransport : Transport {
    LocalHub* hub = nullptr;
    const char* name() const override { return "local"; }
    ~LocalTransport() override { free_window(); }
    int all_reduce_sum(void* d, u64 n, bool is64) override {
        std::vector<u64>& mine = hub->host[rank];
        mine.assign(n, 0);
        std::vector<u32> tmp32;
        if (is64) { CU(ctx, cudaMemcpyAsync(mine.data(), d, 8 * n, cudaMemcpyDeviceToHost, ctx->stream)); }
        else { tmp32.resize(n); CU(ctx, cudaMemcpyAsync(tmp32.data(), d, 4 * n, cudaMemcpyDeviceToHost, ctx->stream)); }
        CU(ctx, spin_sync(ctx->stream));
        if (!is64) for (u64 i = 0; i < n; i++) mine[i] = tmp32[i];
        hub->barrier();
        std::vector<u64> sum(n, 0);
        for (int r = 0; r < size; r++) for (u64 i = 0; i < n; i++) sum[i] += hub->host[r][i];
        if (is64) { CU(ctx, cudaMemcpyAsync(d, sum.data(), 8 * n, cudaMemcpyHostToDevice, ctx->stream)); }
        else { for (u64 i = 0; i < n; i++) tmp32[i] = (u32)sum[i]; CU(ctx, cudaMemcpyAsync(d, tmp32.data(), 4 * n, cudaMemcpyHostToDevice, ctx->stream)); }
        CU(ctx, spin_sync(ctx->stream));
        hub->barrier();
        return DBG_OK;
    }
    int all_gather(const void* d_send, void* d_recv, u64 bytes) override {
        CU(ctx, spin_sync(ctx->stream));
        hub->ptr[rank] = d_send;
        hub->barrier();
        for (int r = 0; r < size; r++) CU(ctx, cudaMemcpyAsync((char*)d_recv + r * bytes, hub->ptr[r], bytes, cudaMemcpyDefault, ctx->stream));
        CU(ctx, spin_sync(ctx->stream));
        hub->barrier();
        return DBG_OK;
    }
    int all_to_all_v(const void* d_send, const u64* soff, const u64* scnt, void* d_recv, const u64* roff, const u64* rcnt) override {
        CU(ctx, spin_sync(ctx->stream));   // the send buffer is complete before anybody pulls from it
        hub->ptr[rank] = d_send; hub->soff[rank] = soff; hub->scnt[rank] = scnt;
        hub->barrier();
        int bad = 0;
        for (int r = 0; r < size; r++) {
            const u64 nb = hub->scnt[r][rank];
            if (nb != rcnt[r]) { bad = 1; continue; }
            if (nb) CU(ctx, cudaMemcpyAsync((char*)d_recv + roff[r], (const char*)hub->ptr[r] + hub->soff[r][rank], nb, cudaMemcpyDefault, ctx->stream));
        }
        CU(ctx, spin_sync(ctx->stream));
        hub->barrier();
        if (bad) DBG_SET_ERR(ctx, DBG_E_INTERNAL, "all_to_all_v: send / receive counts disagree");
        return DBG_OK;
    }
    int exchange_windows() override {
        hub->ptr[rank] = win; hub->device[rank] = ctx->device;
        hub->barrier();
        for (int p = 0; p < size; p++) {
            if (hub->device[p] != ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(hub->device[p], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) DBG_SET_ERR(ctx, DBG_E_CUDA, "peer access %d -> %d: %s", ctx->device, hub->device[p], cudaGetErrorString(e));
                cudaGetLastError();
            }
            peer_ptr[p] = const_cast<void*>(hub->ptr[p]);
        }
        hub->barrier();
        return DBG_OK;
    }
};

// ---- host-side planning helpers (pure functions: exported for the CPU tests) ----
// rank r owns the buckets [bounds[r], bounds[r + 1]); owner(b) = (b * P) >> bits for a power-of-two bucket count
static void owner_bounds(u64 n_buckets, int P, u64* bounds) {
    for (int r = 0; r <= P; r++) bounds[r] = ((u64)r * n_buckets + P - 1) / P;
}
// cut a histogram into P ranges of ~equal mass: P + 1 bin indices, first 0, last nbins
static void quantile_cuts(const u64* hist, u64 nbins, int P, u64* cuts) {
    u64 total = 0;
    for (u64 i = 0; i < nbins; i++) total += hist[i];
    cuts[0] = 0;
    u64 acc = 0, b = 0;
    for (int r = 1; r < P; r++) {
        const u64 target = total * (u64)r / (u64)P;
        while (b < nbins && acc + hist[b] <= target) { acc += hist[b]; b++; }
        cuts[r] = b;
    }
    cuts[P] = nbins;
    for (int r = 1; r <= P; r++) if (cuts[r] < cuts[r - 1]) cuts[r] = cuts[r - 1];
}

__global__ void iota_kernel(u32* out, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)i;
}

struct EvTimer {
    cudaEvent_t ev[16];
    int n = 0;
    EvTimer() { for (auto& e : ev) cudaEventCreate(&e); }
    ~EvTimer() { for (auto& e : ev) cudaEventDestroy(e); }
    void mark(cudaStream_t st) { if (n < 16) cudaEventRecord(ev[n++], st); }
    float ms(int a, int b) { float t = 0; if (a < n && b < n) cudaEventElapsedTime(&t, ev[a], ev[b]); return t; }
};

// Long unitigs / cycles: every rank gathers the whole table and runs the single-GPU compression (complete graph everywhere).
static int fallback_replicated(Transport* T, const Table* shard, int stranded, int reduce_op, Graph** out) {
    Ctx* c = T->ctx;
    const int P = T->size;
    const bool two = shard->k > 32;
    u64 mine = shard->n;
    std::vector<u64> all(P);
    TRY(T->all_gather_host(&mine, 1, all.data()));
    u64 total = 0;
    std::vector<u64> off(P + 1, 0);
    for (int r = 0; r < P; r++) { off[r] = total; total += all[r]; }
    DBuf<u64> lo, hi;
    DBuf<u8> ex;
    DBuf<u16> cn;
    TRY(lo.alloc_pool(c, total)); TRY(ex.alloc_pool(c, total)); TRY(cn.alloc_pool(c, total));
    if (two) TRY(hi.alloc_pool(c, total));
    auto gather = [&](const void* src, void* dst, u64 es) -> int {
        std::vector<u64> so(P, 0), sc(P, mine * es), ro(P), rc(P);
        for (int r = 0; r < P; r++) { ro[r] = off[r] * es; rc[r] = all[r] * es; }
        return T->all_to_all_v(src, so.data(), sc.data(), dst, ro.data(), rc.data());
    };
    TRY(gather(shard->lo, lo.p, 8));
    if (two) TRY(gather(shard->hi, hi.p, 8));
    TRY(gather(shard->exts, ex.p, 1));
    TRY(gather(shard->counts, cn.p, 2));
    dbg_kmer_table* full = nullptr;
    int rc = dbg_table_from_device(reinterpret_cast<dbg_ctx*>(c), shard->k, total, lo.p, two ? hi.p : nullptr, ex.p, cn.p, &full);
    if (rc != DBG_OK) return rc;
    rc = compress_dev(c, &full->t, stranded, reduce_op, out);
    free_table(&full->t);
    return rc;
}

static int multi_reads_to_graph(Transport* T, int k, const SeqSet* s, u32 min_obs, int stranded, int reduce_op, dbg_multi_info* info,
                                Graph** out) {
    Ctx* c = T->ctx;
    cudaStream_t st = c->stream;
    const int P = T->size, me = T->rank;
    *out = nullptr;
    dbg_multi_info I;
    memset(&I, 0, sizeof(I));
    I.n_ranks = P; I.rank = me;
    if (k < 4 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [4,64]", k);
    if (reduce_op < 0 || reduce_op > DBG_REDUCE_SCMAP) DBG_SET_ERR(c, DBG_E_BADARG, "unknown reduce_op %d", reduce_op);
    CU(c, cudaSetDevice(c->device));
    EvTimer tm;
    tm.mark(st);   // 0
    // ---- plan from the TOTAL number of input k-mers ----
    u64 n_local_in = 0;
    TRY(count_input_kmers_dev(c, k, s, &n_local_in));
    u64 tot = n_local_in;
    TRY(T->all_reduce_host(&tot, 1));
    I.n_input_total = tot;
    int p = 0, bbits = 0;
    plan_filter(c, k, tot, &p, &bbits);
    while ((1 << bbits) < P) bbits++;
    I.msp_p = p; I.bucket_bits = bbits;
    const u64 NB = 1ull << bbits;
    u64 bounds[DBG_MAX_RANKS + 1];
    owner_bounds(NB, P, bounds);
    const u64 n_own = bounds[me + 1] - bounds[me];
    // ---- partition the own reads into per-bucket regions of super-k-mer records (no compaction here) ----
    PartRegions* part = nullptr;
    TRY(partition_regions_dev(c, k, s, stranded, p, bbits, &part));
    struct PartGuard { PartRegions* r; ~PartGuard() { if (r) free_part_regions(r); } } pg{part};
    tm.mark(st);   // 1
    const u32 rec_bytes = (u32)part->rec_words * 8;
    u64 so[DBG_MAX_RANKS], sc[DBG_MAX_RANKS], ro[DBG_MAX_RANKS], rc[DBG_MAX_RANKS];   // byte offsets / counts of the later all-to-alls
    // ---- the path's one big exchange, fused with the compaction: every rank learns every rank's per-bucket record counts (one small
    // all-gather), so each sender knows the FINAL position of each of its buckets inside the owner's bucket-contiguous receive
    // window, and one kernel copies the records there over peer memory (NVLink stores).  No send buffer, no gap-closing pass, no
    // per-source merge on the receiver; the closing all-reduce is the barrier and carries the k-mer totals. ----
    DBuf<u32> d_allcnt, d_btot, d_pre, d_cnt_loc;
    DBuf<u64> d_goff, d_sums, d_gb, d_off_loc;
    TRY(d_allcnt.alloc_pool(c, (u64)P * NB)); TRY(d_btot.alloc_pool(c, NB)); TRY(d_pre.alloc_pool(c, NB)); TRY(d_goff.alloc_pool(c, NB + 1));
    TRY(d_sums.alloc_pool(c, 2 * DBG_MAX_RANKS)); TRY(d_gb.alloc_pool(c, DBG_MAX_RANKS + 1));
    TRY(d_off_loc.alloc_pool(c, n_own + 1)); TRY(d_cnt_loc.alloc_pool(c, n_own ? n_own : 1));
    TRY(T->all_gather(part->cnt, d_allcnt.p, NB * 4));
    TRY(bucket_totals_dev(c, d_allcnt.p, P, me, (u32)NB, d_btot.p, d_pre.p));
    TRY(exclusive_scan_u32_to_u64(c, d_btot.p, d_goff.p, NB, d_goff.p + NB));
    for (int r = 0; r <= P; r++) CU(c, cudaMemcpyAsync(d_gb.p + r, d_goff.p + bounds[r], 8, cudaMemcpyDeviceToDevice, st));
    u64 gb[DBG_MAX_RANKS + 1];
    TRY(read_u64(c, d_gb.p, gb, P + 1));
    const u64 n_recv = gb[me + 1] - gb[me];
    TRY(T->ensure_window((n_recv ? n_recv : 1) * rec_bytes));
    ScatterDst D;
    D.P = P; D.me = me;
    for (int r = 0; r < DBG_MAX_RANKS; r++) { D.base[r] = r < P ? reinterpret_cast<u64*>(T->peer_ptr[r]) : nullptr; D.bound[r] = r <= P ? bounds[r] : NB; }
    D.bound[DBG_MAX_RANKS] = NB;
    for (int r = P; r <= DBG_MAX_RANKS; r++) D.bound[r] = NB;
    CU(c, cudaMemsetAsync(d_sums.p, 0, 8 * 2 * DBG_MAX_RANKS, st));
    EvTimer tx;   // the scatter kernel alone (DBG_MULTI_TRACE)
    tx.mark(st);
    TRY(scatter_buckets_dev(c, part, d_goff.p, d_pre.p, D, d_sums.p));
    tx.mark(st);
    u64 sent[2 * DBG_MAX_RANKS];
    TRY(read_u64(c, d_sums.p, sent, 2 * P));
    for (int r = 0; r < P; r++) if (r != me) I.exchange_bytes_sent += sent[P + r] * rec_bytes;
    free_part_regions(part);   // (the scatter kernel has completed: read_u64 waited for it)
    pg.r = nullptr;
    TRY(T->all_reduce_sum(d_sums.p, P, true));   // every rank's stores have landed once this completes; sums[r] = k-mer occurrences rank r holds
    u64 kin[DBG_MAX_RANKS];
    TRY(read_u64(c, d_sums.p, kin, P));
    TRY(local_buckets_dev(c, d_goff.p, d_btot.p, (u32)bounds[me], (u32)n_own, d_off_loc.p, d_cnt_loc.p));
    tm.mark(st);   // 2
    // ---- count the owned buckets, sort the valid k-mers: this rank's shard of the table ----
    Table* shard = nullptr;
    TRY(filter_from_bucketed_dev(c, k, reinterpret_cast<u64*>(T->win), n_recv, d_off_loc.p, d_cnt_loc.p, (u32)n_own, kin[me], tot, min_obs, stranded,
                                 0, &shard));
    d_allcnt.release(); d_btot.release(); d_pre.release(); d_goff.release(); d_off_loc.release(); d_cnt_loc.release();
    tm.mark(st);   // 3
    struct TableGuard { Table* t; ~TableGuard() { if (t) free_table(t); } } tg{shard};
    const u64 V = shard->n;
    I.n_valid_local = V;
    // ================= compression over the sharded table =================
    TRY(arena_begin(c));
    // window of this rank: [walk records: V x 16 B][k-mers lo: V x 8 B][k-mers hi: V x 8 B (K > 32)], each part 256-byte aligned
    const int W = k <= 32 ? 1 : 2;
    std::vector<u64> Vs(P);
    TRY(T->all_gather_host(&V, 1, Vs.data()));
    auto al = [](u64 x) { return (x + 255) & ~255ull; };
    TRY(T->ensure_window(al((V ? V : 1) * 16) + al((V ? V : 1) * 8) * W));
    RecPeers peers;
    for (int r = 0; r < DBG_MAX_RANKS; r++) {
        peers.rec[r] = nullptr; peers.klo[r] = nullptr; peers.khi[r] = nullptr;
        if (r >= P || !T->peer_ptr[r]) continue;
        char* base = reinterpret_cast<char*>(T->peer_ptr[r]);
        const u64 vr = Vs[r] ? Vs[r] : 1;
        peers.rec[r] = reinterpret_cast<const uint4*>(base);
        peers.klo[r] = reinterpret_cast<const u64*>(base + al(vr * 16));
        peers.khi[r] = W == 2 ? reinterpret_cast<const u64*>(base + al(vr * 16) + al(vr * 8)) : nullptr;
    }
    uint4* my_rec = reinterpret_cast<uint4*>(T->win);
    u64* my_klo = const_cast<u64*>(peers.klo[me]);
    u64* my_khi = const_cast<u64*>(peers.khi[me]);
    ShardCfg cfg{P, me, p, bbits, stranded, reduce_op == DBG_REDUCE_SCMAP};
    MsQueries q;
    TRY(ms_links_dev(c, shard, cfg, my_rec, my_klo, my_khi, &q));
    I.n_queries_sent = q.n_total;
    // query counts: M[s][d] = queries rank s has for rank d
    std::vector<u64> M((u64)P * P);
    TRY(T->all_gather_host(q.n_dst, P, M.data()));
    const u32 qb = ms_query_bytes(k);
    u64 n_qin = 0;
    u64 qin_off[DBG_MAX_RANKS + 1];
    for (int r = 0; r < P; r++) { qin_off[r] = n_qin; n_qin += M[(u64)r * P + me]; }
    qin_off[P] = n_qin;
    DBuf<unsigned char> qin;
    DBuf<uint2> rep_out, rep_in;
    TRY(qin.alloc_pool(c, (n_qin ? n_qin : 1) * qb));
    TRY(rep_out.alloc_pool(c, n_qin ? n_qin : 1));
    TRY(rep_in.alloc_pool(c, q.n_total ? q.n_total : 1));
    for (int r = 0; r < P; r++) { so[r] = q.off[r] * qb; sc[r] = q.n_dst[r] * qb; ro[r] = qin_off[r] * qb; rc[r] = M[(u64)r * P + me] * qb; }
    TRY(T->all_to_all_v(q.msg.p, so, sc, qin.p, ro, rc));
    TRY(ms_resolve_dev(c, shard, &q, stranded, qin.p, n_qin, rep_out.p));
    for (int r = 0; r < P; r++) { so[r] = qin_off[r] * 8; sc[r] = M[(u64)r * P + me] * 8; ro[r] = q.off[r] * 8; rc[r] = q.n_dst[r] * 8; }
    TRY(T->all_to_all_v(rep_out.p, so, sc, rep_in.p, ro, rc));
    TRY(ms_apply_dev(c, shard, cfg, &q, rep_in.p, my_rec));
    u32 lerr = 0;
    TRY(ms_link_error(c, &q, &lerr));
    tm.mark(st);   // 4
    // ---- discover + collect in ONE pass of walker rounds (shard_compress.cu, fused walk): a walker never leaves its rank — when its
    // chain continues on another rank it is shipped there with the node collected so far; finished nodes stay where their right end lives ----
    const uint4* rec = my_rec;
    const u32 fib = ms_fitem_bytes(k), pb = ms_path_bytes(k);
    DBuf<u64> d_ends, cur, d_allc;
    TRY(d_ends.alloc_pool(c, 1)); TRY(cur.alloc_pool(c, DBG_MAX_RANKS + 4)); TRY(d_allc.alloc_pool(c, (u64)P * P));
    TRY(ms_count_ends_dev(c, rec, V, d_ends.p));
    u64 n_ends = 0;
    TRY(read_u64(c, d_ends.p, &n_ends));
    // capacities: a link side is crossed at most once per direction, so rank d receives at most (my link sides pointing at d) walkers from me
    u64 ocap[DBG_MAX_RANKS + 1], ooff[DBG_MAX_RANKS + 1], icap = 0, otot = 0;
    for (int r = 0; r < P; r++) { ocap[r] = q.n_dst[r] + 32; ooff[r] = otot; otot += ocap[r]; icap += M[(u64)r * P + me] + 32; }
    const u64 ecap = n_ends + 32;
    DBuf<unsigned char> obox, ibox, nmsg;
    TRY(obox.alloc_pool(c, otot * (u64)fib)); TRY(ibox.alloc_pool(c, icap * (u64)fib));
    TRY(nmsg.alloc_pool(c, ecap * pb));
    u64 n_paths = 0, n_cov = 0;
    std::vector<u64> allc((u64)P * P);
    auto run_rounds = [&]() -> int {
        // per-destination outboxes, the rank's own list of finished nodes as destination P
        WalkOut wo;
        for (int r = 0; r < P; r++) { wo.box[r] = obox.p + ooff[r] * fib; wo.cap[r] = ocap[r]; }
        for (int r = P; r <= DBG_MAX_RANKS; r++) { wo.box[r] = nullptr; wo.cap[r] = 0; }
        wo.box[P] = nmsg.p; wo.cap[P] = ecap;
        wo.cursor = cur.p; wo.P = P;
        CU(c, cudaMemsetAsync(cur.p, 0, 8 * (DBG_MAX_RANKS + 4), st));
        u64 n_in = 0;
        for (int round = 0; round < 4096; round++) {
            if (round == 0) TRY(ms_fwalk_start_dev(c, k, rec, my_klo, my_khi, me, V, 1024u, reduce_op, wo));
            else TRY(ms_fwalk_continue_dev(c, k, rec, my_klo, my_khi, me, ibox.p, n_in, 1024u, reduce_op, wo));
            // everybody learns everybody's outbox counts: the same matrix on every rank decides when the rounds end
            TRY(T->all_gather(cur.p, d_allc.p, 8ull * P));
            u64 own[3];
            CU(c, cudaMemcpyAsync(allc.data(), d_allc.p, 8ull * P * P, cudaMemcpyDeviceToHost, st));
            TRY(read_u64(c, cur.p + P, own, 3));
            if (own[2]) DBG_SET_ERR(c, DBG_E_INTERNAL, "walker buffer overflow");
            u64 inflight = 0;
            for (u64 i = 0; i < (u64)P * P; i++) inflight += allc[i];
            n_paths = own[0]; n_cov = own[1];
            if (!inflight) {
                if (getenv("DBG_MULTI_TRACE") && me == 0) fprintf(stderr, "[dbg multi] walker rounds: %d\n", round + 1);
                return DBG_OK;
            }
            u64 s_o[DBG_MAX_RANKS], s_c[DBG_MAX_RANKS], r_o[DBG_MAX_RANKS], r_c[DBG_MAX_RANKS];
            n_in = 0;
            for (int r = 0; r < P; r++) {
                s_o[r] = ooff[r] * fib; s_c[r] = allc[(u64)me * P + r] * fib;
                r_o[r] = n_in * fib; r_c[r] = allc[(u64)r * P + me] * fib; n_in += allc[(u64)r * P + me];
            }
            if (n_in > icap) DBG_SET_ERR(c, DBG_E_INTERNAL, "walker inbox overflow");
            TRY(T->all_to_all_v(obox.p, s_o, s_c, ibox.p, r_o, r_c));
            CU(c, cudaMemsetAsync(cur.p, 0, 8 * P, st));   // the outboxes are empty again; the own list and its counters carry on
        }
        DBG_SET_ERR(c, DBG_E_INTERNAL, "walker rounds did not terminate");
    };
    TRY(run_rounds());
    u64 red[5] = {V, n_cov, n_paths, lerr == 1 ? 1ull : 0ull, lerr == 2 ? 1ull : 0ull};
    TRY(T->all_reduce_host(red, 5));
    I.n_valid_total = red[0];
    if (red[3]) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "k-mer extension points at a k-mer with no extension back (src/compression.rs:428-434)");
    if (red[4]) DBG_SET_ERR(c, DBG_E_INCONSISTENT_EXTS, "k-mer extensions are not reciprocal");
    tm.mark(st);   // 5
    if (red[1] != red[0]) {
        // long unitigs or cycles: gather the table, compress replicated (complete graph on every rank)
        Graph* g = nullptr;
        TRY(fallback_replicated(T, shard, stranded, reduce_op, &g));
        TRY(T->barrier());
        tm.mark(st);
        I.replicated = 1;
        I.n_nodes_total = g->n_nodes; I.n_bases_total = g->n_bases; I.node0 = 0; I.base0 = 0;
        I.check_ok = I.n_bases_total == I.n_valid_total + I.n_nodes_total * (u64)(k - 1);
        I.ms_partition = tm.ms(0, 1); I.ms_exchange = tm.ms(1, 2); I.ms_count_sort = tm.ms(2, 3); I.ms_links = tm.ms(3, 4);
        I.ms_discover = tm.ms(4, 5); I.ms_layout = 0; I.ms_emit = tm.ms(5, 6); I.ms_total = tm.ms(0, 6);
        c->stats.n_valid = V;
        if (info) *info = I;
        *out = g;
        return DBG_OK;
    }
    EvTimer tl;   // finer timing of the layout stage (DBG_MULTI_TRACE)
    tl.mark(st);
    obox.release(); ibox.release();
    DBuf<u64> pk_lo, pk_hi;
    DBuf<u32> pk_idx;
    TRY(pk_lo.alloc_pool(c, n_paths ? n_paths : 1)); TRY(pk_idx.alloc_pool(c, n_paths ? n_paths : 1));
    if (W == 2) TRY(pk_hi.alloc_pool(c, n_paths ? n_paths : 1));
    TRY(ms_unpack_nodes_dev(c, k, nmsg.p, n_paths, pk_lo.p, pk_hi.p, pk_idx.p));
    // ---- finished nodes -> the rank owning their seed's key range (quantile cuts of the all-reduced seed histogram) ----
    const int hb = std::min(16, 2 * k);
    const u64 nbins = 1ull << hb;
    DBuf<u32> d_hist, d_ghist;
    TRY(d_hist.alloc_pool(c, nbins)); TRY(d_ghist.alloc_pool(c, nbins));
    TRY(ms_key_hist_dev(c, k, pk_lo.p, pk_hi.p, n_paths, hb, d_hist.p));
    CU(c, cudaMemcpyAsync(d_ghist.p, d_hist.p, nbins * 4, cudaMemcpyDeviceToDevice, st));
    TRY(T->all_reduce_sum(d_ghist.p, nbins, false));
    std::vector<u32> h_hist(nbins), h_ghist(nbins);
    CU(c, cudaMemcpyAsync(h_hist.data(), d_hist.p, nbins * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(h_ghist.data(), d_ghist.p, nbins * 4, cudaMemcpyDeviceToHost, st));
    TRY(sync(c));
    u64 cuts[DBG_MAX_RANKS + 1], pbound[DBG_MAX_RANKS + 1];
    {
        std::vector<u64> g64(nbins);
        for (u64 i = 0; i < nbins; i++) g64[i] = h_ghist[i];
        quantile_cuts(g64.data(), nbins, P, cuts);
        u64 acc = 0, b = 0;
        for (int r = 0; r <= P; r++) {   // destination r gets the seeds whose bin lies in [cuts[r], cuts[r + 1])
            while (b < cuts[r]) acc += h_hist[b++];
            pbound[r] = acc;
        }
    }
    DBuf<unsigned char> pmsg_out, pmsg_in;
    TRY(pmsg_out.alloc_pool(c, (n_paths ? n_paths : 1) * pb));
    TRY(ms_scatter_nodes_dev(c, k, nmsg.p, n_paths, P, hb, cuts, pbound, pmsg_out.p));
    nmsg.release();
    u64 psend[DBG_MAX_RANKS];
    for (int r = 0; r < P; r++) psend[r] = pbound[r + 1] - pbound[r];
    std::vector<u64> PM((u64)P * P);
    TRY(T->all_gather_host(psend, P, PM.data()));
    u64 m_own = 0;
    for (int r = 0; r < P; r++) { so[r] = pbound[r] * pb; sc[r] = psend[r] * pb; ro[r] = m_own * pb; rc[r] = PM[(u64)r * P + me] * pb; m_own += PM[(u64)r * P + me]; }
    TRY(pmsg_in.alloc_pool(c, (m_own ? m_own : 1) * pb));
    tl.mark(st);
    TRY(T->all_to_all_v(pmsg_out.p, so, sc, pmsg_in.p, ro, rc));
    tl.mark(st);
    // ---- own seed range: sort, node lengths, offsets ----
    DBuf<u64> nk_lo, nk_hi, nk_lo_b, nk_hi_b, node_len, node_start, d_tot;
    DBuf<u32> ni_a, ni_b, olen;
    const u64 mcap = m_own ? m_own : 1;
    TRY(nk_lo.alloc_pool(c, mcap)); TRY(nk_lo_b.alloc_pool(c, mcap)); TRY(ni_a.alloc_pool(c, mcap)); TRY(ni_b.alloc_pool(c, mcap));
    if (W == 2) { TRY(nk_hi.alloc_pool(c, mcap)); TRY(nk_hi_b.alloc_pool(c, mcap)); }
    TRY(node_len.alloc_pool(c, mcap)); TRY(node_start.alloc_pool(c, mcap)); TRY(olen.alloc_pool(c, mcap)); TRY(d_tot.alloc_pool(c, 1));
    TRY(ms_unpack_nodes_dev(c, k, pmsg_in.p, m_own, nk_lo.p, nk_hi.p, ni_a.p));
    u64 *sk_lo = nk_lo.p, *sk_hi = nk_hi.p;
    u32* sidx = ni_a.p;
    TRY(radix_sort_pairs(c, W, 2 * k, m_own, nk_lo.p, nk_hi.p, ni_a.p, nk_lo_b.p, nk_hi_b.p, ni_b.p, &sk_lo, &sk_hi, &sidx));
    tl.mark(st);
    TRY(ms_node_len_dev(c, k, pmsg_in.p, sidx, m_own, node_len.p, olen.p));
    u64 nb_own = 0;
    if (m_own) {
        TRY(exclusive_scan_u64(c, node_len.p, node_start.p, m_own, d_tot.p));
        TRY(read_u64(c, d_tot.p, &nb_own));
    }
    u64 mine2[2] = {m_own, nb_own};
    std::vector<u64> all2(2 * (u64)P);
    TRY(T->all_gather_host(mine2, 2, all2.data()));
    for (int r = 0; r < P; r++) {
        if (r < me) { I.node0 += all2[2 * r]; I.base0 += all2[2 * r + 1]; }
        I.n_nodes_total += all2[2 * r]; I.n_bases_total += all2[2 * r + 1];
    }
    tm.mark(st);   // 6
    // ---- emit the own run of nodes ----
    Graph* g = &(new dbg_graph())->g;
    g->ctx = c; g->k = k; g->stranded = stranded; g->n_nodes = m_own; g->n_bases = nb_own; g->n_words = (nb_own + 31) / 32;
    struct GraphGuard { Graph* g; ~GraphGuard() { if (g) free_graph(g); } } gg{g};
    if (m_own) {
        DBuf<u64> words;
        DBuf<u8> oexts;
        DBuf<u16> odata;
        TRY(words.alloc_pool(c, g->n_words + 3)); TRY(words.zero());
        TRY(oexts.alloc_pool(c, m_own)); TRY(odata.alloc_pool(c, m_own));
        TRY(ms_emit_dev(c, k, peers, pmsg_in.p, sidx, node_start.p, m_own, reduce_op, words.p, oexts.p, odata.p));
        TRY(sync(c));
        g->words = words.take(); g->start = node_start.take(); g->length = olen.take(); g->exts = oexts.take(); g->data = odata.take();
    }
    tm.mark(st);   // 7
    // nobody may overwrite (next call) or free its walk records while a peer still walks them
    TRY(T->barrier());
    tm.mark(st);   // 8
    TRY(sync(c));
    I.replicated = 0;
    I.check_ok = I.n_bases_total == I.n_valid_total + I.n_nodes_total * (u64)(k - 1);
    I.ms_partition = tm.ms(0, 1); I.ms_exchange = tm.ms(1, 2); I.ms_count_sort = tm.ms(2, 3); I.ms_links = tm.ms(3, 4);
    I.ms_discover = tm.ms(4, 5); I.ms_layout = tm.ms(5, 6); I.ms_emit = tm.ms(6, 7); I.ms_total = tm.ms(0, 8);
    if (getenv("DBG_MULTI_TRACE") && me == 0) {
        fprintf(stderr, "[dbg multi] exchange: scatter kernel %.2f ms for %.0f MB received here\n", tx.ms(0, 1), (double)n_recv * rec_bytes / 1e6);
        float t_col = 0, t_lay = 0;
        cudaEventElapsedTime(&t_col, tm.ev[5], tl.ev[0]);
        cudaEventElapsedTime(&t_lay, tl.ev[3], tm.ev[6]);
        fprintf(stderr, "[dbg multi] layout: (%.2f) | unpack+hist+cuts+scatter %.2f | a2a %.2f | unpack+sort %.2f | len+scan+gather %.2f ms; nodes here %llu\n",
                t_col, tl.ms(0, 1), tl.ms(1, 2), tl.ms(2, 3), t_lay, (unsigned long long)m_own);
    }
    c->stats.n_valid = V; c->stats.n_nodes = m_own; c->stats.n_bases = nb_own;
    c->stats.gpu_launches = c->launches;
    if (info) *info = I;
    gg.g = nullptr;
    *out = g;
    return DBG_OK;
}

}  // namespace dbg

using namespace dbg;

struct dbg_comm {
    Transport* T = nullptr;
};
struct dbg_multi {
    int n = 0;
    dbg_ctx* ctx[DBG_MAX_RANKS] = {nullptr};
    dbg_comm comm[DBG_MAX_RANKS];
    LocalHub hub;
    std::string err;
};

extern "C" {

int dbg_comm_unique_id(void* id_out) {
    NcclApi* api = nccl_api();
    if (!api || !id_out) return DBG_E_CUDA;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return DBG_E_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return DBG_OK;
}

int dbg_comm_create(dbg_ctx* ctx, int n_ranks, int rank, const void* unique_id, dbg_comm** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = &ctx->c;
    *out = nullptr;
    if (n_ranks < 1 || n_ranks > DBG_MAX_RANKS || rank < 0 || rank >= n_ranks || !unique_id) DBG_SET_ERR(c, DBG_E_BADARG, "need 1 <= n_ranks <= %d, 0 <= rank < n_ranks and a unique id", DBG_MAX_RANKS);
    NcclApi* api = nccl_api();
    if (!api) DBG_SET_ERR(c, DBG_E_CUDA, "libnccl.so.2 could not be loaded: the multi-GPU path needs NCCL (there is no fallback)");
    CU(c, cudaSetDevice(c->device));
    NcclTransport* T = new (std::nothrow) NcclTransport();
    if (!T) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    T->api = api; T->ctx = c; T->rank = rank; T->size = n_ranks;
    memset(T->seen, 0, sizeof(T->seen));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = api->CommInitRank(&T->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { T->comm = nullptr; delete T; DBG_SET_ERR(c, DBG_E_CUDA, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(r) : "error"); }
    dbg_comm* h = new dbg_comm();
    h->T = T;
    *out = h;
    return DBG_OK;
}

void dbg_comm_destroy(dbg_comm* comm) {
    if (!comm) return;
    if (comm->T) { cudaSetDevice(comm->T->ctx->device); cudaStreamSynchronize(comm->T->ctx->stream); delete comm->T; }
    delete comm;
}
int dbg_comm_rank(const dbg_comm* comm) { return comm && comm->T ? comm->T->rank : -1; }
int dbg_comm_size(const dbg_comm* comm) { return comm && comm->T ? comm->T->size : 0; }
const char* dbg_comm_transport(const dbg_comm* comm) { return comm && comm->T ? comm->T->name() : ""; }

int dbg_reads_to_graph_multi(dbg_comm* comm, int k, const dbg_seqset* seqs, uint32_t min_kmer_obs, int stranded, int reduce_op,
                             dbg_multi_info* info, dbg_graph** graph_out) {
    if (!comm || !comm->T || !seqs || !graph_out) return DBG_E_BADARG;
    Graph* g = nullptr;
    int rc = multi_reads_to_graph(comm->T, k, &seqs->s, min_kmer_obs, stranded, reduce_op, info, &g);
    *graph_out = rc == DBG_OK ? reinterpret_cast<dbg_graph*>(g) : nullptr;
    return rc;
}

int dbg_multi_create(const int* devices, int n, dbg_multi** out) {
    if (!devices || !out || n < 1 || n > DBG_MAX_RANKS) return DBG_E_BADARG;
    *out = nullptr;
    dbg_multi* m = new (std::nothrow) dbg_multi();
    if (!m) return DBG_E_OOM;
    m->n = n;
    m->hub.size = n;
    bool distinct = true;
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) if (devices[i] == devices[j]) distinct = false;
    for (int i = 0; i < n; i++) {
        int rc = dbg_ctx_create(devices[i], &m->ctx[i]);
        if (rc != DBG_OK) { for (int j = 0; j < i; j++) dbg_ctx_destroy(m->ctx[j]); delete m; return rc; }
    }
    const char* force = getenv("DBG_MULTI_TRANSPORT");
    NcclApi* api = nccl_api();
    const bool use_nccl = distinct && n > 1 && api && !(force && !strcmp(force, "local"));
    if (use_nccl) {
        ncclUniqueId id;
        if (api->GetUniqueId(&id) != ncclSuccess) { for (int i = 0; i < n; i++) dbg_ctx_destroy(m->ctx[i]); delete m; return DBG_E_CUDA; }
        std::vector<std::thread> th;
        std::vector<int> rcs(n, DBG_OK);
        std::vector<dbg_comm*> cm(n, nullptr);
        for (int i = 0; i < n; i++) th.emplace_back([&, i] { rcs[i] = dbg_comm_create(m->ctx[i], n, i, &id, &cm[i]); });
        for (auto& t : th) t.join();
        for (int i = 0; i < n; i++) {
            if (rcs[i] != DBG_OK) {
                for (int j = 0; j < n; j++) { if (cm[j]) dbg_comm_destroy(cm[j]); dbg_ctx_destroy(m->ctx[j]); }
                delete m;
                return rcs[i];
            }
            m->comm[i].T = cm[i]->T;
            cm[i]->T = nullptr;
            delete cm[i];
        }
    } else {
        for (int i = 0; i < n; i++) {
            LocalTransport* T = new LocalTransport();
            T->hub = &m->hub; T->ctx = &m->ctx[i]->c; T->rank = i; T->size = n;
            m->comm[i].T = T;
        }
    }
    *out = m;
    return DBG_OK;
}

void dbg_multi_destroy(dbg_multi* m) {
    if (!m) return;
    for (int i = 0; i < m->n; i++) {
        if (m->comm[i].T) { cudaSetDevice(m->ctx[i]->c.device); cudaStreamSynchronize(m->ctx[i]->c.stream); delete m->comm[i].T; m->comm[i].T = nullptr; }
    }
    for (int i = 0; i < m->n; i++) dbg_ctx_destroy(m->ctx[i]);
    delete m;
}
int dbg_multi_size(const dbg_multi* m) { return m ? m->n : 0; }
dbg_ctx* dbg_multi_ctx(dbg_multi* m, int rank) { return m && rank >= 0 && rank < m->n ? m->ctx[rank] : nullptr; }
const char* dbg_multi_transport(const dbg_multi* m) { return m && m->n ? m->comm[0].T->name() : ""; }

int dbg_multi_reads_to_graph(dbg_multi* m, int k, const dbg_seqset* const* seqs, uint32_t min_kmer_obs, int stranded, int reduce_op,
                             dbg_multi_info* infos, dbg_graph** graphs_out) {
    if (!m || !seqs || !graphs_out) return DBG_E_BADARG;
    std::vector<std::thread> th;
    std::vector<int> rcs(m->n, DBG_OK);
    for (int i = 0; i < m->n; i++) {
        graphs_out[i] = nullptr;
        th.emplace_back([&, i] {
            cudaSetDevice(m->ctx[i]->c.device);
            rcs[i] = dbg_reads_to_graph_multi(&m->comm[i], k, seqs[i], min_kmer_obs, stranded, reduce_op, infos ? infos + i : nullptr, graphs_out + i);
        });
    }
    for (auto& t : th) t.join();
    for (int i = 0; i < m->n; i++) if (rcs[i] != DBG_OK) return rcs[i];
    return DBG_OK;
}

/* pure host helpers of the multi-GPU plan (CPU tests) */
int dbg_plan_owner_bounds(uint64_t n_buckets, int n_ranks, uint64_t* bounds_out) {
    if (!bounds_out || n_ranks < 1 || n_ranks > DBG_MAX_RANKS) return DBG_E_BADARG;
    owner_bounds(n_buckets, n_ranks, reinterpret_cast<u64*>(bounds_out));
    return DBG_OK;
}
int dbg_plan_quantile_cuts(const uint64_t* hist, uint64_t n_bins, int n_ranks, uint64_t* cuts_out) {
    if (!hist || !cuts_out || n_ranks < 1 || n_ranks > DBG_MAX_RANKS) return DBG_E_BADARG;
    quantile_cuts(reinterpret_cast<const u64*>(hist), n_bins, n_ranks, reinterpret_cast<u64*>(cuts_out));
    return DBG_OK;
}

/* Layout of the fused exchange (scatter_buckets_kernel + bucket_totals_kernel + the scan, restated on the host for the CPU tests):
 * all_counts[s * n_buckets + b] = records of bucket b on rank s.  For sender `me`: dst_off[b] = position (in records) of its chunk of
 * bucket b inside the window of the bucket's owner = (records of the owner's buckets below b, all senders) + (records of bucket b
 * on the senders below me).  recv_total[r] = records rank r receives. */
int dbg_plan_exchange_layout(const uint32_t* all_counts, int n_ranks, int me, uint64_t n_buckets, uint64_t* dst_off /* n_buckets */,
                             uint64_t* recv_total /* n_ranks */) {
    if (!all_counts || !dst_off || !recv_total || n_ranks < 1 || n_ranks > DBG_MAX_RANKS || me < 0 || me >= n_ranks) return DBG_E_BADARG;
    u64 bounds[DBG_MAX_RANKS + 1];
    owner_bounds(n_buckets, n_ranks, bounds);
    for (int r = 0; r < n_ranks; r++) {
        u64 goff = 0;   // goff[b] - goff[bounds[r]] of the device code
        for (u64 b = bounds[r]; b < bounds[r + 1]; b++) {
            u64 tot = 0, pre = 0;
            for (int s2 = 0; s2 < n_ranks; s2++) {
                const u64 v = all_counts[(u64)s2 * n_buckets + b];
                if (s2 < me) pre += v;
                tot += v;
            }
            dst_off[b] = goff + pre;
            goff += tot;
        }
        recv_total[r] = goff;
    }
    return DBG_OK;
}

}  // extern "C"
