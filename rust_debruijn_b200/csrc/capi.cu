// extern "C" boundary (include/dbg_b200.h): opaque handles over the device-resident objects, host <->
// device copies, the synth-v1 generator.  No torch types, no exceptions across the ABI.
#include <new>

#include "common.cuh"

using namespace dbg;


namespace dbg {

void free_seqset(SeqSet* s) {
    if (!s) return;
    if (s->n_pending > 0) spin_sync(s->ctx->copy_stream);   // an asynchronous upload nobody consumed: let it land first
    for (int i = 0; i < 8; i++) if (s->pend_ev[i]) cudaEventDestroy(s->pend_ev[i]);
    if (s->owned) {
        cudaStream_t st = s->ctx->stream;
        if (s->words) cudaFreeAsync(s->words, st);
        if (s->start) cudaFreeAsync(s->start, st);
        if (s->length) cudaFreeAsync(s->length, st);
        if (s->seq_exts) cudaFreeAsync(s->seq_exts, st);
    }
    delete reinterpret_cast<dbg_seqset*>(s);
}
void free_table(Table* t) {
    if (!t) return;
    cudaStream_t st = t->ctx->stream;
    if (t->lo) cudaFreeAsync(t->lo, st);
    if (t->hi) cudaFreeAsync(t->hi, st);
    if (t->exts) cudaFreeAsync(t->exts, st);
    if (t->counts) cudaFreeAsync(t->counts, st);
    if (t->all_lo) cudaFreeAsync(t->all_lo, st);
    if (t->all_hi) cudaFreeAsync(t->all_hi, st);
    if (t->colors) cudaFreeAsync(t->colors, st);
    delete reinterpret_cast<dbg_kmer_table*>(t);
}
void free_graph(Graph* g) {
    if (!g) return;
    cudaStream_t st = g->ctx->stream;
    if (g->words) cudaFreeAsync(g->words, st);
    if (g->start) cudaFreeAsync(g->start, st);
    if (g->length) cudaFreeAsync(g->length, st);
    if (g->exts) cudaFreeAsync(g->exts, st);
    if (g->data) cudaFreeAsync(g->data, st);
    delete reinterpret_cast<dbg_graph*>(g);
}

// ---- synth-v1: counter-based splitmix64, identical bits on CPU and GPU ----
__host__ __device__ __forceinline__ u64 sm64(u64 x) {
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void synth_reads_kernel(u64 R, u64 s0, u64 s1, u64 s2, u64 s3, u32 err_thr, u64 G, u64* words, u64 n_words,
                                   u64* start, u32* length) {
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < R) { start[w] = 150 * w; length[w] = 150; }
    if (w >= n_words) return;
    u64 out = 0;
    u64 cur_read = ~0ull, rstart = 0, strand = 0;
    for (int t = 0; t < 32; t++) {
        u64 pos = w * 32 + t;
        u64 i = pos / 150;
        if (i >= R) break;
        u32 b = (u32)(pos - i * 150);
        if (i != cur_read) {
            cur_read = i;
            rstart = sm64(s1 + i) % (G - 149);
            strand = sm64(s2 + i) >> 63;
        }
        u64 gi = strand ? rstart + 149 - b : rstart + b;
        u32 base = (u32)(sm64(s0 + (gi >> 5)) >> (62 - 2 * (gi & 31))) & 3u;
        if (strand) base = 3u - base;
        u64 h = sm64(s3 + 256 * i + b);
        if ((u32)(h & 0xFFFFFF) < err_thr) base = (base + 1 + (u32)((h >> 24) % 3)) & 3u;
        out |= (u64)base << (62 - 2 * t);
    }
    words[w] = out;
}

int synth_reads_dev(Ctx* c, u64 R, u64 seed, u32 err_thr, SeqSet** out) {
    *out = nullptr;
    if (R == 0) DBG_SET_ERR(c, DBG_E_BADARG, "n_reads == 0");
    dbg_seqset* h = new (std::nothrow) dbg_seqset();
    SeqSet* s = &h->s;
    memset(s->pend_ev, 0, sizeof(s->pend_ev));
    s->ctx = c;
    s->n_seqs = R;
    s->n_words = (150 * R + 31) / 32;
    s->uniform_len = 150;
    s->max_len = 150;
    s->contiguous = true;
    s->base0 = 0;
    s->total_end = 150 * R;
    DBuf<u64> words, start;
    DBuf<u32> length;
    TRY(words.alloc_pool(c, s->n_words + 2));
    TRY(words.zero());
    TRY(start.alloc_pool(c, R));
    TRY(length.alloc_pool(c, R));
    u64 G = (150 * R + 49) / 50;
    u64 n = s->n_words > R ? s->n_words : R;
    synth_reads_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(R, sm64(4 * seed + 0), sm64(4 * seed + 1), sm64(4 * seed + 2),
                                                                sm64(4 * seed + 3), err_thr, G, words.p, s->n_words, start.p,
                                                                length.p);
    TRY(check_launch(c, "synth_reads"));
    TRY(sync(c));
    s->words = words.take(); s->start = start.take(); s->length = length.take();
    *out = s;
    return DBG_OK;
}

}  // namespace dbg

#define CTX(ctx) (&(ctx)->c)
#define NULLCHK(ctx, p) do { if (!(p)) DBG_SET_ERR(CTX(ctx), DBG_E_BADARG, "null argument: %s", #p); } while (0)

extern "C" {

int dbg_ctx_create(int device, dbg_ctx** out) {
    if (!out) return DBG_E_BADARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return DBG_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return DBG_E_CUDA;
    dbg_ctx* h = new (std::nothrow) dbg_ctx();
    if (!h) return DBG_E_OOM;
    Ctx* c = &h->c;
    c->device = device;
    memset(&c->stats, 0, sizeof(c->stats));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return DBG_E_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return DBG_E_CUDA; }
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->stream); delete h; return DBG_E_CUDA; }
    cudaMemPoolProps pp;
    memset(&pp, 0, sizeof(pp));
    pp.allocType = cudaMemAllocationTypePinned;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    if (cudaMemPoolCreate(&c->pool, &pp) != cudaSuccess) { cudaStreamDestroy(c->stream); delete h; return DBG_E_CUDA; }
    unsigned long long thr = ~0ull;  // keep freed blocks: steady-state calls do no driver allocation
    cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (cudaMallocHost((void**)&c->h_scratch, 64 * 8) != cudaSuccess) { cudaMemPoolDestroy(c->pool); cudaStreamDestroy(c->stream); delete h; return DBG_E_CUDA; }
    for (int i = 0; i < 12; i++) cudaEventCreate(&c->ev[i]);
    *out = h;
    return DBG_OK;
}

void dbg_ctx_destroy(dbg_ctx* ctx) {
    if (!ctx) return;
    Ctx* c = CTX(ctx);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 12; i++) cudaEventDestroy(c->ev[i]);
    cudaFreeHost(c->h_scratch);
    if (c->arena) cudaFree(c->arena);
    cudaMemPoolDestroy(c->pool);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete ctx;
}

const char* dbg_last_error(const dbg_ctx* ctx) { return ctx ? ctx->c.err.c_str() : "null ctx"; }

int dbg_stats_get(const dbg_ctx* ctx, dbg_stats* out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = ctx->c.stats;
    out->gpu_launches = ctx->c.launches;
    return DBG_OK;
}

int dbg_ctx_set_param(dbg_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    if (!strcmp(name, "msp_p")) {
        if (value < 0 || value > 16) DBG_SET_ERR(c, DBG_E_BADARG, "msp_p must be in [0,16]");
        c->msp_p = (int)value;
    } else if (!strcmp(name, "bucket_occ")) {
        if (value < 0) DBG_SET_ERR(c, DBG_E_BADARG, "bucket_occ must be >= 0");
        c->target_bucket_occ = (int)value;
    } else if (!strcmp(name, "dedup")) {
        c->dedup = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
    } else if (!strcmp(name, "mem_budget_bytes")) {
        c->mem_budget_bytes = value > 0 ? (u64)value : 0;
    } else if (!strcmp(name, "valid_est_div")) {
        c->valid_est_div = value > 0 ? (u64)value : 0;
    } else if (!strcmp(name, "direct_partition")) {
        c->direct_partition = value ? 1 : 0;
    } else if (!strcmp(name, "direct_min_tiles")) {
        c->direct_min_tiles = value > 0 ? (u64)value : 0;
    } else if (!strcmp(name, "fast_compress")) {
        c->no_fast_compress = value ? 0 : 1;
    } else {
        DBG_SET_ERR(c, DBG_E_BADARG, "unknown parameter %s", name);
    }
    return DBG_OK;
}

int dbg_ctx_synchronize(dbg_ctx* ctx) {
    if (!ctx) return DBG_E_BADARG;
    cudaSetDevice(ctx->c.device);
    return sync(CTX(ctx));
}

void* dbg_ctx_stream(dbg_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }

// ---- sequences --------------------------------------------------------------------------------------
int dbg_seqset_upload(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, const uint64_t* start,
                      const uint32_t* length, const uint8_t* seq_exts, uint64_t n_seqs, dbg_seqset** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    cudaSetDevice(c->device);
    if (n_seqs && (!start || !length)) DBG_SET_ERR(c, DBG_E_BADARG, "start/length are required");
    if (n_words && !words) DBG_SET_ERR(c, DBG_E_BADARG, "words is null");
    // validate extents and detect the uniform layout (start[i] = i*L, length[i] = L) on the host
    u32 max_len = 0;
    bool uniform = n_seqs > 0, contiguous = n_seqs > 0;
    for (u64 i = 0; i < n_seqs; i++) {
        const u64 limit = n_words * 32;   // overflow-safe: start near 2^64 must not wrap past the check
        if (start[i] > limit || length[i] > limit - start[i]) DBG_SET_ERR(c, DBG_E_BADARG, "sequence %llu runs past the packed words", (unsigned long long)i);
        u64 e = (u64)start[i] + length[i];
        if (length[i] > max_len) max_len = length[i];
        if (uniform && (length[i] != length[0] || start[i] != i * (u64)length[0])) uniform = false;
        if (contiguous && i + 1 < n_seqs && start[i + 1] != e) contiguous = false;
    }
    dbg_seqset* h = new (std::nothrow) dbg_seqset();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    SeqSet* s = &h->s;
    memset(s->pend_ev, 0, sizeof(s->pend_ev));
    s->ctx = c; s->n_seqs = n_seqs; s->n_words = n_words; s->max_len = max_len;
    s->uniform_len = uniform ? length[0] : 0;
    s->contiguous = contiguous;
    if (contiguous) { s->base0 = start[0]; s->total_end = (u64)start[n_seqs - 1] + length[n_seqs - 1]; }
    DBuf<u64> dw, ds;
    DBuf<u32> dl;
    DBuf<u8> de;
    int rc = DBG_OK;
    do {
        if ((rc = dw.alloc_pool(c, n_words + 2)) != DBG_OK) break;
        if ((rc = dw.zero()) != DBG_OK) break;
        if (n_words && cudaMemcpyAsync(dw.p, words, n_words * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = DBG_E_CUDA; break; }
        if (!uniform) {
            if ((rc = ds.alloc_pool(c, n_seqs)) != DBG_OK) break;
            if ((rc = dl.alloc_pool(c, n_seqs)) != DBG_OK) break;
            if (n_seqs && cudaMemcpyAsync(ds.p, start, n_seqs * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = DBG_E_CUDA; break; }
            if (n_seqs && cudaMemcpyAsync(dl.p, length, n_seqs * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = DBG_E_CUDA; break; }
        }
        if (seq_exts) {
            if ((rc = de.alloc_pool(c, n_seqs)) != DBG_OK) break;
            if (n_seqs && cudaMemcpyAsync(de.p, seq_exts, n_seqs, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = DBG_E_CUDA; break; }
        }
    } while (0);
    if (rc != DBG_OK) {
        if (rc == DBG_E_CUDA) c->err = std::string("seqset upload: ") + cudaGetErrorString(cudaGetLastError());
        delete h;
        return rc;
    }
    s->words = dw.take();
    if (!uniform) { s->start = ds.take(); s->length = dl.take(); }
    if (seq_exts) s->seq_exts = de.take();
    *out = h;
    return DBG_OK;
}

static int upload_uniform_impl(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, uint64_t n_seqs, uint32_t read_len,
                               const uint8_t* seq_exts, bool pipelined, dbg_seqset** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    cudaSetDevice(c->device);
    if (n_seqs == 0 || read_len == 0) DBG_SET_ERR(c, DBG_E_BADARG, "n_seqs and read_len must be > 0");
    if (!words || n_seqs * (u64)read_len > n_words * 32) DBG_SET_ERR(c, DBG_E_BADARG, "reads run past the packed words");
    dbg_seqset* h = new (std::nothrow) dbg_seqset();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    SeqSet* s = &h->s;
    memset(s->pend_ev, 0, sizeof(s->pend_ev));
    s->ctx = c; s->n_seqs = n_seqs; s->n_words = n_words; s->max_len = read_len; s->uniform_len = read_len;
    s->contiguous = true; s->base0 = 0; s->total_end = n_seqs * (u64)read_len;
    DBuf<u64> dw;
    DBuf<u8> de;
    int rc = dw.alloc_pool(c, n_words + 2);
    if (rc == DBG_OK && cudaMemsetAsync(dw.p + n_words, 0, 16, c->stream) != cudaSuccess) rc = DBG_E_CUDA;
    if (rc == DBG_OK && seq_exts) {
        rc = de.alloc_pool(c, n_seqs);
        if (rc == DBG_OK && cudaMemcpyAsync(de.p, seq_exts, n_seqs, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = DBG_E_CUDA;
    }
    if (rc == DBG_OK) {
        const int nch = (pipelined && n_words >= (1u << 22)) ? 8 : 1;   // >= 32 MB: worth overlapping
        if (nch == 1) {
            if (cudaMemcpyAsync(dw.p, words, n_words * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = DBG_E_CUDA;
        } else {
            // the buffer was allocated on c->stream: order the copy stream after it, then chunked copies + events
            cudaEvent_t e0;
            cudaEventCreateWithFlags(&e0, cudaEventDisableTiming);
            cudaEventRecord(e0, c->stream);
            cudaStreamWaitEvent(c->copy_stream, e0, 0);
            cudaEventDestroy(e0);
            for (int i = 0; i < nch && rc == DBG_OK; i++) {
                u64 w0 = n_words * i / nch, w1 = n_words * (i + 1) / nch;
                if (cudaMemcpyAsync(dw.p + w0, words + w0, (w1 - w0) * 8, cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess) rc = DBG_E_CUDA;
                if (rc == DBG_OK && cudaEventCreateWithFlags(&s->pend_ev[i], cudaEventDisableTiming) != cudaSuccess) rc = DBG_E_CUDA;
                if (rc == DBG_OK && cudaEventRecord(s->pend_ev[i], c->copy_stream) != cudaSuccess) rc = DBG_E_CUDA;
                s->pend_words_end[i] = w1;
                s->n_pending = i + 1;
            }
        }
    }
    if (rc != DBG_OK) {
        if (rc == DBG_E_CUDA) c->err = std::string("seqset upload: ") + cudaGetErrorString(cudaGetLastError());
        cudaStreamSynchronize(c->copy_stream);
        for (int i = 0; i < 8; i++) if (s->pend_ev[i]) cudaEventDestroy(s->pend_ev[i]);
        delete h;
        return rc;
    }
    s->words = dw.take();
    if (seq_exts) s->seq_exts = de.take();
    *out = h;
    return DBG_OK;
}

int dbg_seqset_upload_uniform(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, uint64_t n_seqs, uint32_t read_len,
                              const uint8_t* seq_exts, dbg_seqset** out) {
    return upload_uniform_impl(ctx, words, n_words, n_seqs, read_len, seq_exts, false, out);
}

int dbg_seqset_upload_uniform_async(dbg_ctx* ctx, const uint64_t* words, uint64_t n_words, uint64_t n_seqs, uint32_t read_len,
                                    const uint8_t* seq_exts, dbg_seqset** out) {
    return upload_uniform_impl(ctx, words, n_words, n_seqs, read_len, seq_exts, true, out);
}

// ---- ingest: DnaString::from_acgt_bytes (src/dna_string.rs:224-250, AVX2 twin src/bitops_avx2.rs:8-132) + PackedDnaStringSet::add
// (src/dna_string.rs:811-821) for a whole batch of ASCII sequences.  Thread = one output word (32 bases of the
// concatenation): finds the sequence holding its first base (binary search over the output offsets), then walks
// the ASCII bytes, crossing sequence boundaries as needed.  A/a=0 C/c=1 G/g=2 T/t=3, anything else -> A (counted).
// SipHash-1-3 with a zero key over the byte stream [len(name) as u64 LE][name][pos as u64 LE]: what Rust's
// std::collections::hash_map::DefaultHasher yields for `read_name.hash(&mut h); pos.hash(&mut h_clone); h_clone.finish()`
// (DnaString::from_acgt_bytes_hashn, src/dna_string.rs:254-278: Hash for [u8] writes the length prefix, then the bytes;
// usize writes its 8 native-endian bytes).
__host__ __device__ __forceinline__ u64 sip_rotl(u64 x, int b) { return (x << b) | (x >> (64 - b)); }
struct Sip13 {
    u64 v0, v1, v2, v3;
    __host__ __device__ void init() { v0 = 0x736f6d6570736575ull; v1 = 0x646f72616e646f6dull; v2 = 0x6c7967656e657261ull; v3 = 0x7465646279746573ull; }
    __host__ __device__ void round() {
        v0 += v1; v1 = sip_rotl(v1, 13); v1 ^= v0; v0 = sip_rotl(v0, 32);
        v2 += v3; v3 = sip_rotl(v3, 16); v3 ^= v2;
        v0 += v3; v3 = sip_rotl(v3, 21); v3 ^= v0;
        v2 += v1; v1 = sip_rotl(v1, 17); v1 ^= v2; v2 = sip_rotl(v2, 32);
    }
    __host__ __device__ void word(u64 m) { v3 ^= m; round(); v0 ^= m; }
    __host__ __device__ u64 finish(u64 b) { v3 ^= b; round(); v0 ^= b; v2 ^= 0xff; round(); round(); round(); return v0 ^ v1 ^ v2 ^ v3; }
};
__host__ __device__ inline u64 sip13_name_pos(const u8* name, u32 L, u64 pos) {
    Sip13 h;
    h.init();
    h.word((u64)L);
    u64 m = 0;
    const u32 total = L + 8;
    for (u32 i = 0; i < total; i++) {
        const u64 byte = i < L ? name[i] : (pos >> (8 * (i - L))) & 0xffull;
        m |= byte << (8 * (i & 7));
        if ((i & 7) == 7) { h.word(m); m = 0; }
    }
    return h.finish(m | ((u64)((16 + L) & 0xffu) << 56));
}
// plain SipHash-1-3 (zero key) of a byte string: the primitive alone, exported so that tests can pin it to an independent
// implementation (CPython's bytes hash under PYTHONHASHSEED=0 is the same function)
static u64 sip13_bytes(const u8* p, u64 n) {
    Sip13 h;
    h.init();
    u64 m = 0;
    for (u64 i = 0; i < n; i++) {
        m |= (u64)p[i] << (8 * (i & 7));
        if ((i & 7) == 7) { h.word(m); m = 0; }
    }
    return h.finish(m | ((n & 0xff) << 56));
}

__global__ void ascii_pack_kernel(const u8* __restrict__ ascii, const u64* __restrict__ in_start, const u64* __restrict__ out_start,
                                  const u32* __restrict__ length, u64 n_seqs, u64 n_bases, u64* __restrict__ words,
                                  u64* __restrict__ n_invalid, const u8* __restrict__ names, const u64* __restrict__ name_start,
                                  const u32* __restrict__ name_len) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 bad = 0;
    if (w * 32 < n_bases) {
        const u64 b0 = w * 32;
        u64 lo = 0, hi = n_seqs;   // last sequence with out_start <= b0 (skips empty sequences sharing an offset)
        while (lo < hi) { u64 m = (lo + hi) >> 1; if (out_start[m] <= b0) lo = m + 1; else hi = m; }
        u64 si = lo - 1;
        u64 off = b0 - out_start[si];
        u64 out = 0;
        const int nb = (int)min((u64)32, n_bases - b0);
        for (int t = 0; t < nb; t++) {
            while (off >= length[si]) { si++; off = 0; }
            const u32 c = ascii[in_start[si] + off];
            const u32 up = c & 0xDFu;   // fold case
            const bool ok = up == 'A' || up == 'C' || up == 'G' || up == 'T';
            u32 v = (c >> 1) & 3u;      // A 0, C 1, T 2, G 3
            v ^= v >> 1;                // swap 2 <-> 3
            if (!ok) {   // from_acgt_bytes: 'A'; from_acgt_bytes_hashn: repeatable pseudo-random base from (read name, position)
                v = names ? (u32)(sip13_name_pos(names + name_start[si], name_len[si], off) & 3ull) : 0u;
                bad++;
            }
            out |= (u64)v << (62 - 2 * t);
            off++;
        }
        words[w] = out;
    }
    for (int o = 16; o; o >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, o);
    __shared__ u32 s_bad[8];
    if ((threadIdx.x & 31) == 0) s_bad[threadIdx.x >> 5] = bad;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += s_bad[i];
        if (t) atomicAdd(n_invalid, (u64)t);   // one atomic per CTA
    }
}

static int seqset_from_ascii_impl(dbg_ctx* ctx, const uint8_t* ascii, uint64_t n_bytes, const uint64_t* start, const uint32_t* length,
                                  const uint8_t* names, uint64_t n_name_bytes, const uint64_t* name_start, const uint32_t* name_len,
                                  const uint8_t* seq_exts, uint64_t n_seqs, uint64_t* n_invalid, dbg_seqset** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    cudaSetDevice(c->device);
    if (n_seqs && (!start || !length || !ascii)) DBG_SET_ERR(c, DBG_E_BADARG, "null argument");
    if (names) {
        if (!name_start || !name_len) DBG_SET_ERR(c, DBG_E_BADARG, "name_start / name_len are required with names");
        for (u64 i = 0; i < n_seqs; i++)
            if (name_start[i] > n_name_bytes || name_len[i] > n_name_bytes - name_start[i]) DBG_SET_ERR(c, DBG_E_BADARG, "name %llu runs past the name buffer", (unsigned long long)i);
    }
    std::vector<u64> ostart(n_seqs + 1, 0);
    u32 max_len = 0;
    bool uniform = n_seqs > 0;
    for (u64 i = 0; i < n_seqs; i++) {
        if (start[i] > n_bytes || length[i] > n_bytes - start[i]) DBG_SET_ERR(c, DBG_E_BADARG, "sequence %llu runs past the ASCII buffer", (unsigned long long)i);
        ostart[i + 1] = ostart[i] + length[i];
        if (length[i] > max_len) max_len = length[i];
        if (length[i] != length[0]) uniform = false;
    }
    const u64 n_bases = ostart[n_seqs], n_words = (n_bases + 31) / 32;
    dbg_seqset* h = new (std::nothrow) dbg_seqset();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    SeqSet* s = &h->s;
    memset(s->pend_ev, 0, sizeof(s->pend_ev));
    s->ctx = c; s->n_seqs = n_seqs; s->n_words = n_words; s->max_len = max_len;
    s->uniform_len = (uniform && n_seqs) ? length[0] : 0;
    s->contiguous = n_seqs > 0; s->base0 = 0; s->total_end = n_bases;
    DBuf<u64> dw, ds, din, dbad, dns;
    DBuf<u32> dl, dnl;
    DBuf<u8> de, dasc, dnm;
    int rc = arena_begin(c);
    do {
        if (rc != DBG_OK) break;
        if ((rc = dw.alloc_pool(c, n_words + 2)) != DBG_OK) break;
        if ((rc = dw.zero()) != DBG_OK) break;
        if ((rc = ds.alloc_pool(c, n_seqs + 1)) != DBG_OK) break;
        if ((rc = dl.alloc_pool(c, n_seqs)) != DBG_OK) break;
        if ((rc = din.alloc(c, n_seqs)) != DBG_OK) break;
        if ((rc = dasc.alloc(c, n_bytes)) != DBG_OK) break;
        if ((rc = dbad.alloc(c, 1)) != DBG_OK) break;
        if ((rc = dbad.zero()) != DBG_OK) break;
        if (seq_exts) {
            if ((rc = de.alloc_pool(c, n_seqs)) != DBG_OK) break;
            if (n_seqs && cudaMemcpyAsync(de.p, seq_exts, n_seqs, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = DBG_E_CUDA; break; }
        }
        if (n_seqs) {
            if (cudaMemcpyAsync(ds.p, ostart.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaMemcpyAsync(dl.p, length, n_seqs * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaMemcpyAsync(din.p, start, n_seqs * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                (n_bytes && cudaMemcpyAsync(dasc.p, ascii, n_bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)) { rc = DBG_E_CUDA; break; }
        }
        if (names && n_seqs) {
            if ((rc = dnm.alloc(c, n_name_bytes)) != DBG_OK) break;
            if ((rc = dns.alloc(c, n_seqs)) != DBG_OK) break;
            if ((rc = dnl.alloc(c, n_seqs)) != DBG_OK) break;
            if ((n_name_bytes && cudaMemcpyAsync(dnm.p, names, n_name_bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) ||
                cudaMemcpyAsync(dns.p, name_start, n_seqs * 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaMemcpyAsync(dnl.p, name_len, n_seqs * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = DBG_E_CUDA; break; }
        }
        if (n_words) {
            ascii_pack_kernel<<<grid_for(n_words, 256), 256, 0, c->stream>>>(dasc.p, din.p, ds.p, dl.p, n_seqs, n_bases, dw.p, dbad.p,
                                                                           names ? dnm.p : nullptr, dns.p, dnl.p);
            if ((rc = check_launch(c, "ascii_pack")) != DBG_OK) break;
        }
        u64 bad = 0;
        if ((rc = read_u64(c, dbad.p, &bad)) != DBG_OK) break;   // also: ostart may go out of scope after this sync
        if (n_invalid) *n_invalid = bad;
    } while (0);
    if (rc != DBG_OK) {
        if (rc == DBG_E_CUDA) c->err = std::string("seqset_from_ascii: ") + cudaGetErrorString(cudaGetLastError());
        delete h;
        return rc;
    }
    s->words = dw.take(); s->start = ds.take(); s->length = dl.take();
    if (seq_exts) s->seq_exts = de.take();
    *out = h;
    return DBG_OK;
}

int dbg_seqset_from_ascii(dbg_ctx* ctx, const uint8_t* ascii, uint64_t n_bytes, const uint64_t* start, const uint32_t* length,
                          const uint8_t* seq_exts, uint64_t n_seqs, uint64_t* n_invalid, dbg_seqset** out) {
    return seqset_from_ascii_impl(ctx, ascii, n_bytes, start, length, nullptr, 0, nullptr, nullptr, seq_exts, n_seqs, n_invalid, out);
}
int dbg_seqset_from_ascii_hashn(dbg_ctx* ctx, const uint8_t* ascii, uint64_t n_bytes, const uint64_t* start, const uint32_t* length,
                                const uint8_t* names, uint64_t n_name_bytes, const uint64_t* name_start, const uint32_t* name_len,
                                const uint8_t* seq_exts, uint64_t n_seqs, uint64_t* n_invalid, dbg_seqset** out) {
    if (ctx && !names) DBG_SET_ERR(CTX(ctx), DBG_E_BADARG, "names is null (use dbg_seqset_from_ascii)");
    return seqset_from_ascii_impl(ctx, ascii, n_bytes, start, length, names, n_name_bytes, name_start, name_len, seq_exts, n_seqs, n_invalid, out);
}
uint64_t dbg_siphash13(const uint8_t* bytes, uint64_t n) { return sip13_bytes(bytes, n); }

int dbg_seqset_wrap_device(dbg_ctx* ctx, const uint64_t* d_words, uint64_t n_words, const uint64_t* d_start,
                           const uint32_t* d_length, const uint8_t* d_seq_exts, uint64_t n_seqs, uint32_t max_len,
                           dbg_seqset** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    if (n_seqs && (!d_start || !d_length || !d_words)) DBG_SET_ERR(c, DBG_E_BADARG, "null device pointer");
    dbg_seqset* h = new (std::nothrow) dbg_seqset();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    SeqSet* s = &h->s;
    memset(s->pend_ev, 0, sizeof(s->pend_ev));
    s->ctx = c; s->owned = false;
    s->words = (u64*)d_words; s->n_words = n_words;
    s->start = (u64*)d_start; s->length = (u32*)d_length; s->seq_exts = (u8*)d_seq_exts;
    s->n_seqs = n_seqs; s->max_len = max_len; s->uniform_len = 0;
    *out = h;
    return DBG_OK;
}

int dbg_seqset_synth(dbg_ctx* ctx, uint64_t n_reads, uint64_t seed, uint32_t err_thr, dbg_seqset** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    cudaSetDevice(ctx->c.device);
    SeqSet* s = nullptr;
    int rc = synth_reads_dev(CTX(ctx), n_reads, seed, err_thr, &s);
    *out = reinterpret_cast<dbg_seqset*>(s);
    return rc;
}

uint64_t dbg_seqset_len(const dbg_seqset* s) { return s ? s->s.n_seqs : 0; }
uint64_t dbg_seqset_n_words(const dbg_seqset* s) { return s ? s->s.n_words : 0; }

int dbg_seqset_copy_out(const dbg_seqset* h, uint64_t* words, uint64_t* start, uint32_t* length) {
    if (!h) return DBG_E_BADARG;
    SeqSet* s = const_cast<SeqSet*>(&h->s);
    Ctx* c = s->ctx;
    cudaSetDevice(c->device);
    TRY(seqset_ready(c, s));
    if (words && s->n_words) CU(c, cudaMemcpyAsync(words, s->words, s->n_words * 8, cudaMemcpyDeviceToHost, c->stream));
    if (s->uniform_len && (!s->start || !s->length)) {
        for (u64 i = 0; i < s->n_seqs; i++) {
            if (start) start[i] = i * (u64)s->uniform_len;
            if (length) length[i] = s->uniform_len;
        }
    } else {
        if (start && s->n_seqs) CU(c, cudaMemcpyAsync(start, s->start, s->n_seqs * 8, cudaMemcpyDeviceToHost, c->stream));
        if (length && s->n_seqs) CU(c, cudaMemcpyAsync(length, s->length, s->n_seqs * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    return sync(c);
}

void dbg_seqset_free(dbg_seqset* s) { if (s) { cudaSetDevice(s->s.ctx->device); free_seqset(&s->s); } }

// ---- filter_kmers -------------------------------------------------------------------------------------
int dbg_filter_kmers(dbg_ctx* ctx, int k, const dbg_seqset* seqs, uint32_t min_kmer_obs, int stranded,
                     int report_all_kmers, uint64_t memory_size_gb, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    NULLCHK(ctx, seqs);
    cudaSetDevice(ctx->c.device);
    Table* t = nullptr;
    int rc = filter_kmers_dev(CTX(ctx), k, &seqs->s, min_kmer_obs, stranded != 0, report_all_kmers != 0, memory_size_gb, &t);
    *out = reinterpret_cast<dbg_kmer_table*>(t);
    return rc;
}

int dbg_filter_kmers_host(dbg_ctx* ctx, int k, const uint64_t* words, uint64_t n_words, const uint64_t* start,
                          const uint32_t* length, const uint8_t* seq_exts, uint64_t n_seqs, uint32_t min_kmer_obs,
                          int stranded, int report_all_kmers, uint64_t memory_size_gb, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    dbg_seqset* s = nullptr;
    int rc = dbg_seqset_upload(ctx, words, n_words, start, length, seq_exts, n_seqs, &s);
    if (rc != DBG_OK) return rc;
    rc = dbg_filter_kmers(ctx, k, s, min_kmer_obs, stranded, report_all_kmers, memory_size_gb, out);
    dbg_seqset_free(s);
    return rc;
}

uint64_t dbg_table_len(const dbg_kmer_table* t) { return t ? t->t.n : 0; }
uint64_t dbg_table_all_len(const dbg_kmer_table* t) { return t ? t->t.n_all : 0; }
uint64_t dbg_table_n_input(const dbg_kmer_table* t) { return t ? t->t.n_input : 0; }
int dbg_table_k(const dbg_kmer_table* t) { return t ? t->t.k : 0; }

int dbg_table_copy_out(const dbg_kmer_table* h, uint64_t* kmers_lo, uint64_t* kmers_hi, uint8_t* exts,
                       uint16_t* counts, uint64_t* all_lo, uint64_t* all_hi) {
    if (!h) return DBG_E_BADARG;
    const Table* t = &h->t;
    Ctx* c = t->ctx;
    cudaSetDevice(c->device);
    if (t->n) {
        if (kmers_lo) CU(c, cudaMemcpyAsync(kmers_lo, t->lo, t->n * 8, cudaMemcpyDeviceToHost, c->stream));
        if (kmers_hi && t->hi) CU(c, cudaMemcpyAsync(kmers_hi, t->hi, t->n * 8, cudaMemcpyDeviceToHost, c->stream));
        if (exts) CU(c, cudaMemcpyAsync(exts, t->exts, t->n, cudaMemcpyDeviceToHost, c->stream));
        if (counts) CU(c, cudaMemcpyAsync(counts, t->counts, t->n * 2, cudaMemcpyDeviceToHost, c->stream));
    }
    if (t->n_all) {
        if (all_lo) CU(c, cudaMemcpyAsync(all_lo, t->all_lo, t->n_all * 8, cudaMemcpyDeviceToHost, c->stream));
        if (all_hi && t->all_hi) CU(c, cudaMemcpyAsync(all_hi, t->all_hi, t->n_all * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    return sync(c);
}

__global__ void pack_vals_kernel(const u8* exts, const u16* counts, u32* val, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) val[i] = (u32)exts[i] | ((u32)counts[i] << 8);
}
__global__ void unpack_vals_kernel2(const u32* val, u8* exts, u16* counts, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { u32 v = val[i]; exts[i] = (u8)v; counts[i] = (u16)(v >> 8); }
}
// bad = 1: not ascending / duplicate; bad = 2: bits set above the 2k key bits (VarIntKmer equality is on raw storage,
// src/kmer.rs:438-442: such keys would sort and compare differently from the k-mers they spell)
__global__ void check_sorted_unique_kernel(const u64* lo, const u64* hi, u64 n, u32* bad, u64 mask_lo, u64 mask_hi) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((lo[i] & ~mask_lo) || (hi && (hi[i] & ~mask_hi))) { *bad = 2; return; }
    if (i + 1 >= n) return;
    bool ok = hi ? (hi[i] < hi[i + 1] || (hi[i] == hi[i + 1] && lo[i] < lo[i + 1])) : lo[i] < lo[i + 1];
    if (!ok) atomicMax(bad, 1u);
}

static int table_from_arrays(dbg_ctx* ctx, int k, uint64_t n, const void* kmers_lo, const void* kmers_hi,
                             const void* exts, const void* counts, cudaMemcpyKind kind, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    cudaSetDevice(c->device);
    if (k < 2 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [2,64]", k);
    if (n && (!kmers_lo || !exts || !counts || (k > 32 && !kmers_hi))) DBG_SET_ERR(c, DBG_E_BADARG, "null array");
    const int W = k <= 32 ? 1 : 2;
    dbg_kmer_table* h = new (std::nothrow) dbg_kmer_table();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    Table* t = &h->t;
    t->ctx = c; t->k = k; t->n = n; t->n_input = 0;
    *out = h;
    if (n == 0) return DBG_OK;
    DBuf<u64> alo, ahi, blo, bhi;
    DBuf<u32> av, bv, bad;
    DBuf<u8> de;
    DBuf<u16> dc;
    auto fail = [&](int rc) { free_table(t); *out = nullptr; return rc; };
#define T2(x) do { int _r = (x); if (_r != DBG_OK) return fail(_r); } while (0)
#define CU2(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { c->err = std::string("table_from_host: ") + cudaGetErrorString(_e); return fail(DBG_E_CUDA); } } while (0)
    T2(arena_begin(c));
    T2(alo.alloc_pool(c, n)); T2(blo.alloc_pool(c, n)); T2(av.alloc(c, n)); T2(bv.alloc(c, n)); T2(de.alloc_pool(c, n)); T2(dc.alloc_pool(c, n)); T2(bad.alloc(c, 1));
    if (W == 2) { T2(ahi.alloc_pool(c, n)); T2(bhi.alloc_pool(c, n)); }
    CU2(cudaMemcpyAsync(alo.p, kmers_lo, n * 8, kind, c->stream));
    if (W == 2) CU2(cudaMemcpyAsync(ahi.p, kmers_hi, n * 8, kind, c->stream));
    CU2(cudaMemcpyAsync(de.p, exts, n, kind, c->stream));
    CU2(cudaMemcpyAsync(dc.p, counts, n * 2, kind, c->stream));
    pack_vals_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(de.p, dc.p, av.p, n);
    T2(check_launch(c, "pack_vals"));
    u64 *rlo, *rhi;
    u32* rv;
    T2(radix_sort_pairs(c, W, 2 * k, n, alo.p, ahi.p, av.p, blo.p, bhi.p, bv.p, &rlo, &rhi, &rv));
    unpack_vals_kernel2<<<grid_for(n, 256), 256, 0, c->stream>>>(rv, de.p, dc.p, n);
    T2(check_launch(c, "unpack_vals"));
    T2(bad.zero());
    check_sorted_unique_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(rlo, W == 2 ? rhi : nullptr, n, bad.p, make_kp(k).mask_lo, make_kp(k).mask_hi);
    T2(check_launch(c, "check_sorted_unique"));
    u32 hbad = 0;
    CU2(cudaMemcpyAsync(c->h_scratch, bad.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CU2(spin_sync(c->stream));
    hbad = *(u32*)c->h_scratch;
    if (hbad) { c->err = hbad == 2 ? "k-mer words have bits set above the 2k key bits" : "duplicate k-mers in table"; return fail(DBG_E_BADARG); }
    t->lo = (rlo == alo.p) ? alo.take() : blo.take();
    if (W == 2) t->hi = (rhi == ahi.p) ? ahi.take() : bhi.take();
    t->exts = de.take();
    t->counts = dc.take();
#undef T2
#undef CU2
    return DBG_OK;
}

int dbg_table_from_host(dbg_ctx* ctx, int k, uint64_t n, const uint64_t* kmers_lo, const uint64_t* kmers_hi,
                        const uint8_t* exts, const uint16_t* counts, dbg_kmer_table** out) {
    return table_from_arrays(ctx, k, n, kmers_lo, kmers_hi, exts, counts, cudaMemcpyHostToDevice, out);
}
int dbg_table_from_device(dbg_ctx* ctx, int k, uint64_t n, const void* d_kmers_lo, const void* d_kmers_hi,
                          const void* d_exts, const void* d_counts, dbg_kmer_table** out) {
    return table_from_arrays(ctx, k, n, d_kmers_lo, d_kmers_hi, d_exts, d_counts, cudaMemcpyDeviceToDevice, out);
}
int dbg_table_from_device_sorted(dbg_ctx* ctx, int k, uint64_t n, const void* d_lo, const void* d_hi, const void* d_exts,
                                 const void* d_counts, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    cudaSetDevice(c->device);
    if (k < 2 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [2,64]", k);
    if (n && (!d_lo || !d_exts || !d_counts || (k > 32 && !d_hi))) DBG_SET_ERR(c, DBG_E_BADARG, "null array");
    const bool two = k > 32;
    dbg_kmer_table* h = new (std::nothrow) dbg_kmer_table();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    Table* t = &h->t;
    t->ctx = c; t->k = k; t->n = n;
    if (n == 0) { *out = h; return DBG_OK; }
    DBuf<u64> lo, hi;
    DBuf<u8> ex;
    DBuf<u16> cn;
    DBuf<u32> bad;
    int rc = arena_begin(c);
    if (rc == DBG_OK) rc = lo.alloc_pool(c, n);
    if (rc == DBG_OK && two) rc = hi.alloc_pool(c, n);
    if (rc == DBG_OK) rc = ex.alloc_pool(c, n);
    if (rc == DBG_OK) rc = cn.alloc_pool(c, n);
    if (rc == DBG_OK) rc = bad.alloc(c, 1);
    if (rc == DBG_OK) rc = bad.zero();
    if (rc != DBG_OK) { delete h; return rc; }
    cudaMemcpyAsync(lo.p, d_lo, n * 8, cudaMemcpyDeviceToDevice, c->stream);
    if (two) cudaMemcpyAsync(hi.p, d_hi, n * 8, cudaMemcpyDeviceToDevice, c->stream);
    cudaMemcpyAsync(ex.p, d_exts, n, cudaMemcpyDeviceToDevice, c->stream);
    cudaMemcpyAsync(cn.p, d_counts, n * 2, cudaMemcpyDeviceToDevice, c->stream);
    check_sorted_unique_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(lo.p, two ? hi.p : nullptr, n, bad.p, make_kp(k).mask_lo, make_kp(k).mask_hi);
    c->launches++;
    cudaMemcpyAsync(c->h_scratch, bad.p, 4, cudaMemcpyDeviceToHost, c->stream);
    if (spin_sync(c->stream) != cudaSuccess) { delete h; DBG_SET_ERR(c, DBG_E_CUDA, "table_from_device_sorted: %s", cudaGetErrorString(cudaGetLastError())); }
    if (*(u32*)c->h_scratch) { const bool hb = *(u32*)c->h_scratch == 2; delete h; DBG_SET_ERR(c, DBG_E_BADARG, hb ? "k-mer words have bits set above the 2k key bits" : "arrays are not ascending and distinct"); }
    t->lo = lo.take();
    if (two) t->hi = hi.take();
    t->exts = ex.take();
    t->counts = cn.take();
    *out = h;
    return DBG_OK;
}
int dbg_table_alloc(dbg_ctx* ctx, int k, uint64_t n, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    cudaSetDevice(c->device);
    if (k < 2 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [2,64]", k);
    dbg_kmer_table* h = new (std::nothrow) dbg_kmer_table();
    if (!h) DBG_SET_ERR(c, DBG_E_OOM, "host allocation failed");
    Table* t = &h->t;
    t->ctx = c; t->k = k; t->n = n;
    DBuf<u64> lo, hi;
    DBuf<u8> ex;
    DBuf<u16> cn;
    int rc = lo.alloc_pool(c, n);
    if (rc == DBG_OK && k > 32) rc = hi.alloc_pool(c, n);
    if (rc == DBG_OK) rc = ex.alloc_pool(c, n);
    if (rc == DBG_OK) rc = cn.alloc_pool(c, n);
    if (rc != DBG_OK) { delete h; return rc; }
    t->lo = lo.take();
    if (k > 32) t->hi = hi.take();
    t->exts = ex.take();
    t->counts = cn.take();
    *out = h;
    return DBG_OK;
}
int dbg_remove_censored_exts(dbg_ctx* ctx, dbg_kmer_table* table, int stranded, int sharded) {
    if (!ctx) return DBG_E_BADARG;
    NULLCHK(ctx, table);
    cudaSetDevice(ctx->c.device);
    return remove_censored_exts_dev(CTX(ctx), &table->t, stranded != 0, sharded != 0);
}
int dbg_table_prefix_hist(dbg_ctx* ctx, const dbg_kmer_table* t, int bits, void* d_hist) {
    if (!ctx || !d_hist) return DBG_E_BADARG;
    NULLCHK(ctx, t);
    cudaSetDevice(ctx->c.device);
    return table_prefix_hist_dev(CTX(ctx), &t->t, bits, (u32*)d_hist);
}
int dbg_table_device_ptrs(const dbg_kmer_table* t, void** kmers_lo, void** kmers_hi, void** exts, void** counts) {
    if (!t) return DBG_E_BADARG;
    if (kmers_lo) *kmers_lo = t->t.lo;
    if (kmers_hi) *kmers_hi = t->t.hi;
    if (exts) *exts = t->t.exts;
    if (counts) *counts = t->t.counts;
    return DBG_OK;
}

// ---- multi-GPU building blocks ------------------------------------------------------------------------------
int dbg_plan_filter(dbg_ctx* ctx, int k, uint64_t n_total, int* msp_p, int* bucket_bits) {
    if (!ctx || !msp_p || !bucket_bits) return DBG_E_BADARG;
    if (k < 4 || k > 64) DBG_SET_ERR(CTX(ctx), DBG_E_BADARG, "k=%d outside [4,64]", k);
    plan_filter(CTX(ctx), k, n_total, msp_p, bucket_bits);
    return DBG_OK;
}
uint64_t dbg_seqset_count_kmers(dbg_ctx* ctx, int k, const dbg_seqset* seqs) {
    if (!ctx || !seqs) return 0;
    const SeqSet* s = &seqs->s;
    if (s->uniform_len) return s->uniform_len >= (u32)k ? (u64)(s->uniform_len - k + 1) * s->n_seqs : 0;
    // general layouts: the partition stage counts on the device; do the same reduction here through a 0-bucket plan
    Partition* P = nullptr;
    cudaSetDevice(ctx->c.device);
    if (partition_reads_dev(CTX(ctx), k, s, 0, k - 3 < 1 ? 1 : (k - 3 > 12 ? 12 : k - 3), 0, &P) != DBG_OK) return 0;
    u64 n = P->n_input;
    free_partition(P);
    return n;
}
int dbg_partition_reads(dbg_ctx* ctx, int k, const dbg_seqset* seqs, int stranded, int msp_p, int bucket_bits,
                        dbg_partition** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    NULLCHK(ctx, seqs);
    cudaSetDevice(ctx->c.device);
    Partition* P = nullptr;
    int rc = partition_reads_dev(CTX(ctx), k, &seqs->s, stranded != 0, msp_p, bucket_bits, &P);
    *out = reinterpret_cast<dbg_partition*>(P);
    return rc;
}
uint64_t dbg_partition_n_records(const dbg_partition* p) { return p ? p->p.n_rec : 0; }
uint64_t dbg_partition_n_input(const dbg_partition* p) { return p ? p->p.n_input : 0; }
uint32_t dbg_partition_record_bytes(const dbg_partition* p) { return p ? (uint32_t)p->p.rec_words * 8 : 0; }
void* dbg_partition_records_dev(const dbg_partition* p) { return p ? (void*)p->p.rec : nullptr; }
int dbg_partition_bucket_counts(const dbg_partition* p, uint32_t* host_counts) {
    if (!p || !host_counts) return DBG_E_BADARG;
    Ctx* c = p->p.ctx;
    cudaSetDevice(c->device);
    CU(c, cudaMemcpyAsync(host_counts, p->p.bucket_count, ((u64)1 << p->p.bbits) * 4, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}
void dbg_partition_free(dbg_partition* p) { if (p) { cudaSetDevice(p->p.ctx->device); free_partition(&p->p); } }
int dbg_filter_from_records(dbg_ctx* ctx, int k, const void* d_records, uint64_t n_records, const uint32_t* h_counts,
                            uint32_t n_src, uint32_t n_local_buckets, uint64_t n_input_kmers_total,
                            uint32_t min_kmer_obs, int stranded, int report_all_kmers, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    cudaSetDevice(ctx->c.device);
    Table* t = nullptr;
    int rc = filter_from_records_dev(CTX(ctx), k, (const u64*)d_records, n_records, h_counts, n_src, n_local_buckets,
                                     n_input_kmers_total, min_kmer_obs, stranded != 0, report_all_kmers != 0, &t);
    *out = reinterpret_cast<dbg_kmer_table*>(t);
    return rc;
}

// ---- sharded compression ------------------------------------------------------------------------------------
int dbg_graph_from_device(dbg_ctx* ctx, int k, int stranded, uint64_t n_nodes, uint64_t n_bases, const void* d_words,
                          const void* d_start, const void* d_length, const void* d_exts, const void* d_data,
                          dbg_graph** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    cudaSetDevice(ctx->c.device);
    Graph* g = nullptr;
    int rc = graph_from_device_dev(CTX(ctx), k, stranded != 0, n_nodes, n_bases, (const u64*)d_words, (const u64*)d_start,
                                   (const u32*)d_length, (const u8*)d_exts, (const u16*)d_data, &g);
    *out = reinterpret_cast<dbg_graph*>(g);
    return rc;
}

void dbg_table_free(dbg_kmer_table* t) { if (t) { cudaSetDevice(t->t.ctx->device); free_table(&t->t); } }

// ---- compress ------------------------------------------------------------------------------------------
int dbg_compress_kmers_with_hash(dbg_ctx* ctx, const dbg_kmer_table* index, int stranded, int reduce_op,
                                 dbg_graph** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    NULLCHK(ctx, index);
    cudaSetDevice(ctx->c.device);
    Graph* g = nullptr;
    int rc = compress_dev(CTX(ctx), &index->t, stranded != 0, reduce_op, &g);
    *out = reinterpret_cast<dbg_graph*>(g);
    return rc;
}

uint64_t dbg_graph_len(const dbg_graph* g) { return g ? g->g.n_nodes : 0; }
uint64_t dbg_graph_n_bases(const dbg_graph* g) { return g ? g->g.n_bases : 0; }
uint64_t dbg_graph_n_words(const dbg_graph* g) { return g ? g->g.n_words : 0; }
int dbg_graph_stranded(const dbg_graph* g) { return g ? g->g.stranded : 0; }
int dbg_graph_k(const dbg_graph* g) { return g ? g->g.k : 0; }

int dbg_graph_copy_out(const dbg_graph* h, uint64_t* words, uint64_t* start, uint32_t* length, uint8_t* exts,
                       uint16_t* data) {
    if (!h) return DBG_E_BADARG;
    const Graph* g = &h->g;
    Ctx* c = g->ctx;
    cudaSetDevice(c->device);
    if (g->n_nodes) {
        if (words && g->n_words) CU(c, cudaMemcpyAsync(words, g->words, g->n_words * 8, cudaMemcpyDeviceToHost, c->stream));
        if (start) CU(c, cudaMemcpyAsync(start, g->start, g->n_nodes * 8, cudaMemcpyDeviceToHost, c->stream));
        if (length) CU(c, cudaMemcpyAsync(length, g->length, g->n_nodes * 4, cudaMemcpyDeviceToHost, c->stream));
        if (exts) CU(c, cudaMemcpyAsync(exts, g->exts, g->n_nodes, cudaMemcpyDeviceToHost, c->stream));
        if (data) CU(c, cudaMemcpyAsync(data, g->data, g->n_nodes * 2, cudaMemcpyDeviceToHost, c->stream));
    }
    return sync(c);
}

// bincode 1.x (fixed-width little-endian integers, u64 sequence lengths) image of the crate's serde-derived BaseGraph<K, u16>:
//   sequences.sequence.storage: Vec<u64> | sequences.sequence.len: usize | sequences.start: Vec<usize> | sequences.length: Vec<u32>
//   | exts: Vec<Exts{val: u8}> | data: Vec<u16> | stranded: bool | phantom: PhantomData<K> (no bytes)
// (field order: src/graph.rs:43-50, src/dna_string.rs:72-76, 762-767, src/lib.rs:577-580)
static u64 graph_bincode_size(const Graph* g) {
    return 8 + g->n_words * 8 + 8 + 8 + g->n_nodes * 8 + 8 + g->n_nodes * 4 + 8 + g->n_nodes + 8 + g->n_nodes * 2 + 1;
}
int dbg_graph_serialize(const dbg_graph* h, uint8_t* out, uint64_t cap, uint64_t* n_bytes) {
    if (!h || !n_bytes) return DBG_E_BADARG;
    const Graph* g = &h->g;
    Ctx* c = g->ctx;
    const u64 need = graph_bincode_size(g);
    *n_bytes = need;
    if (!out || cap < need) return DBG_OK;   // size query
    cudaSetDevice(c->device);
    u8* q = out;
    auto put64 = [&](u64 v) { memcpy(q, &v, 8); q += 8; };
    const u64 m = g->n_nodes;
    put64(g->n_words);
    u8* p_words = q; q += g->n_words * 8;
    put64(g->n_bases);
    put64(m);
    u8* p_start = q; q += m * 8;
    put64(m);
    u8* p_len = q; q += m * 4;
    put64(m);
    u8* p_exts = q; q += m;
    put64(m);
    u8* p_data = q; q += m * 2;
    *q++ = g->stranded ? 1 : 0;
    if (m) {
        if (g->n_words) CU(c, cudaMemcpyAsync(p_words, g->words, g->n_words * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(p_start, g->start, m * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(p_len, g->length, m * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(p_exts, g->exts, m, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(p_data, g->data, m * 2, cudaMemcpyDeviceToHost, c->stream));
    }
    return sync(c);
}
int dbg_graph_deserialize(dbg_ctx* ctx, int k, const uint8_t* bytes, uint64_t n_bytes, dbg_graph** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    Ctx* c = CTX(ctx);
    *out = nullptr;
    if (!bytes || k < 2 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "null buffer or k outside [2,64]");
    cudaSetDevice(c->device);
    const u8* q = bytes;
    const u8* end = bytes + n_bytes;
    bool bad = false;
    auto get64 = [&]() -> u64 { if ((u64)(end - q) < 8) { bad = true; return 0; } u64 v; memcpy(&v, q, 8); q += 8; return v; };
    auto skip = [&](u64 n, u64 es) -> const u8* { if (bad || n > (u64)(end - q) / es) { bad = true; return q; } const u8* r = q; q += n * es; return r; };
    const u64 nw = get64();
    const u8* p_words = skip(nw, 8);
    const u64 nb = get64();
    const u64 m1 = get64(); const u8* p_start = skip(m1, 8);
    const u64 m2 = get64(); const u8* p_len = skip(m2, 4);
    const u64 m3 = get64(); const u8* p_exts = skip(m3, 1);
    const u64 m4 = get64(); const u8* p_data = skip(m4, 2);
    if (bad || end - q < 1) DBG_SET_ERR(c, DBG_E_BADARG, "truncated BaseGraph image");
    const int stranded = *q++ != 0;
    if (m1 != m2 || m1 != m3 || m1 != m4 || nw != (nb + 31) / 32) DBG_SET_ERR(c, DBG_E_BADARG, "inconsistent BaseGraph image");
    // stage through device buffers and adopt them (graph_from_device_dev copies device -> device)
    DBuf<u64> dw, ds;
    DBuf<u32> dl;
    DBuf<u8> de;
    DBuf<u16> dd;
    TRY(dw.alloc_pool(c, nw + 1)); TRY(ds.alloc_pool(c, m1 + 1)); TRY(dl.alloc_pool(c, m1 + 1)); TRY(de.alloc_pool(c, m1 + 1)); TRY(dd.alloc_pool(c, m1 + 1));
    if (nw) CU(c, cudaMemcpyAsync(dw.p, p_words, nw * 8, cudaMemcpyHostToDevice, c->stream));
    if (m1) {
        CU(c, cudaMemcpyAsync(ds.p, p_start, m1 * 8, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(dl.p, p_len, m1 * 4, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(de.p, p_exts, m1, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(dd.p, p_data, m1 * 2, cudaMemcpyHostToDevice, c->stream));
    }
    TRY(sync(c));   // the source buffer is the caller's
    Graph* g = nullptr;
    TRY(graph_from_device_dev(c, k, stranded, m1, nb, dw.p, ds.p, dl.p, de.p, dd.p, &g));
    *out = reinterpret_cast<dbg_graph*>(g);
    return DBG_OK;
}

int dbg_graph_edges(dbg_ctx* ctx, const dbg_graph* graph, uint32_t* target, uint8_t* flags) {
    if (!ctx) return DBG_E_BADARG;
    NULLCHK(ctx, graph);
    cudaSetDevice(ctx->c.device);
    return graph_edges_dev(CTX(ctx), &graph->g, target, flags);
}

int dbg_graph_fix_exts(dbg_ctx* ctx, dbg_graph* graph, const uint8_t* valid_nodes) {
    if (!ctx) return DBG_E_BADARG;
    NULLCHK(ctx, graph);
    cudaSetDevice(ctx->c.device);
    return graph_fix_exts_dev(CTX(ctx), &graph->g, valid_nodes);
}

int dbg_graph_combine(dbg_ctx* ctx, const dbg_graph* const* graphs, uint32_t n_graphs, dbg_graph** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    NULLCHK(ctx, graphs);
    cudaSetDevice(ctx->c.device);
    std::vector<const Graph*> gs(n_graphs);
    for (uint32_t i = 0; i < n_graphs; i++) {
        if (!graphs[i]) DBG_SET_ERR(CTX(ctx), DBG_E_BADARG, "graphs[%u] is null", i);
        if (graphs[i]->g.ctx != CTX(ctx)) DBG_SET_ERR(CTX(ctx), DBG_E_BADARG, "graphs[%u] belongs to another context", i);
        gs[i] = &graphs[i]->g;
    }
    Graph* g = nullptr;
    int rc = graph_combine_dev(CTX(ctx), gs.data(), n_graphs, &g);
    if (rc == DBG_OK) *out = reinterpret_cast<dbg_graph*>(g);
    return rc;
}

int dbg_compress_graph(dbg_ctx* ctx, const dbg_graph* graph, int stranded, int reduce_op, const uint64_t* censor_nodes,
                       uint64_t n_censor, dbg_graph** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    NULLCHK(ctx, graph);
    cudaSetDevice(ctx->c.device);
    Graph* g = nullptr;
    int rc = compress_graph_dev(CTX(ctx), &graph->g, stranded, reduce_op, (const u64*)censor_nodes, n_censor, &g);
    if (rc == DBG_OK) *out = reinterpret_cast<dbg_graph*>(g);
    return rc;
}

int dbg_graph_is_compressed(dbg_ctx* ctx, const dbg_graph* graph, int scmap_join_test, int64_t* pair_out) {
    if (!ctx || !pair_out) return DBG_E_BADARG;
    NULLCHK(ctx, graph);
    cudaSetDevice(ctx->c.device);
    long long pr = -1;
    int rc = graph_is_compressed_dev(CTX(ctx), &graph->g, scmap_join_test, &pr);
    *pair_out = (int64_t)pr;
    return rc;
}

void dbg_graph_free(dbg_graph* g) { if (g) { cudaSetDevice(g->g.ctx->device); free_graph(&g->g); } }

// ---- fused ------------------------------------------------------------------------------------------------
int dbg_reads_to_graph(dbg_ctx* ctx, int k, const dbg_seqset* seqs, uint32_t min_kmer_obs, int stranded,
                       int reduce_op, dbg_kmer_table** table_out, dbg_graph** graph_out) {
    if (!ctx || !graph_out) return DBG_E_BADARG;
    *graph_out = nullptr;
    if (table_out) *table_out = nullptr;
    dbg_kmer_table* t = nullptr;
    int rc = dbg_filter_kmers(ctx, k, seqs, min_kmer_obs, stranded, 0, 0, &t);
    if (rc != DBG_OK) return rc;
    dbg_stats s1 = ctx->c.stats;
    rc = dbg_compress_kmers_with_hash(ctx, t, stranded, reduce_op, graph_out);
    // keep the filter-stage timings next to the compress-stage ones
    ctx->c.stats.ms_partition = s1.ms_partition; ctx->c.stats.ms_count = s1.ms_count; ctx->c.stats.ms_sort = s1.ms_sort;
    if (rc != DBG_OK || !table_out) dbg_table_free(t);
    else *table_out = t;
    return rc;
}

int dbg_reads_to_graph_host(dbg_ctx* ctx, int k, const uint64_t* words, uint64_t n_words, const uint64_t* start,
                            const uint32_t* length, const uint8_t* seq_exts, uint64_t n_seqs, uint32_t min_kmer_obs,
                            int stranded, int reduce_op, dbg_kmer_table** table_out, dbg_graph** graph_out) {
    if (!ctx || !graph_out) return DBG_E_BADARG;
    *graph_out = nullptr;
    dbg_seqset* s = nullptr;
    int rc = dbg_seqset_upload(ctx, words, n_words, start, length, seq_exts, n_seqs, &s);
    if (rc != DBG_OK) return rc;
    rc = dbg_reads_to_graph(ctx, k, s, min_kmer_obs, stranded, reduce_op, table_out, graph_out);
    dbg_seqset_free(s);
    return rc;
}

int dbg_reads_to_graph_host_uniform(dbg_ctx* ctx, int k, const uint64_t* words, uint64_t n_words, uint64_t n_seqs,
                                    uint32_t read_len, const uint8_t* seq_exts, uint32_t min_kmer_obs, int stranded,
                                    int reduce_op, dbg_kmer_table** table_out, dbg_graph** graph_out) {
    if (!ctx || !graph_out) return DBG_E_BADARG;
    *graph_out = nullptr;
    dbg_seqset* s = nullptr;
    // pipelined: the packed reads go up in chunks on the copy stream while the partition kernel already works on
    // the chunks that have arrived (the host buffer stays borrowed until this call returns)
    int rc = upload_uniform_impl(ctx, words, n_words, n_seqs, read_len, seq_exts, true, &s);
    if (rc != DBG_OK) return rc;
    rc = dbg_reads_to_graph(ctx, k, s, min_kmer_obs, stranded, reduce_op, table_out, graph_out);
    spin_sync(ctx->c.copy_stream);
    dbg_seqset_free(s);
    return rc;
}

int dbg_msp_kmer_buckets(dbg_ctx* ctx, int k, int p, const dbg_seqset* seqs, int stranded, uint32_t* out_bucket,
                         uint64_t n_out) {
    if (!ctx) return DBG_E_BADARG;
    NULLCHK(ctx, seqs);
    if (n_out && !out_bucket) DBG_SET_ERR(CTX(ctx), DBG_E_BADARG, "null output");
    cudaSetDevice(ctx->c.device);
    return msp_kmer_buckets_dev(CTX(ctx), k, p, &seqs->s, stranded != 0, out_bucket, n_out);
}

int dbg_msp_sequence(dbg_ctx* ctx, int k, int p, const dbg_seqset* seqs, int rc, const uint32_t* permutation, uint64_t cap,
                     uint64_t* n_intervals, uint32_t* seq, uint32_t* start, uint32_t* len, uint32_t* bucket, uint8_t* exts) {
    if (!ctx || !n_intervals) return DBG_E_BADARG;
    NULLCHK(ctx, seqs);
    cudaSetDevice(ctx->c.device);
    u64 n = 0;
    int rcode = msp_sequence_dev(CTX(ctx), k, p, &seqs->s, rc != 0, permutation, cap, &n, seq, start, len, bucket, exts);
    *n_intervals = n;
    return rcode;
}

int dbg_filter_kmers_colorset(dbg_ctx* ctx, int k, const dbg_seqset* seqs, const uint8_t* labels, uint32_t min_kmer_obs, int stranded,
                              uint64_t memory_size_gb, dbg_kmer_table** out) {
    if (!ctx || !out) return DBG_E_BADARG;
    *out = nullptr;
    NULLCHK(ctx, seqs);
    cudaSetDevice(ctx->c.device);
    Table* t = nullptr;
    int rc = filter_kmers_colorset_dev(CTX(ctx), k, &seqs->s, labels, min_kmer_obs, stranded, memory_size_gb, &t);
    if (rc == DBG_OK) *out = reinterpret_cast<dbg_kmer_table*>(t);
    return rc;
}
int dbg_table_colorsets(const dbg_kmer_table* h, uint64_t* masks) {
    if (!h || !masks) return DBG_E_BADARG;
    const Table* t = &h->t;
    Ctx* c = t->ctx;
    if (!t->colors && t->n) DBG_SET_ERR(c, DBG_E_BADARG, "the table was not built by dbg_filter_kmers_colorset");
    cudaSetDevice(c->device);
    if (t->n) CU(c, cudaMemcpyAsync(masks, t->colors, t->n * 8, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}

}  // extern "C"
