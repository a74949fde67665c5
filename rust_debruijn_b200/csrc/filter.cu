// filter::filter_kmers with CountFilter, rebuilt for B200 (replaces src/filter.rs:138-231).
//
// The reference materialises one 16-byte tuple per k-mer OCCURRENCE, scatters them into 256
// prefix buckets and merge-sorts each bucket (src/filter.rs:186-219).  Here the N occurrences are
// never written to HBM:
//
//   P1  msp_partition   2-bit p-mer scores, sliding-window minimum (minimum-substring partitioning, src/msp.rs:207-276
//                       with a hash permutation and rc = !stranded), maximal runs of k-mers with equal bucket -> one
//                       16/32-byte SUPER-K-MER record (bases + boundary Exts nibbles, Exts::from_slice_bounds
//                       src/lib.rs:645-660).  ~1.7 B per k-mer instead of 9-16.  Contiguous layouts: tile-per-CTA kernel
//                       (msp_tile_kernel); anything else: warp per sequence chunk (msp_partition_kernel).
//                       Direct mode (large device-resident inputs): a sampling pass over 1/16 of the tiles sizes one
//                       region per bucket, the main pass writes every record into its bucket's region.
//   P1b scatter         staging mode only (small inputs, multi-pass, pipelined uploads, the multi-GPU partition):
//                       records -> contiguous per-bucket ranges (histogram + exclusive scan + scatter).
//   P2  count           one CTA per bucket (persistent, atomic work queue): records are deduplicated, cut into
//                       length-sorted tasks and expanded with ROLLING fwd / reverse-complement k-mers (KmerExtsIter
//                       src/lib.rs:812-841, min_rc_flip :224-231, Exts::rc :746), inserted into an open-addressed
//                       table in SHARED memory: key CAS, Exts OR, count add — CountFilter::summarize
//                       (src/filter.rs:52-63).  A table overflow splits the bucket by hash class
//                       (always terminates: the class hash is a bijection of the key).
//   P3  sort            only the V valid (k-mer, exts, count) records are radix sorted (scan_sort.cu)
//                       to give the ascending order of src/filter.rs:205-219.
//
// Result arrays are bit-identical to the reference's valid_kmers / valid_exts / valid_data (and
// all_kmers) before BoomHashMap2::new permutes them: they are fully determined (ascending, unique).
#include "common.cuh"

namespace dbg {

static const int P1_WARPS = 8;
static const int P1_THREADS = P1_WARPS * 32;
static const int CHUNK = 512;        // k-mers per work item
static const int SC_N = CHUNK + 64;  // p-mer scores per item (K - p <= 63)
static const int SW_N = 24;          // staged 2-bit words per item: (CHUNK + 64 + 2)/32 + 3
static const int WCHUNK = 256;       // record slots a warp reserves per global atomic
static const u32 INVALID_BUCKET = 0xffffffffu;

struct P1Args {
    const u64* words; u64 n_words;
    const u64* start; const u32* length; const u8* seq_exts;
    u64 n_seqs; u32 uniform_len;
    const u32* item_seq; const u32* item_j0; u64 n_items;  // item_seq == nullptr => item i = sequence i, j0 = 0
    int p; int stranded; u32 bucket_mask; int maxk;
    u32 bk_lo, bk_span;  // this pass keeps records of buckets [bk_lo, bk_lo + bk_span) (multi-pass planner)
    u64* rec; u32* rec_bucket; u64 capacity;                 // staging
    u64* cursor; u32* bucket_count; u32* overflow;
    // tile kernel only.  mode 0: records -> staging + bucket ids (scattered afterwards); mode 1: bucket histogram of the
    // visited tiles only (sampling pass); mode 2: records straight into per-bucket regions [bucket_start[b], +bucket_cap[b])
    int mode;
    const u64* bucket_start; const u32* bucket_cap; u32* bucket_fill;
};

template <int W>
__global__ void __launch_bounds__(P1_THREADS) msp_partition_kernel(KP kp, P1Args a) {
    constexpr int RW = RecLayout<W>::WORDS;
    __shared__ u64 s_w[P1_WARPS][SW_N];
    __shared__ u32 s_sc[P1_WARPS][SC_N];
    __shared__ u32 s_bk[P1_WARPS][CHUNK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1;
    u64* sw = s_w[warp];
    u32* sc = s_sc[warp];
    u32* bk = s_bk[warp];
    const int K = kp.k, p = a.p, wlen = K - p + 1;
    u64 chunk_base = 0;
    u32 chunk_used = WCHUNK;  // forces a reservation on first use

    for (u64 item = (u64)blockIdx.x * P1_WARPS + warp; item < a.n_items; item += (u64)gridDim.x * P1_WARPS) {
        u64 s = a.item_seq ? a.item_seq[item] : item;
        u32 j0 = a.item_seq ? a.item_j0[item] : 0;
        u32 L = a.uniform_len ? a.uniform_len : a.length[s];
        u64 st = a.uniform_len ? s * (u64)a.uniform_len : a.start[s];
        if (L < (u32)K) continue;  // shorter than K: no k-mers (src/lib.rs:783,813)
        u32 nk = L - K + 1;
        u32 nkc = nk - j0 < (u32)CHUNK ? nk - j0 : (u32)CHUNK;
        u32 sx = a.seq_exts ? a.seq_exts[s] : 0;
        u32 lo_b = j0 > 0 ? j0 - 1 : 0;   // first staged base (left flank if any)
        u32 rel = j0 - lo_b;              // staged offset of k-mer j0's first base
        u32 nb = nkc + K + 1;             // enough for right flank (may run past L: masked by logic below)
        // ---- stage 2-bit words of this chunk, re-aligned so staged base 0 = sequence base lo_b ----
        for (u32 t = lane; t < (nb + 31) / 32 + 2 && t < (u32)SW_N; t += 32) {
            u64 b = st + lo_b + 32ull * t;
            u64 wi = b >> 5;
            int sh = (int)(b & 31) * 2;
            u64 hi = wi < a.n_words ? a.words[wi] : 0;
            u64 lo = (wi + 1) < a.n_words ? a.words[wi + 1] : 0;
            sw[t] = sh ? (hi << sh) | (lo >> (64 - sh)) : hi;
        }
        __syncwarp();
        // ---- p-mer scores ----
        u32 np = nkc + K - p;
        for (u32 q = lane; q < np; q += 32) {
            u32 x = (u32)(bases64(sw, rel + q) >> (64 - 2 * p));
            sc[q] = pmer_score(x, p, a.stranded != 0);
        }
        __syncwarp();
        // ---- sliding-window minimum -> bucket of every k-mer (msp.rs:207-248, order-free form) ----
        for (u32 j = lane; j < nkc; j += 32) {
            u32 m = sc[j];
            for (int t = 1; t < wlen; t++) m = min(m, sc[j + t]);
            bk[j] = m & a.bucket_mask;
        }
        __syncwarp();
        // ---- runs of equal bucket (capped at maxk k-mers) -> records ----
        int last_start = 0;
        for (u32 g0 = 0; g0 <= nkc; g0 += 32) {
            u32 j = g0 + lane;
            bool change = (j < nkc) && (j > 0) && (bk[j] != bk[j - 1]);
            u32 nat = __ballot_sync(0xffffffffu, change);
            int limit = (int)min(g0 + 32, nkc);
            u32 all = 0;
            int cur = last_start;
            for (;;) {  // warp-uniform: merge natural starts with forced (length cap) starts
                int nx = nat ? (int)g0 + __ffs(nat) - 1 : limit;
                while (nx - cur > a.maxk) { cur += a.maxk; all |= 1u << (cur - (int)g0); }
                if (!nat) break;
                cur = nx;
                all |= 1u << (nx - (int)g0);
                nat &= nat - 1;
            }
            if (limit == (int)nkc && nkc - g0 < 32) all |= 1u << (nkc - g0);  // virtual start closes the tail
            if (all) {
                u32 cnt = __popc(all);
                if (chunk_used + cnt > (u32)WCHUNK) {
                    u64 cb = 0;
                    if (lane == 0) cb = atomicAdd(a.cursor, (u64)WCHUNK);
                    chunk_base = __shfl_sync(0xffffffffu, cb, 0);
                    chunk_used = 0;
                    if (chunk_base + WCHUNK > a.capacity && lane == 0) *a.overflow = 1;
                }
                if ((all >> lane) & 1u) {
                    u32 lower = all & lt;
                    int prev = lower ? (int)g0 + 31 - __clz(lower) : last_start;
                    int n = (int)j - prev;              // 1..maxk k-mers: [prev, j)
                    u32 off = rel + prev;               // staged offset of the run's first base
                    u32 nbase = n + K - 1;
                    u32 a_pos = j0 + prev;              // sequence position of first base
                    u32 e_pos = a_pos + nbase;          // sequence position of the base after the run
                    u32 ln = a_pos == 0 ? (sx & 0xfu) : (1u << base_at(sw, off - 1));
                    u32 rn = e_pos >= L ? (sx >> 4) & 0xfu : (1u << base_at(sw, off + nbase));
                    u64 hdr = ((u64)n << 8) | (rn << 4) | ln;
                    u64 slot = chunk_base + chunk_used + __popc(lower);
                    if (slot < a.capacity) {
                        u64 r[RW];
#pragma unroll
                        for (int t = 0; t < RW; t++) r[t] = (32u * t < nbase) ? bases64(sw, off + 32 * t) : 0;
                        // clear everything past the last base, then drop the header into the low 14 bits
                        int lastw = (int)((nbase - 1) >> 5);
                        int used = (int)(nbase - 32 * lastw);  // 1..32 bases in the last occupied word
#pragma unroll
                        for (int t = 0; t < RW; t++) {
                            if (t == lastw && used < 32) r[t] &= ~0ull << (64 - 2 * used);
                            if (t > lastw) r[t] = 0;
                        }
                        r[RW - 1] |= hdr;
                        u32 b = bk[prev];
                        if (b - a.bk_lo < a.bk_span) {  // other buckets belong to another pass: the slot stays INVALID
                            if constexpr (RW == 2) {
                                *reinterpret_cast<ulonglong2*>(a.rec + slot * 2) = make_ulonglong2(r[0], r[1]);
                            } else {
                                *reinterpret_cast<ulonglong2*>(a.rec + slot * 4) = make_ulonglong2(r[0], r[1]);
                                *reinterpret_cast<ulonglong2*>(a.rec + slot * 4 + 2) = make_ulonglong2(r[2], r[RW - 1]);
                            }
                            a.rec_bucket[slot] = b;
                            atomicAdd(&a.bucket_count[b], 1u);
                        }
                    }
                }
                chunk_used += cnt;
                last_start = (int)g0 + 31 - __clz(all);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// P1 fast path: partition of CONTIGUOUSLY packed sequences (PackedDnaStringSet::add layout:
// start[i+1] = start[i] + length[i]; uniform-length reads are the arithmetic special case).
//
// Round-2 form: no CTA-wide barrier in the steady state.  A CTA is 8 consumer warps + 1 producer warp.
// The producer streams BLOCKS of packed words (8 warp tiles, ~1 KB) from HBM into a 4-deep
// shared-memory ring with bulk asynchronous copies (cp.async.bulk = TMA, completion on the slot's
// "full" mbarrier) and waits on the slot's "empty" mbarrier in between.  The consumer warps CLAIM warp
// tiles from a ticket counter and process each on their own, synchronising with __syncwarp only; a
// finished tile arrives on the slot's "empty" mbarrier.  Warps of a CTA drift up to 3 blocks apart,
// warps of different CTAs are independent: the phases below (which used to be separated by
// __syncthreads and left half the issue slots idle) overlap freely across the 40 resident consumer
// warps of an SM.
//
// A warp tile EXAMINES 512 consecutive k-mer start positions (16 per lane) and OWNS the first
// tpw = 512 - (w - 1) of them (w = K - p + 1 p-mers per window): the 512 p-mer scores it computes are
// then exactly the ones its owned windows need, so no halo is ever recomputed.
//   A  scores     lane = 4 consecutive p-mers x 4 rounds: 3 LDS.32 + funnel shifts, canonical + hash, STS.128
//   B  buckets    lane = 4 consecutive k-mers x 4 rounds: register-tiled window minimum
//   C  runs       lane = 16 positions: validity / sequence-start / bucket-change masks, run starts and
//                 closers by bit tricks + two warp scans; closed runs go to the warp's queue
//   D  records    lane = one queued run: slot from the bucket's cursor (direct mode) or from the warp's
//                 staging chunk, 16/32-byte record assembled from the staged bases, one vector store
// ------------------------------------------------------------------------------------------------
static const int WT_CW = 8;                          // consumer warps per CTA
static const int WT_THREADS = (WT_CW + 1) * 32;      // + 1 producer warp
static const int WT_CTAS = 5;                        // resident CTAs per SM (42.6 KB static shared memory each)
static const int WP = 512;                           // positions a warp tile examines
static const int WSC = WP + 64 + 8;                  // p-mer scores per warp tile (+ what the un-owned windows may touch)
static const int WBK = WP + WP / 16 + 8;             // bucket of every examined position, padded (bkpad)
static const int WBM = 24;                           // bitmap words per warp tile (>= (31 + WP + 64 + 2) / 32 + 3)
static const int NST = 4;                            // ring depth
#ifndef W2_DIRECT
#define W2_DIRECT 1
#endif
static const int SB_BYTES = 1088;                    // bytes staged per block (8 warp tiles + flanks), multiple of 16
static const int SBW = 288;                          // u32 words per ring slot
static const int TP = 4096;                          // (size unit of the direct_min_tiles parameter only)

struct TileArgs {
    u64 base0;      // global position of the first base (start[0])
    u64 total_end;  // global position one past the last base
    u64 tile0;      // first block of this launch
    u64 n_tiles;    // one past the last block of this launch
    u32 tstride;    // visit every tstride-th block (1 = all; > 1 = sampling pass)
};
// positions one block covers (host and device agree on this): 8 warp tiles of 512 - (w - 1) owned positions
__host__ __device__ __forceinline__ u32 tile_owned(int k, int p) { return (u32)(WP - (k - p)); }

// ---- mbarrier + bulk-copy primitives (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, u32 bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u32 addr, u32 parity) {
    u32 ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// consumer side: the data is almost always there already (the ring runs 3 blocks ahead)
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    const u32 addr = smem_u32(bar);
    while (!mbar_try_wait(addr, parity)) __nanosleep(64);
}
// producer side: the wait for a free slot lasts a whole block; the suspend-time hint lets the hardware park the warp
__device__ __forceinline__ void mbar_wait_suspended(u64* bar, u32 parity) {
    const u32 addr = smem_u32(bar);
    u32 ok = 0;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
        // (12-15 % of the kernel's executed instructions are this loop; an added __nanosleep between polls changed neither that
        // share nor the kernel time: the polls take issue slots nobody else wanted)
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"((u64)__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The ring slots hold the packed u64 words exactly as they lie in HBM; read as u32, the word of BASE-ORDER index j
// (16 bases each, first base in the top bits) sits at u32 index j ^ 1 (little-endian halves of a u64).
__device__ __forceinline__ u32 st_word(const u32* st, u32 j) { return st[j ^ 1u]; }
__device__ __forceinline__ u32 st_bits(const u32* st, u32 b) {  // 32 bits starting at staged base b
    const u32 i = b >> 4;
    return __funnelshift_l(st_word(st, i + 1), st_word(st, i), 2 * (b & 15));
}
__device__ __forceinline__ u32 st_base(const u32* st, u32 b) { return (st_word(st, b >> 4) >> (30 - 2 * (b & 15))) & 3u; }
__device__ __forceinline__ u32 bm_bit(const u32* bm, u32 b) { return (bm[b >> 5] >> (b & 31)) & 1u; }
__device__ __forceinline__ u32 bm_bits32(const u32* bm, u32 b) {  // 32 bitmap bits starting at index b (LSB = b)
    u32 i = b >> 5;
    return __funnelshift_r(bm[i], bm[i + 1], b & 31);
}
__device__ __forceinline__ u64 seq_lower_bound(const u64* __restrict__ start, u64 n, u64 g) {
    u64 lo = 0, hi = n;
    while (lo < hi) { u64 m = (lo + hi) >> 1; if (start[m] < g) lo = m + 1; else hi = m; }
    return lo;
}
// index of the sequence holding base g: last i with start[i] <= g (skips zero-length sequences sharing a start)
__device__ __forceinline__ u64 seq_index_of(const u64* __restrict__ start, u64 n, u64 g) {
    u64 lo = 0, hi = n;
    while (lo < hi) { u64 m = (lo + hi) >> 1; if (start[m] <= g) lo = m + 1; else hi = m; }
    return lo - 1;
}

// One record: run of nn k-mers starting at tile position ps.  st = ring slot (staged origin sb, global), ofs = staged
// index of tile position 0, bm = the warp's boundary bitmap whose bit 0 is staged index wo.
template <int W>
__device__ __forceinline__ void tile_emit_record(const P1Args& a, const TileArgs& ta, int K, const u32* st, const u32* bm, u32 wo,
                                                 u32 bkt, u64 sb, u32 ofs, int ps, int nn, u64 slot, bool ok) {
    // `slot` / `ok` may depend on an atomic that is still in flight (direct partition): nothing below touches them
    // until the final store, so the record is assembled while the atomic travels
    constexpr int RW = RecLayout<W>::WORDS;
    const u32 eb = ofs + (u32)ps;            // staged index of the run's first base
    const u32 nbase = (u32)nn + K - 1;
    const u64 gpos = sb + eb;
    const bool at_first = bm_bit(bm, eb - wo) != 0;                                              // run starts a sequence
    const bool at_last = bm_bit(bm, eb + nbase - wo) != 0 || (gpos + nbase >= ta.total_end);   // run ends a sequence
    u32 ln, rn;
    if (at_first) {  // KmerExtsIter: first k-mer takes the sequence-level left nibble (lib.rs:820-824)
        u32 sx = 0;
        if (a.seq_exts) {
            u64 si = a.uniform_len ? (gpos - ta.base0) / a.uniform_len : seq_index_of(a.start, a.n_seqs, gpos);
            sx = a.seq_exts[si];
        }
        ln = sx & 0xfu;
    } else {
        ln = 1u << st_base(st, eb - 1);
    }
    if (at_last) {   // last k-mer takes the sequence-level right nibble (lib.rs:826-830)
        u32 sx = 0;
        if (a.seq_exts) {
            u64 ge = gpos + nbase;
            u64 si = a.uniform_len ? (ge - 1 - ta.base0) / a.uniform_len : seq_index_of(a.start, a.n_seqs, ge - 1);
            sx = a.seq_exts[si];
        }
        rn = (sx >> 4) & 0xfu;
    } else {
        rn = 1u << st_base(st, eb + nbase);
    }
    const u64 hdr = ((u64)nn << 8) | (rn << 4) | ln;
    u64 r[RW];
    {   // the record's 2 * RW + 1 staged words are loaded once each, then realigned in registers
        const u32 j0 = eb >> 4, sh = 2 * (eb & 15);
        u32 wv[2 * RW + 1];
#pragma unroll
        for (int t = 0; t < 2 * RW + 1; t++) wv[t] = st_word(st, j0 + t);
#pragma unroll
        for (int t = 0; t < RW; t++)
            r[t] = ((u64)__funnelshift_l(wv[2 * t + 1], wv[2 * t], sh) << 32) | __funnelshift_l(wv[2 * t + 2], wv[2 * t + 1], sh);
    }
    const int lastw = (int)((nbase - 1) >> 5);
    const int used = (int)(nbase - 32 * lastw);
#pragma unroll
    for (int t = 0; t < RW; t++) {
        if (t == lastw && used < 32) r[t] &= ~0ull << (64 - 2 * used);
        if (t > lastw) r[t] = 0;
    }
    r[RW - 1] |= hdr;
    if (!ok) {
        *a.overflow = 1;   // region / staging buffer too small: record dropped, the host retries (staging) or falls back to staging (direct)
        return;
    }
    if constexpr (RW == 2) {
        *reinterpret_cast<ulonglong2*>(a.rec + slot * 2) = make_ulonglong2(r[0], r[1]);
    } else {
        *reinterpret_cast<ulonglong2*>(a.rec + slot * 4) = make_ulonglong2(r[0], r[1]);
        *reinterpret_cast<ulonglong2*>(a.rec + slot * 4 + 2) = make_ulonglong2(r[2], r[RW - 1]);
    }
    if (a.mode == 0) {
        a.rec_bucket[slot] = bkt;
        atomicAdd(&a.bucket_count[bkt], 1u);
    }
}

__device__ __forceinline__ u32 bkpad(u32 x) { return x + (x >> 4); }  // padded index: lane stride 17 words, conflict-free

template <int W>
__global__ void __launch_bounds__(WT_THREADS, WT_CTAS) msp_tile_kernel(KP kp, P1Args a, TileArgs ta) {
    __shared__ __align__(128) u32 s_stage[NST][SBW];
    __shared__ __align__(16) u32 s_sc_all[WT_CW][WSC];
    __shared__ u32 s_bk_all[WT_CW][WBK];
    __shared__ u32 s_bm_all[WT_CW][WBM], s_vm_all[WT_CW][WBM];
    __shared__ __align__(8) u64 s_full[NST];   // per ring slot: "the block's bytes have landed" (1 arrival + transaction bytes)
    __shared__ __align__(8) u64 s_empty[NST];  // per ring slot: "all 8 warp tiles of the block are finished" (8 arrivals)
    __shared__ u32 s_ticket;                   // next warp tile of this CTA to be processed
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = kp.k, p = a.p, wlen = K - p + 1;
    const u32 tpw = tile_owned(K, p);          // positions a warp tile owns
    const u64 bp = (u64)WT_CW * tpw;           // positions per block
    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WT_CW); }
        s_ticket = 0;
        mbar_fence_init();
    }
    __syncthreads();   // the only CTA-wide barrier: the mbarriers exist
    const u64 blk_step = (u64)gridDim.x * ta.tstride;
    const u64 blk_first = ta.tile0 + (u64)blockIdx.x * ta.tstride;
    const bool tma_ok = (reinterpret_cast<unsigned long long>(a.words) & 15ull) == 0;

    if (warp == WT_CW) {
        // ---- producer warp: block i of this CTA -> ring slot i % NST by one bulk asynchronous copy (TMA) that completes
        // on the slot's "full" mbarrier, issued as soon as the slot's previous block was released ("empty" mbarrier).
        // Measured alternatives without a producer warp (the finisher of a block's last tile refills a slot, elected by
        // a shared counter or by tile number) ran 5-14% slower: the dedicated warp keeps the ring one block deeper and
        // its polling only takes issue slots nobody else wanted. ----
        u32 i = 0;
        for (u64 blk = blk_first; blk < ta.n_tiles; blk += blk_step, i++) {
            const u32 s = i % NST;
            if (i >= (u32)NST) mbar_wait_suspended(&s_empty[s], ((i / NST) - 1) & 1);
            const u64 g0B = ta.base0 + blk * bp;
            const u64 sbB = (g0B > ta.base0 ? g0B - 1 : g0B) & ~63ull;   // 64 bases = 16 bytes: bulk-copy alignment
            const u64 ow = sbB >> 5;                                      // first packed word of the block
            const u64 avail = ow < a.n_words ? (a.n_words - ow) * 8 : 0;
            const u32 copy = tma_ok ? (u32)min((u64)SB_BYTES, avail & ~15ull) : 0u;
            // whatever the bulk copy cannot take (end of the buffer, unaligned caller memory) goes through registers
            u64* dst = reinterpret_cast<u64*>(s_stage[s]);
            for (u32 j = copy / 8 + lane; j < (u32)SB_BYTES / 8; j += 32) dst[j] = (ow + j) < a.n_words ? a.words[ow + j] : 0ull;
            __syncwarp();
            if (lane == 0) {
                if (copy) {
                    mbar_arrive_expect_tx(&s_full[s], copy);
                    bulk_g2s(dst, a.words + ow, copy, &s_full[s]);
                } else {
                    mbar_arrive(&s_full[s]);
                }
            }
        }
        return;
    }

    // ---- consumer warps ----
    u32* const sc = s_sc_all[warp];
    u32* const bk = s_bk_all[warp];
    u32* const bm = s_bm_all[warp];   // bit b: a sequence starts at staged index wo + b (or the data ends there)
    u32* const vm = s_vm_all[warp];   // bit b: a valid k-mer starts at staged index wo + b
    u32* const queue = sc;            // closed runs (start | n << 16); reuses the score array after phase B
    const bool part = a.bk_span <= a.bucket_mask || a.maxk < 17;   // multi-pass planner: only some buckets belong to this pass
    u64 chunk_base = 0;               // staging mode: the warp's current chunk of record slots
    u32 chunk_used = 0, chunk_cap = 0;
    // Warp tiles are CLAIMED, not assigned: one ticket counter per CTA, ticket = 8 * (block of this CTA) + (warp tile of
    // that block), handed out in order.  A warp that is ahead simply takes more tiles and nobody waits for the slowest
    // warp of the CTA (with a fixed warp -> tile map the fast warps spun on the ring: 19% of the executed instructions).
    // A warp holding a ticket of block i keeps that block's ring slot from being released, so the slot's "full" barrier
    // can be at most one phase ahead of what the warp waits for.
    for (;;) {
        u32 ticket = 0;
        if (lane == 0) ticket = atomicAdd(&s_ticket, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        const u32 i = ticket >> 3, wt = ticket & 7u;                   // block of this CTA, warp tile of the block
        const u64 blk = blk_first + (u64)i * blk_step;
        if (blk >= ta.n_tiles) break;
        const u32 s = i % NST, use = i / NST;
        const u64 g0B = ta.base0 + blk * bp;
        const u64 sbB = (g0B > ta.base0 ? g0B - 1 : g0B) & ~63ull;    // global position of staged index 0
        const u64 g0 = g0B + (u64)wt * tpw;                            // global position of this warp tile's x = 0
        mbar_wait(&s_full[s], use & 1);
        if (g0 < ta.total_end) {
            const u32* st = s_stage[s];
            const u32 ofs = (u32)(g0 - sbB);                           // staged index of x = 0
            const u32 wo = (g0 > ta.base0 ? ofs - 1 : ofs) & ~31u;     // staged index of bitmap bit 0 (covers the left flank)
            const u32 nbits = (ofs - wo) + WP + K + 2;                 // bitmap bits that may be touched
            const u64 gw = sbB + wo;                                   // global position of bitmap bit 0
            // ---- sequence-boundary bitmap ----
            if (lane < WBM) bm[lane] = 0;
            __syncwarp();
            if (a.uniform_len) {
                const u32 L = a.uniform_len;
                const u64 i0 = gw <= ta.base0 ? 0 : (gw - ta.base0 + L - 1) / L;
                for (u64 q = i0 + lane; q <= a.n_seqs; q += 32) {
                    const u64 g = ta.base0 + q * L;
                    if (g >= gw + nbits) break;
                    atomicOr(&bm[(g - gw) >> 5], 1u << ((g - gw) & 31));
                }
            } else {
                const u64 i0 = seq_lower_bound(a.start, a.n_seqs, gw);
                for (u64 q = i0 + lane; q < a.n_seqs; q += 32) {
                    const u64 g = a.start[q];
                    if (g >= gw + nbits) break;
                    atomicOr(&bm[(g - gw) >> 5], 1u << ((g - gw) & 31));
                }
                if (lane == 0 && ta.total_end >= gw && ta.total_end < gw + nbits)
                    atomicOr(&bm[(ta.total_end - gw) >> 5], 1u << ((ta.total_end - gw) & 31));
            }
            // ---- phase A: p-mer scores of the 512 examined positions, 4 consecutive positions per lane and round ----
#pragma unroll
            for (int it = 0; it < WP / 128; it++) {
                const u32 q0 = 128 * it + 4 * lane;
                const u32 b = ofs + q0, j = b >> 4, sh = 2 * (b & 15);
                const u32 w0 = st_word(st, j), w1 = st_word(st, j + 1), w2 = st_word(st, j + 2);
                const u32 hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
                uint4 v;
                v.x = pmer_score(hi >> (32 - 2 * p), p, a.stranded != 0);
                v.y = pmer_score(__funnelshift_l(lo, hi, 2) >> (32 - 2 * p), p, a.stranded != 0);
                v.z = pmer_score(__funnelshift_l(lo, hi, 4) >> (32 - 2 * p), p, a.stranded != 0);
                v.w = pmer_score(__funnelshift_l(lo, hi, 6) >> (32 - 2 * p), p, a.stranded != 0);
                *reinterpret_cast<uint4*>(&sc[q0]) = v;
            }
            __syncwarp();
            // ---- validity bitmap: a k-mer at b is valid iff no boundary in (b, b+K) and b lies before the end of data.
            // OR over the K-1 following boundary bits by doubling on a 96-bit register window. ----
            if (lane < WBM - 3) {
                const u32 j = lane;
                const u32 w0 = bm[j], w1 = bm[j + 1], w2 = bm[j + 2];
                u32 y0 = __funnelshift_r(w0, w1, 1), y1 = __funnelshift_r(w1, w2, 1), y2 = w2 >> 1;
                int have = 1;            // y covers offsets 1 .. have
                const int need = K - 1;  // offsets 1 .. K-1
                while (have * 2 <= need) {
                    const int sft = have;      // < 32
                    const u32 z0 = __funnelshift_r(y0, y1, sft), z1 = __funnelshift_r(y1, y2, sft), z2 = y2 >> sft;
                    y0 |= z0; y1 |= z1; y2 |= z2;
                    have *= 2;
                }
                if (have < need) {
                    const int sft = need - have;  // < have <= 32
                    const u32 z0 = __funnelshift_r(y0, y1, sft), z1 = __funnelshift_r(y1, y2, sft);
                    y0 |= z0; y1 |= z1;
                }
                const u32 inv = need > 0 ? y0 : 0;
                const u64 endi = ta.total_end - gw;  // bitmap index of the end of data
                const u32 in_range = (u64)32 * j + 32 <= endi ? 0xffffffffu : ((u64)32 * j >= endi ? 0u : ((1u << (endi - 32 * j)) - 1));
                vm[j] = ~inv & in_range;
            }
            // ---- phase B: window minimum (w >= 4) for 4 consecutive k-mers per lane and round ----
#pragma unroll 1
            for (int it = 0; it < WP / 128; it++) {
                const u32 x0 = 128 * it + 4 * lane;
                const uint4 f = *reinterpret_cast<const uint4*>(&sc[x0]);
                u32 c = f.w;  // running min over indices 3 .. wlen-1
                int e = 4;
                for (; e + 3 < wlen; e += 4) {
                    const uint4 v = *reinterpret_cast<const uint4*>(&sc[x0 + e]);
                    c = min(min(c, v.x), min(v.y, min(v.z, v.w)));
                }
                for (; e < wlen; e++) c = min(c, sc[x0 + e]);
                const u32 t0 = sc[x0 + wlen], t1 = sc[x0 + wlen + 1], t2 = sc[x0 + wlen + 2];
                const u32 pb = bkpad(x0);  // x0 % 4 == 0: the four padded indices are consecutive
                bk[pb] = min(min(f.x, f.y), min(f.z, c)) & a.bucket_mask;
                bk[pb + 1] = min(min(f.y, f.z), min(c, t0)) & a.bucket_mask;
                bk[pb + 2] = min(min(f.z, c), min(t0, t1)) & a.bucket_mask;
                bk[pb + 3] = min(min(c, t0), min(t1, t2)) & a.bucket_mask;
            }
            __syncwarp();
            // ---- phase C: runs.  The lane owns 16 consecutive positions and works on 16/17-bit masks: valid,
            // first-of-sequence, bucket-differs.  A set bit in `smask` starts a run or (first invalid position after a
            // run) closes one.  Positions >= tpw belong to the next warp tile: invalid here, which closes the last run. ----
            u32 wtotal, nrec, qbase;
            u32 smask, cmask;
            int carry;
            const u32 x0 = 16 * lane;
            {
                const u32 b0 = ofs + x0 - wo;
                const u32 own = tpw > x0 ? (tpw - x0 >= 16 ? 0xffffu : (1u << (tpw - x0)) - 1) : 0u;
                const u32 vmk = bm_bits32(vm, b0) & own;
                const u32 fm = bm_bits32(bm, b0) & 0xffffu;
                u32 dm = 0;
                {
                    u32 prevb = bk[bkpad(x0 ? x0 - 1 : 0)];
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        const u32 cur = bk[bkpad(x0 + t)];
                        dm |= (cur != prevb ? 1u : 0u) << t;
                        prevb = cur;
                    }
                }
                const u32 up = __shfl_up_sync(0xffffffffu, vmk, 1);
                const u32 pv = (vmk << 1) | (lane ? (up >> 15) & 1u : 0u);           // bits 0..16: valid(x-1)
                const u32 start = vmk & (fm | ~pv | dm | (lane == 0 ? 1u : 0u));
                const u32 closer = ~vmk & pv & (lane == 31 ? 0x1ffffu : 0xffffu);    // bit 16 (x = 512) only for the last lane
                smask = start | closer;
                cmask = smask & pv;
                // carry: position of the last set bit of smask in any lower lane (only read when a run is open there)
                const int mylast = smask ? (int)x0 + 31 - __clz(smask) : -1;
                const u32 below = __ballot_sync(0xffffffffu, smask != 0) & ((1u << lane) - 1);
                carry = __shfl_sync(0xffffffffu, mylast, below ? 31 - __clz(below) : 0);
                if (!below) carry = 0;
                // records this lane will push (runs longer than maxk are cut)
                if (!part) {
                    // one pass over all buckets.  maxk >= 26 > 16 positions per lane: only a run that began in a lower lane
                    // can be longer than maxk, and only the lane's first set bit can close it
                    nrec = (u32)__popc(cmask);
                    if (cmask & (smask & (0u - smask))) {
                        const int n = (int)x0 + __ffs(smask) - 1 - carry;
                        if (n > a.maxk) nrec += (u32)((n - 1) / a.maxk);
                    }
                } else {
                    nrec = 0;
                    u32 m = cmask;
                    while (m) {
                        const int bit = __ffs(m) - 1;
                        m &= m - 1;
                        const u32 lower = smask & ((1u << bit) - 1);
                        const int prev = lower ? (int)x0 + 31 - __clz(lower) : carry;
                        const int n = (int)x0 + bit - prev;
                        if (bk[bkpad((u32)prev)] - a.bk_lo < a.bk_span)   // runs of other buckets belong to another pass
                            nrec += n <= a.maxk ? 1u : (u32)((n + a.maxk - 1) / a.maxk);
                    }
                }
                u32 inc = nrec;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                wtotal = __shfl_sync(0xffffffffu, inc, 31);
                qbase = inc - nrec;
            }
            __syncwarp();   // every lane is done with the scores' last readers (phase B) before the queue overwrites them
            {
                u32 qpos = qbase;
                u32 m = cmask;
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const u32 lower = smask & ((1u << bit) - 1);
                    const int prev = lower ? (int)x0 + 31 - __clz(lower) : carry;
                    const int xe = (int)x0 + bit;
                    if (part && bk[bkpad((u32)prev)] - a.bk_lo >= a.bk_span) continue;
                    for (int ps = prev; ps < xe; ps += a.maxk) {
                        queue[qpos] = (u32)ps | ((u32)min(a.maxk, xe - ps) << 16);   // qpos < wtotal <= WP: one run per position at most
                        qpos++;
                    }
                }
            }
            __syncwarp();
            // ---- phase D: one lane per record.  mode 0: slots from the warp's staging chunk (a new chunk is reserved from
            // the global cursor when the current one cannot take this tile's records; unused slots stay INVALID and are
            // skipped by the scatter pass); mode 2: the record's slot comes from its bucket's own cursor (no staging, no
            // scatter pass); mode 1: histogram only ----
            u64 slot0 = 0;
            bool chunk_ok = true;
            if (a.mode == 0 && wtotal) {
                if (chunk_used + wtotal > chunk_cap) {
                    chunk_cap = max((u32)WCHUNK, wtotal);
                    u64 cb = 0;
                    if (lane == 0) cb = atomicAdd(a.cursor, (u64)chunk_cap);
                    chunk_base = __shfl_sync(0xffffffffu, cb, 0);
                    chunk_used = 0;
                }
                slot0 = chunk_base + chunk_used;
                chunk_used += wtotal;
                chunk_ok = chunk_base + chunk_cap <= a.capacity;
            }
            // two records per lane and round: both bucket-cursor atomics are in flight before either record is assembled
            for (u32 q = lane; q < wtotal; q += 64) {
                const u32 q1 = q + 32;
                const bool has1 = q1 < wtotal;
                const u32 ent0 = queue[q], ent1 = has1 ? queue[q1] : 0u;
                const int ps0 = (int)(ent0 & 0xffffu), nn0 = (int)(ent0 >> 16);
                const int ps1 = (int)(ent1 & 0xffffu), nn1 = (int)(ent1 >> 16);
                const u32 bkt0 = bk[bkpad((u32)ps0)], bkt1 = bk[bkpad((u32)ps1)];
                if (a.mode == 1) {
                    atomicAdd(&a.bucket_count[bkt0], 1u);
                    if (has1) atomicAdd(&a.bucket_count[bkt1], 1u);
                    continue;
                }
                u64 slotA = slot0 + q, slotB = slot0 + q1;
                bool okA = chunk_ok, okB = chunk_ok;
                if (a.mode == 2) {
                    const u32 r0 = atomicAdd(&a.bucket_fill[bkt0], 1u);   // no branch on r0 / r1 here: see tile_emit_record
                    const u32 r1 = has1 ? atomicAdd(&a.bucket_fill[bkt1], 1u) : 0u;
                    slotA = a.bucket_start[bkt0] + r0;
                    okA = r0 < a.bucket_cap[bkt0] && slotA < a.capacity;
                    slotB = a.bucket_start[bkt1] + r1;
                    okB = r1 < a.bucket_cap[bkt1] && slotB < a.capacity;
                }
                tile_emit_record<W>(a, ta, K, st, bm, wo, bkt0, sbB, ofs, ps0, nn0, slotA, okA);
                if (has1) tile_emit_record<W>(a, ta, K, st, bm, wo, bkt1, sbB, ofs, ps1, nn1, slotB, okB);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);   // this warp tile no longer reads ring slot s
    }
}

// direct partition, between the sampling pass and the main pass: region capacity of every bucket from its sampled
// count (estimate + 6 sigma of the thinning noise + slack)
__global__ void bucket_caps_kernel(const u32* __restrict__ sample, u32 nb, float scale, u32* __restrict__ cap) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    // The sample saw s of the bucket's T records (s ~ Poisson(T / scale)).  The capacity is scale x an UPPER confidence bound of
    // the Poisson mean given s (z = 6.5: about 1e-10 per bucket), not "estimate + 6 sigma of the estimate": a bucket whose sample
    // came out low has a low estimate AND a low sigma, and with 2^18 buckets one such bucket per call is the rule, not the
    // exception (measured: region 109792, 331 records, 4 sampled of 22 expected -> capacity 325 -> the whole partition redone by
    // the staging path, 18 ms instead of 8 on one of four ranks).
    const float z = 6.5f, sf = (float)sample[b];
    const float lam = sf + z * sqrtf(sf + 0.25f * z * z) + 0.5f * z * z;
    cap[b] = (u32)(lam * scale) + 64u;
}
// after the main pass: records actually stored per bucket (cursor clipped to the capacity) and their total
__global__ void __launch_bounds__(256) bucket_fill_final_kernel(const u32* __restrict__ fill, const u32* __restrict__ cap,
                                                                const u64* __restrict__ start, u64 bound, u32 nb,
                                                                u32* __restrict__ cnt, u64* __restrict__ total) {
    __shared__ u64 s_w[8];
    u64 v = 0;
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        u32 c = min(fill[b], cap[b]);
        const u64 s0 = start[b];
        if (s0 + c > bound) c = s0 < bound ? (u32)(bound - s0) : 0u;   // (overflow case: results are discarded, stay in bounds)
        cnt[b] = c;
        v += c;
    }
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        if (t) atomicAdd(total, t);
    }
}
__global__ void check_total_kernel(const u64* total, u64 bound, u32* overflow) {
    if (*total > bound) *overflow = 1;
}

// records (staging order) -> per-bucket contiguous ranges.  Four independent slots per thread: the chain bucket id ->
// bucket offset / cursor atomic -> 16-byte copy is pure latency, so four chains are kept in flight per thread.
template <int RW>
__global__ void __launch_bounds__(256) scatter_records_kernel(const u64* __restrict__ rec, const u32* __restrict__ rec_bucket,
                                                              u64 n_slots, const u64* __restrict__ bucket_off,
                                                              u32* __restrict__ bucket_fill, u64* __restrict__ out) {
    constexpr int U = 4;
    const u64 base = (u64)blockIdx.x * (256 * U) + threadIdx.x;
    u32 b[U];
    u64 pos[U];
    ulonglong2 r0[U], r1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const u64 i = base + (u64)u * 256;
        b[u] = i < n_slots ? rec_bucket[i] : INVALID_BUCKET;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const u64 i = base + (u64)u * 256;
        pos[u] = 0;
        if (b[u] != INVALID_BUCKET) {
            pos[u] = bucket_off[b[u]] + atomicAdd(&bucket_fill[b[u]], 1u);
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(rec + i * RW);
            r0[u] = src[0];
            if (RW == 4) r1[u] = src[1];
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (b[u] == INVALID_BUCKET) continue;
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + pos[u] * RW);
        dst[0] = r0[u];
        if (RW == 4) dst[1] = r1[u];
    }
}

// ------------------------------------------------------------------------------------------------
// P2: per-bucket counting in shared memory
// ------------------------------------------------------------------------------------------------
// P2_VLIST: a k-mer whose count crosses min_obs is appended to a shared-memory list when it happens, so the emission writes
// the <= VL_CAP valid k-mers of a bucket without scanning the table (the scan stays for report_all and for buckets with more)
#ifndef P2_VLIST
#define P2_VLIST 1
#endif
static const int VL_CAP = 2048;
static const int MAX_PROBE = 96;
static const int SPLIT_STACK = 64;

template <int W> struct P2Cfg;
#ifndef P2_CAP1
#define P2_CAP1 8192
#define P2_THR1 512
#define P2_CTA1 2
#define P2_RC1 1024
#endif
// K <= 32: 8 B key + 4 B val = 96 KB tables, 2 CTAs/SM of 512 threads (4 x 256 threads with 48 KB tables measured 5% slower:
// more bucket splits).  RC = records per chunk.
template <> struct P2Cfg<1> { static const int CAP = P2_CAP1; static const int THREADS = P2_THR1; static const int CTAS = P2_CTA1; static const int RC = P2_RC1; };
#ifndef P2_CAP2
#define P2_CAP2 8192
#define P2_THR2 1024
#define P2_CTA2 1
#define P2_RC2 1024
#endif
template <> struct P2Cfg<2> { static const int CAP = P2_CAP2; static const int THREADS = P2_THR2; static const int CTAS = P2_CTA2; static const int RC = P2_RC2; };  // 16 B key + 4 B val = 160 KB, 1 CTA/SM

struct P2Args {
    u64* rec; u32* mult;  // records (deduplicated in place per bucket when mult != nullptr) and their multiplicities
    int mult_ready;       // records are already deduplicated (retry): bucket b holds dedup_cnt[b] records
    u32* dedup_cnt;       // per bucket: records left after deduplication
    const u64* bucket_start; const u32* bucket_cnt; u32 n_buckets;   // bucket b = records [start[b], start[b] + cnt[b])
    u32 min_obs; int stranded; int report_all;
    int task_len;         // k-mers per task of the expansion loop (8, or 16 when records hold more than 32 k-mers)
    u64* out_lo; u64* out_hi; u32* out_val; u64 cap_valid;
    u64* all_lo; u64* all_hi; u64 cap_all;
    u64* counters;  // [0] queue, [1] n_valid, [2] n_all(distinct), [3] splits, [4] error
};

__device__ __forceinline__ Kmer<2> cas128_shared(Kmer<2>* addr, Kmer<2> cmp, Kmer<2> val) {
    Kmer<2> old;
    u32 sa = (u32)__cvta_generic_to_shared(addr);
    asm volatile(
        "{\n .reg .b128 c, s, d;\n mov.b128 c, {%3, %4};\n mov.b128 s, {%5, %6};\n"
        " atom.shared.cas.b128 d, [%2], c, s;\n mov.b128 {%0, %1}, d;\n}\n"
        : "=l"(old.lo), "=l"(old.hi)
        : "r"(sa), "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi)
        : "memory");
    return old;
}

template <int W>
struct SmemTable {
    Kmer<W>* keys;
    u32* vals;  // count << 8 | exts
    __device__ __forceinline__ int find_or_insert(Kmer<W> key, u32 h, int cap, u32& fresh);   // fresh += 1 when the key is new
};
template <>
__device__ __forceinline__ int SmemTable<1>::find_or_insert(Kmer<1> key, u32 h, int cap, u32& fresh) {
    u32 slot = (h >> 18) & (cap - 1);  // class selection uses the low <= 18 bits
    u64* k64 = reinterpret_cast<u64*>(keys);
    for (int pr = 0; pr < MAX_PROBE; pr++) {
        u64 cur = *reinterpret_cast<volatile u64*>(k64 + slot);
        if (cur == key.lo) return (int)slot;
        if (cur == ~0ull) {
            u64 old = atomicCAS(k64 + slot, ~0ull, key.lo);
            if (old == ~0ull) { fresh++; return (int)slot; }
            if (old == key.lo) return (int)slot;
        }
        slot = (slot + 1) & (cap - 1);
    }
    return -1;
}
template <>
__device__ __forceinline__ int SmemTable<2>::find_or_insert(Kmer<2> key, u32 h, int cap, u32& fresh) {
    u32 slot = (h >> 18) & (cap - 1);
    const Kmer<2> empty{~0ull, ~0ull};
    for (int pr = 0; pr < MAX_PROBE; pr++) {
        volatile u64* kp = reinterpret_cast<volatile u64*>(keys + slot);
        if (kp[0] == key.lo && kp[1] == key.hi) return (int)slot;  // both halves equal => not torn
        Kmer<2> old = cas128_shared(keys + slot, empty, key);
        if (old.lo == ~0ull && old.hi == ~0ull) { fresh++; return (int)slot; }
        if (old.lo == key.lo && old.hi == key.hi) return (int)slot;
        slot = (slot + 1) & (cap - 1);
    }
    return -1;
}

template <int W>
__global__ void __launch_bounds__(P2Cfg<W>::THREADS, P2Cfg<W>::CTAS) count_kernel(KP kp, P2Args a) {
    constexpr int CAP = P2Cfg<W>::CAP;
    constexpr int P2T = P2Cfg<W>::THREADS;
    constexpr int P2_RC = P2Cfg<W>::RC;
    constexpr int RW = RecLayout<W>::WORDS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Kmer<W>* keys = reinterpret_cast<Kmer<W>*>(smem_raw);
    u32* vals = reinterpret_cast<u32*>(smem_raw + sizeof(Kmer<W>) * CAP);
    __shared__ u64 s_scan[33];
    constexpr int MAXT = W == 1 ? 4 : 8;   // tasks per record: 4 x 8 k-mers (<= 32 k-mers per 16-byte record), 8 x 8 (<= 59 per 32-byte record)
    __shared__ u16 s_task[P2_RC * MAXT];  // (record in chunk) | (task in record << 10), sorted by task length
    __shared__ u32 s_cls[17];             // per task length: counter / cursor; [0] = number of tasks
    __shared__ u32 s_ptot, s_next, s_wr_ok, s_cnt;
    __shared__ u64 s_r0;
    __shared__ u32 s_wsum[32];
    __shared__ u32 s_bucket, s_overflow, s_sp_cnt, s_sp_exts;
    __shared__ u64 s_pf_r0;   // the NEXT bucket's records, prefetched into L2 while this bucket is expanded
    __shared__ u32 s_pf_cnt;
    __shared__ u64 s_base_valid, s_base_all;
    __shared__ u32 s_stack[SPLIT_STACK];  // (residue << 6) | bits ; residue < 2^26 (deeper => error)
#if P2_VLIST
    __shared__ u32 s_vcnt, s_na;
    __shared__ u16 s_vlist[VL_CAP];
    const u32 mo = a.min_obs ? a.min_obs : 1u;
#endif
    SmemTable<W> tab{keys, vals};
    const int K = kp.k;
    const int tl = a.task_len;

    // thread 0 keeps one bucket in flight: the queue atomic and the bucket's (start, count) loads are issued while the
    // previous bucket is still being processed, so the CTA never idles on an L2 round trip between buckets
    u32 nb_id = 0, nb_cnt = 0;
    u64 nb_r0 = 0;
    if (threadIdx.x == 0) {
        nb_id = (u32)atomicAdd(&a.counters[0], 1ull);
        if (nb_id < a.n_buckets) { nb_r0 = a.bucket_start[nb_id]; nb_cnt = a.bucket_cnt[nb_id]; }
    }
    for (;;) {
        // (no barrier here: every path back to this point ends with one, after the last read of the values below)
        if (threadIdx.x == 0) {
            s_bucket = nb_id; s_r0 = nb_r0; s_cnt = nb_cnt;
            s_stack[0] = 0;
            if (nb_id < a.n_buckets) nb_id = (u32)atomicAdd(&a.counters[0], 1ull);   // result needed one bucket later
        }
        __syncthreads();
        const u32 b = s_bucket;
        if (b >= a.n_buckets) break;
        const u64 r0 = s_r0;
        u64 r1 = r0 + s_cnt;
        if (threadIdx.x == 0 && nb_id < a.n_buckets) { nb_r0 = a.bucket_start[nb_id]; nb_cnt = a.bucket_cnt[nb_id]; }
        if (r0 == r1) { __syncthreads(); continue; }
        const bool small_bucket = (r1 - r0) * 63ull < (1ull << 24);   // no count can overflow the 24-bit field
        if constexpr (W == 1) {
            if (a.mult && a.mult_ready) {
                r1 = r0 + a.dedup_cnt[b];
            } else if (a.mult) {
                // ---- P2a: deduplicate this bucket's records.  At sequencing coverage c most super-k-mers of a
                // genomic site occur ~c/2 times byte-identically; each distinct record is expanded once below and
                // its k-mers are counted with the record's multiplicity.  The 16-byte records are hashed into a
                // table that borrows the (not yet used) k-mer table memory; distinct records are written back
                // over the front of the bucket's own range (never ahead of what has been read). ----
                constexpr int RCAP = CAP / 2;   // 20 B per entry inside the 12 B x CAP table memory
                Kmer<2>* rkeys = reinterpret_cast<Kmer<2>*>(smem_raw);
                u32* rcnt = reinterpret_cast<u32*>(smem_raw + sizeof(Kmer<2>) * RCAP);
                u64 dbase = 0;
                for (u64 c0 = r0; c0 < r1; c0 += RCAP / 2) {
                    const u32 nrc = (u32)min((u64)(RCAP / 2), r1 - c0);
#if P2_VLIST
                    // the chunk's records of this thread are loaded up front, under the table clear: the probe loop's atomics are
                    // fences the next load could not be hoisted over (one L2 round trip per record otherwise)
                    constexpr int RPT = RCAP / 2 / P2T;
                    ulonglong2 vv[RPT];
#pragma unroll
                    for (int u = 0; u < RPT; u++) {
                        const u32 i = threadIdx.x + u * P2T;
                        vv[u] = i < nrc ? __ldcg(reinterpret_cast<const ulonglong2*>(a.rec + (c0 + i) * 2)) : make_ulonglong2(0, 0);
                    }
#endif
                    for (int i = threadIdx.x; i < RCAP; i += P2T) {
                        reinterpret_cast<u64*>(rkeys)[2 * i] = ~0ull;
                        reinterpret_cast<u64*>(rkeys)[2 * i + 1] = ~0ull;
                        rcnt[i] = 0;
                    }
#if P2_VLIST
                    if (threadIdx.x == 0) s_vcnt = 0;
#endif
                    __syncthreads();
#if P2_VLIST
                    // the thread whose CAS claims an empty slot appends the slot to a list (a chunk holds <= RCAP / 2 <= VL_CAP
                    // records): the compaction below walks the list, not the table
                    static_assert(RCAP / 2 <= VL_CAP, "one list entry per record of a chunk");
#pragma unroll
                    for (int u = 0; u < RPT; u++) {
                        const u32 i = threadIdx.x + u * P2T;
                        if (i >= nrc) break;
                        Kmer<2> key{vv[u].x, vv[u].y};
                        u32 slot = Ops<2>::hash32(key) >> 8 & (RCAP - 1);
                        const Kmer<2> empty{~0ull, ~0ull};
                        for (;;) {
                            volatile u64* kp2 = reinterpret_cast<volatile u64*>(rkeys + slot);
                            if (kp2[0] == key.lo && kp2[1] == key.hi) break;
                            Kmer<2> old = cas128_shared(rkeys + slot, empty, key);
                            if (old.lo == ~0ull && old.hi == ~0ull) { s_vlist[atomicAdd(&s_vcnt, 1u)] = (u16)slot; break; }
                            if (old.lo == key.lo && old.hi == key.hi) break;
                            slot = (slot + 1) & (RCAP - 1);
                        }
                        atomicAdd(&rcnt[slot], 1u);
                    }
                    __syncthreads();   // every record of the chunk has been read: the in-place writes below are safe
                    const u32 nd = s_vcnt;
                    for (u32 q = threadIdx.x; q < nd; q += P2T) {
                        const int sl = s_vlist[q];
                        const u64 o = r0 + dbase + q;
                        *reinterpret_cast<ulonglong2*>(a.rec + o * 2) = make_ulonglong2(rkeys[sl].lo, rkeys[sl].hi);
                        a.mult[o] = rcnt[sl];
                    }
                    dbase += nd;
                    __syncthreads();
                }
#else
                    for (u32 i = threadIdx.x; i < nrc; i += P2T) {
                        ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(a.rec + (c0 + i) * 2));
                        Kmer<2> key{v.x, v.y};
                        u32 slot = Ops<2>::hash32(key) >> 8 & (RCAP - 1);
                        const Kmer<2> empty{~0ull, ~0ull};
                        for (;;) {
                            volatile u64* kp2 = reinterpret_cast<volatile u64*>(rkeys + slot);
                            if (kp2[0] == key.lo && kp2[1] == key.hi) break;
                            Kmer<2> old = cas128_shared(rkeys + slot, empty, key);
                            if ((old.lo == ~0ull && old.hi == ~0ull) || (old.lo == key.lo && old.hi == key.hi)) break;
                            slot = (slot + 1) & (RCAP - 1);
                        }
                        atomicAdd(&rcnt[slot], 1u);
                    }
                    __syncthreads();
                    // compact: RCAP / P2T consecutive slots per thread
                    u32 mine = 0;
#pragma unroll
                    for (int j = 0; j < RCAP / P2T; j++) mine += rcnt[threadIdx.x * (RCAP / P2T) + j] != 0;
                    {
                        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                        u32 inc = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                        if (lane == 31) s_wsum[warp] = inc;
                        __syncthreads();
                        if (warp == 0) {
                            u32 w = lane < P2T / 32 ? s_wsum[lane] : 0, winc = w;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
                            s_wsum[lane] = winc - w;
                            if (lane == 31) s_ptot = winc;
                        }
                        __syncthreads();
                        u64 o = r0 + dbase + s_wsum[warp] + inc - mine;
#pragma unroll
                        for (int j = 0; j < RCAP / P2T; j++) {
                            int sl = threadIdx.x * (RCAP / P2T) + j;
                            u32 cnt = rcnt[sl];
                            if (cnt) {
                                *reinterpret_cast<ulonglong2*>(a.rec + o * 2) = make_ulonglong2(rkeys[sl].lo, rkeys[sl].hi);
                                a.mult[o] = cnt;
                                o++;
                            }
                        }
                        dbase += s_ptot;
                    }
                    __syncthreads();
                }
#endif
                if (threadIdx.x == 0) { atomicAdd(&a.counters[5], dbase); a.dedup_cnt[b] = (u32)dbase; }
                r1 = r0 + dbase;
                __threadfence_block();
                __syncthreads();
            }
        } else {
            if (a.mult) {
                // ---- P2a for 32-byte records: the table is keyed by a 128-bit hash of the record (one ATOMS.CAS.128);
                // the slot owner stores the full record, everybody else verifies it before adding to the multiplicity.
                // A mismatch (128-bit hash collision) only disables deduplication for that chunk. ----
                constexpr int RCAP = 2 * P2T;   // one record per thread and chunk, table at most half full
                static_assert(W == 1 || 52 * RCAP <= (int)(sizeof(Kmer<2>) + 4) * CAP, "dedup table fits the k-mer table memory");
                Kmer<2>* hkeys = reinterpret_cast<Kmer<2>*>(smem_raw);
                u64* srec = reinterpret_cast<u64*>(smem_raw + sizeof(Kmer<2>) * RCAP);
                u32* rcnt = reinterpret_cast<u32*>(smem_raw + (sizeof(Kmer<2>) + 32) * RCAP);
                u64 dbase = 0;
                for (u64 c0 = r0; c0 < r1; c0 += RCAP / 2) {
                    const u32 nrc = (u32)min((u64)(RCAP / 2), r1 - c0);
                    for (int i = threadIdx.x; i < RCAP; i += P2T) {
                        reinterpret_cast<u64*>(hkeys)[2 * i] = ~0ull;
                        reinterpret_cast<u64*>(hkeys)[2 * i + 1] = ~0ull;
                        rcnt[i] = 0;
                    }
                    if (threadIdx.x == 0) s_sp_cnt = 0;   // borrowed as the "verification failed" flag of this chunk
                    __syncthreads();
                    u64 w[4] = {0, 0, 0, 0};
                    int my_slot = -1;
                    const bool have = threadIdx.x < nrc;   // RCAP / 2 == P2T: one record per thread
                    if (have) {
                        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(a.rec + (c0 + threadIdx.x) * 4);
                        ulonglong2 v0 = __ldcg(src), v1 = __ldcg(src + 1);
                        w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; w[3] = v1.y;
                        Kmer<2> hk;
                        {
                            u64 x = w[0] ^ (w[1] * 0x9E3779B97F4A7C15ull) ^ (w[2] * 0xC2B2AE3D27D4EB4Full) ^ (w[3] * 0x165667B19E3779F9ull);
                            x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
                            u64 y = w[3] ^ (w[2] * 0x9E3779B97F4A7C15ull) ^ (w[1] * 0xC2B2AE3D27D4EB4Full) ^ (w[0] * 0x27D4EB2F165667C5ull);
                            y ^= y >> 29; y *= 0xBF58476D1CE4E5B9ull; y ^= y >> 32;
                            hk.lo = x; hk.hi = y & ~1ull;   // never the all-ones EMPTY marker
                        }
                        u32 slot = (u32)(hk.lo >> 20) & (RCAP - 1);
                        const Kmer<2> empty{~0ull, ~0ull};
                        for (;;) {
                            volatile u64* kp2 = reinterpret_cast<volatile u64*>(hkeys + slot);
                            if (kp2[0] == hk.lo && kp2[1] == hk.hi) break;
                            Kmer<2> old = cas128_shared(hkeys + slot, empty, hk);
                            if (old.lo == ~0ull && old.hi == ~0ull) {   // slot owner: publish the record itself
                                srec[4 * slot] = w[0]; srec[4 * slot + 1] = w[1]; srec[4 * slot + 2] = w[2]; srec[4 * slot + 3] = w[3];
                                break;
                            }
                            if (old.lo == hk.lo && old.hi == hk.hi) break;
                            slot = (slot + 1) & (RCAP - 1);
                        }
                        my_slot = (int)slot;
                    }
                    __syncthreads();
                    if (have) {
                        bool same = srec[4 * my_slot] == w[0] && srec[4 * my_slot + 1] == w[1] && srec[4 * my_slot + 2] == w[2] &&
                                    srec[4 * my_slot + 3] == w[3];
                        if (same) atomicAdd(&rcnt[my_slot], 1u);
                        else s_sp_cnt = 1;
                    }
                    __syncthreads();
                    const bool failed = s_sp_cnt != 0;
                    u32 mine = 0;
                    if (failed) mine = have ? 1u : 0u;   // keep every record of the chunk as is (multiplicity 1)
                    else {
#pragma unroll
                        for (int j = 0; j < RCAP / P2T; j++) mine += rcnt[threadIdx.x * (RCAP / P2T) + j] != 0;
                    }
                    {
                        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                        u32 inc = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                        if (lane == 31) s_wsum[warp] = inc;
                        __syncthreads();
                        if (warp == 0) {
                            u32 ww = lane < P2T / 32 ? s_wsum[lane] : 0, winc = ww;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
                            s_wsum[lane] = winc - ww;
                            if (lane == 31) s_ptot = winc;
                        }
                        __syncthreads();   // every record of the chunk is in registers / smem by now: in-place writes are safe
                        u64 o = r0 + dbase + s_wsum[warp] + inc - mine;
                        if (failed) {
                            if (have) {
                                ulonglong2* dst = reinterpret_cast<ulonglong2*>(a.rec + o * 4);
                                dst[0] = make_ulonglong2(w[0], w[1]);
                                dst[1] = make_ulonglong2(w[2], w[3]);
                                a.mult[o] = 1;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < RCAP / P2T; j++) {
                                int sl = threadIdx.x * (RCAP / P2T) + j;
                                u32 cnt = rcnt[sl];
                                if (cnt) {
                                    ulonglong2* dst = reinterpret_cast<ulonglong2*>(a.rec + o * 4);
                                    dst[0] = make_ulonglong2(srec[4 * sl], srec[4 * sl + 1]);
                                    dst[1] = make_ulonglong2(srec[4 * sl + 2], srec[4 * sl + 3]);
                                    a.mult[o] = cnt;
                                    o++;
                                }
                            }
                        }
                        dbase += s_ptot;
                    }
                    __syncthreads();
                }
                if (threadIdx.x == 0) { atomicAdd(&a.counters[5], dbase); a.dedup_cnt[b] = (u32)dbase; s_sp_cnt = 0; }
                r1 = r0 + dbase;
                __threadfence_block();
                __syncthreads();
            }
        }
        // the split stack's depth is mirrored in a register by every thread (pushes and pops are uniform decisions)
        int nstack = 1;
        while (nstack > 0) {
            const u32 top = s_stack[--nstack];
            const u32 cbits = top & 63u, cres = top >> 6;
            const u32 cmask = cbits ? ((1u << cbits) - 1) : 0;
            for (int i = threadIdx.x; i < CAP; i += P2T) {
                if (W == 1) reinterpret_cast<u64*>(keys)[i] = ~0ull;
                else { reinterpret_cast<u64*>(keys)[2 * i] = ~0ull; reinterpret_cast<u64*>(keys)[2 * i + 1] = ~0ull; }
                vals[i] = 0;
            }
            if (threadIdx.x == 0) { s_overflow = 0; s_sp_cnt = 0; s_sp_exts = 0; }
#if P2_VLIST
            if (threadIdx.x == 32) { s_vcnt = 0; s_na = 0; }
            u32 fresh = 0;
#else
            u32 fresh = 0;
#endif
            __syncthreads();
            // ---- expand records, insert.  The bucket is processed in chunks of RC records.  Every record is cut into
            // TASKS of <= tl consecutive k-mers and the chunk's tasks are counting-sorted by length (descending) in
            // shared memory, so the 32 lanes of a warp roll tasks of (almost always) the same length: one uniform
            // loop with no record switching inside it, whatever the record lengths are. ----
            for (u64 c0 = r0; c0 < r1; c0 += P2_RC) {
                const u32 nrc = (u32)min((u64)P2_RC, r1 - c0);
                const int lane = threadIdx.x & 31;
                if (threadIdx.x <= 16) s_cls[threadIdx.x] = 0;
                __syncthreads();
                u32 myn[P2_RC / P2T];
#pragma unroll
                for (int j = 0; j < P2_RC / P2T; j++) {   // class sizes
                    const u32 idx = j * P2T + threadIdx.x;
                    const u32 n = idx < nrc ? ((u32)__ldcg(a.rec + (c0 + idx) * RW + (RW - 1)) >> 8) & 63u : 0;
                    myn[j] = n;
                    const u32 full = n / tl, rem = n - full * tl;
                    const u32 fsum = __reduce_add_sync(0xffffffffu, full);
                    if (lane == 0 && fsum) atomicAdd(&s_cls[tl], fsum);
                    if (rem) atomicAdd(&s_cls[rem], 1u);
                }
                __syncthreads();
                if (threadIdx.x == 0) {   // class starts, longest tasks first
                    u32 acc = 0;
                    for (int l = tl; l >= 1; l--) { u32 c = s_cls[l]; s_cls[l] = acc; acc += c; }
                    s_cls[0] = acc;
                    s_next = 0;
                    if (c0 == r0) { s_pf_r0 = nb_r0; s_pf_cnt = nb_id < a.n_buckets ? nb_cnt : 0u; }
                }
                __syncthreads();
                if (c0 == r0 && cbits == 0) {
                    // the dedup phase of the next bucket starts with one dependent load per record: pull its records into L2 now
                    const char* pf = reinterpret_cast<const char*>(a.rec + s_pf_r0 * RW);
                    const u32 lines = (s_pf_cnt * (u32)(RW * 8) + 127u) >> 7;
                    for (u32 i = threadIdx.x; i < lines; i += P2T) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (u64)i * 128));
                }
#pragma unroll
                for (int j = 0; j < P2_RC / P2T; j++) {   // scatter (record, first k-mer) into the class ranges
                    const u32 idx = j * P2T + threadIdx.x;
                    const u32 n = myn[j];
                    for (u32 jt = 0; jt < (u32)MAXT; jt++) {
                        const u32 t0 = jt * tl;
                        const bool has = t0 < n;
                        if (!__any_sync(0xffffffffu, has)) break;
                        const u32 len = has ? min((u32)tl, n - t0) : 0;
                        const bool isfull = len == (u32)tl;
                        const u32 fm = __ballot_sync(0xffffffffu, isfull);
                        u32 base = 0;
                        if (fm && lane == __ffs(fm) - 1) base = atomicAdd(&s_cls[tl], (u32)__popc(fm));
                        if (fm) base = __shfl_sync(0xffffffffu, base, __ffs(fm) - 1);
                        u32 pos = base + __popc(fm & ((1u << lane) - 1));
                        if (has && !isfull) pos = atomicAdd(&s_cls[len], 1u);
                        if (has) s_task[pos] = (u16)(idx | (jt << 10));
                    }
                }
                __syncthreads();
                const u32 NT = s_cls[0];
                // every lane takes its next task from a shared cursor, longest tasks first (LPT): lanes of a warp hold
                // tasks of (almost) equal length, and no warp waits at the barrier below for longer than one short task.
                // Per-lane claims on purpose: no warp-level primitive inside this loop, so lanes may drift freely.
                for (;;) {
                    const u32 q = atomicAdd(&s_next, 1u);
                    if (q >= NT) break;
                    {
                    const u32 task = s_task[q];
                    const u32 ri = task & 1023u;
                    int t = (int)((task >> 10) * tl);
                    u64 s[RW];
                    {
                        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(a.rec + (c0 + ri) * RW);
                        ulonglong2 v0 = __ldcg(src);   // L2 path: records may have been rewritten by this CTA (dedup)
                        s[0] = v0.x; s[1] = v0.y;
                        if constexpr (RW == 4) { ulonglong2 v1 = __ldcg(src + 1); s[2] = v1.x; s[RW - 1] = v1.y; }
                    }
                    const u32 mcur = a.mult ? __ldcg(a.mult + c0 + ri) : 1u;
                    const u32 hdr = (u32)s[RW - 1] & 0x3fffu;
                    const int n = (int)(hdr >> 8);
                    const int tend = min(n, t + tl);
                    const u32 rn = (hdr >> 4) & 0xfu, ln = hdr & 0xfu;
                    u32 prev_first = 0;
                    if (t > 0) {  // task starts inside the record: drop t bases from the front (static register indices only)
                        const int pb = t - 1, pw = pb >> 5;
                        u64 wsel = s[0];
                        if (pw == 1) wsel = s[1];
                        if constexpr (RW == 4) { if (pw == 2) wsel = s[2]; if (pw == 3) wsel = s[3]; }
                        prev_first = (u32)(wsel >> (62 - 2 * (pb & 31))) & 3u;
                        const int ws = t >> 5, bs = 2 * (t & 31);
                        if constexpr (RW == 2) {
                            if (ws) { s[0] = s[1]; s[1] = 0; }
                        } else {
                            if (ws & 1) { s[0] = s[1]; s[1] = s[2]; s[2] = s[3]; s[3] = 0; }
                            if (ws & 2) { s[0] = s[2]; s[1] = s[3]; s[2] = 0; s[3] = 0; }
                        }
                        if (bs) {
#pragma unroll
                            for (int q2 = 0; q2 < RW - 1; q2++) s[q2] = (s[q2] << bs) | (s[q2 + 1] >> (64 - bs));
                            s[RW - 1] <<= bs;
                        }
                    }
                    Kmer<W> fwd;
                    if constexpr (W == 1) {
                        fwd.lo = s[0] >> (64 - 2 * K);
                    } else {
                        const int sh = 128 - 2 * K;  // first K (33..64) bases = top 2K bits of (s0:s1)
                        fwd.hi = sh ? (s[0] >> sh) : s[0];
                        fwd.lo = sh ? ((s[1] >> sh) | (s[0] << (64 - sh))) : s[1];
                    }
                    Kmer<W> rcv = Ops<W>::rc(kp, fwd);
                    // the (at most 16) bases a task can touch, as two 32-bit windows: PF = bases t .. t+15 (the first bases of its
                    // k-mers), NX = bases t+K .. t+K+15 (the bases shifted in); no shift register of the whole record in the loop
                    u32 PF = (u32)(s[0] >> 32), NX;
                    if constexpr (W == 1) {
                        NX = K < 32 ? (u32)(((s[0] << (2 * K)) | (s[1] >> (64 - 2 * K))) >> 32) : (u32)(s[1] >> 32);
                    } else {
                        const int kk = K - 32;   // 1..32
                        NX = kk < 32 ? (u32)(((s[1] << (2 * kk)) | (s[2] >> (64 - 2 * kk))) >> 32) : (u32)(s[2] >> 32);
                    }
                    // left nibble of the forward Exts: the record's own left mask for its first k-mer, else the base in front
                    u32 pn = t == 0 ? ln : (1u << prev_first);
                    for (; t < tend; t++) {
                        const u32 nb = NX >> 30;   // base t+K (only meaningful when t < n-1)
                        NX <<= 2;
                        const u32 cf = PF >> 30;   // first base of this k-mer
                        PF <<= 2;
                        const bool use_rc = !a.stranded && !(fwd < rcv);   // lib.rs:224-231 (equality -> flipped), filter.rs:190-196
                        // forward Exts, then Exts::rc (lib.rs:746) when the rc is the canonical form: complement + swap of the
                        // nibbles sends bit i to bit 7 - i, i.e. an 8-bit reversal
                        const u32 ef = pn | ((t == n - 1 ? rn : (1u << nb)) << 4);
                        const u32 e = use_rc ? __brev(ef) >> 24 : ef;
                        const Kmer<W> key = use_rc ? rcv : fwd;
                        const u32 h = Ops<W>::hash32(key);
                        if ((h & cmask) == cres) {
                            bool special;
                            if constexpr (W == 1) special = key.lo == ~0ull;
                            else special = key.lo == ~0ull && key.hi == ~0ull;
                            if (special) {  // all-T k-mer at K = 32/64 stranded collides with the EMPTY sentinel
                                atomicAdd(&s_sp_cnt, mcur);
                                atomicOr(&s_sp_exts, e);
                            } else {
                                const int slot = tab.find_or_insert(key, h, CAP, fresh);
                                if (slot < 0) {   // table full: stop everybody's claims (no flag polled in the loop), the bucket is split
                                    atomicExch(&s_overflow, 1u);
                                    atomicMax(&s_next, 0x40000000u);
                                    break;
                                }
                                u32 v = *reinterpret_cast<volatile u32*>(vals + slot);
                                if (e & ~v) atomicOr(vals + slot, e);
                                if (small_bucket) {
                                    // the bucket's total occurrences fit the 24-bit field: plain add, clamped at emission
#if P2_VLIST
                                    const u32 old = atomicAdd(vals + slot, mcur << 8) >> 8;
                                    if (old < mo && old + mcur >= mo) {
                                        const u32 qv = atomicAdd(&s_vcnt, 1u);
                                        if (qv < VL_CAP) s_vlist[qv] = (u16)slot;
                                    }
#else
                                    atomicAdd(vals + slot, mcur << 8);
#endif
                                } else {
                                    // saturating count (filter.rs:57): stop adding once 65535 is reached; steps of <= 255
                                    // keep the transient overshoot far below the 24-bit field
                                    for (u32 rem = mcur;;) {
                                        u32 stp = min(rem, 255u);
#if P2_VLIST
                                        if ((v >> 8) < 65535u) {
                                            const u32 old = atomicAdd(vals + slot, stp << 8) >> 8;
                                            if (old < mo && old + stp >= mo) {
                                                const u32 qv = atomicAdd(&s_vcnt, 1u);
                                                if (qv < VL_CAP) s_vlist[qv] = (u16)slot;
                                            }
                                        }
#else
                                        if ((v >> 8) < 65535u) atomicAdd(vals + slot, stp << 8);
#endif
                                        rem -= stp;
                                        if (!rem) break;
                                        v = *reinterpret_cast<volatile u32*>(vals + slot);
                                    }
                                }
                            }
                        }
                        // roll to the next k-mer of this record
                        pn = 1u << cf;
                        fwd = Ops<W>::ext_right(kp, fwd, nb);
                        rcv = Ops<W>::roll_rc(kp, rcv, nb);
                    }
                    }
                }
#if P2_VLIST
                if (fresh) { atomicAdd(&s_na, fresh); fresh = 0; }
#endif
                __syncthreads();
                if (s_overflow) break;   // (uniform: every thread reads the flag after the barrier)
            }
            if (s_overflow) {
                const bool deep = cbits >= 18 || nstack + 2 > SPLIT_STACK;
                if (threadIdx.x == 0) {
                    if (deep) {
                        a.counters[4] = 1;  // cannot happen for distinct keys below 2^26 classes; reported as internal error
                    } else {
                        s_stack[nstack] = ((cres | (1u << cbits)) << 6) | (cbits + 1);
                        s_stack[nstack + 1] = (cres << 6) | (cbits + 1);
                        atomicAdd(&a.counters[3], 1ull);
                    }
                }
                if (!deep) nstack += 2;
                __syncthreads();
                continue;
            }
#if P2_VLIST
            if (!a.report_all && s_vcnt <= (u32)VL_CAP) {
                // ---- emit from the list: one reservation, one k-mer per thread and round ----
                const u32 nvl = s_vcnt;
                if (threadIdx.x == 0) {
                    const u32 sc = s_sp_cnt > 65535u ? 65535u : s_sp_cnt;
                    const u32 spv = (s_sp_cnt && sc >= a.min_obs) ? 1u : 0u;
                    const u32 tv = nvl + spv, ta = s_na + (s_sp_cnt ? 1u : 0u);
                    s_base_valid = tv ? atomicAdd(&a.counters[1], (u64)tv) : 0;
                    atomicAdd(&a.counters[2], (u64)ta);
                    const bool over = s_base_valid + tv > a.cap_valid;
                    if (over) a.counters[4] = 2;
                    s_wr_ok = over ? 0u : 1u;
                    if (!over && spv) {
                        a.out_lo[s_base_valid + nvl] = ~0ull;
                        if (W == 2) a.out_hi[s_base_valid + nvl] = ~0ull;
                        a.out_val[s_base_valid + nvl] = (s_sp_exts & 0xffu) | (sc << 8);
                    }
                }
                __syncthreads();
                if (s_wr_ok) {
                    const u64 pv0 = s_base_valid;
                    for (u32 q = threadIdx.x; q < nvl; q += P2T) {
                        const int i = s_vlist[q];
                        const u32 v = vals[i];
                        u32 cc = v >> 8;
                        cc = cc > 65535u ? 65535u : cc;
                        a.out_lo[pv0 + q] = W == 1 ? reinterpret_cast<u64*>(keys)[i] : reinterpret_cast<u64*>(keys)[2 * i];
                        if (W == 2) a.out_hi[pv0 + q] = reinterpret_cast<u64*>(keys)[2 * i + 1];
                        a.out_val[pv0 + q] = (v & 0xffu) | (cc << 8);
                    }
                }
                __syncthreads();
                continue;
            }
#endif
            // ---- emit: count, block scan, reserve, write ----
            u32 nv = 0, na = 0;
            u32 occm = 0, valm = 0;   // bit j: slot threadIdx.x + j * P2T is occupied / valid (CAP / P2T <= 32 slots per thread)
            static_assert(CAP / P2T <= 32, "slot masks are 32 bits");
#pragma unroll
            for (int j = 0; j < CAP / P2T; j++) {
                const int i = threadIdx.x + j * P2T;
                bool occ = W == 1 ? reinterpret_cast<u64*>(keys)[i] != ~0ull
                                  : !(reinterpret_cast<u64*>(keys)[2 * i] == ~0ull && reinterpret_cast<u64*>(keys)[2 * i + 1] == ~0ull);
                if (occ) {
                    na++;
                    occm |= 1u << j;
                    u32 c = vals[i] >> 8;
                    c = c > 65535u ? 65535u : c;
                    if (c >= a.min_obs) { nv++; valm |= 1u << j; }
                }
            }
            if (threadIdx.x == 0 && s_sp_cnt) {
                na++;
                u32 c = s_sp_cnt > 65535u ? 65535u : s_sp_cnt;
                if (c >= a.min_obs) nv++;
            }
            u64 tot;
            u64 packed = ((u64)na << 32) | nv;
            u64 ex;
            {   // block exclusive scan of (na, nv) packed in one u64 (CAP < 2^31 so no carry between halves)
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                u64 inc = packed;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    u64 t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) s_scan[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    u64 w = lane < P2T / 32 ? s_scan[lane] : 0;
                    u64 winc = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        u64 t = __shfl_up_sync(0xffffffffu, winc, o);
                        if (lane >= o) winc += t;
                    }
                    s_scan[lane] = winc - w;
                    if (lane == 31) s_scan[32] = winc;
                }
                __syncthreads();
                tot = s_scan[32];
                ex = s_scan[warp] + inc - packed;
            }
            if (threadIdx.x == 0) {
                u32 tv = (u32)tot, ta = (u32)(tot >> 32);
                s_base_valid = tv ? atomicAdd(&a.counters[1], (u64)tv) : 0;
                s_base_all = atomicAdd(&a.counters[2], (u64)ta);
                const bool over = s_base_valid + tv > a.cap_valid || (a.report_all && s_base_all + ta > a.cap_all);
                if (over) a.counters[4] = 2;
                s_wr_ok = over ? 0u : 1u;
            }
            __syncthreads();
            if (s_wr_ok) {
                u64 pv = s_base_valid + (u32)ex, pa = s_base_all + (u32)(ex >> 32);
                // only the slots the first pass marked: valid ones (a few per cent of the table on reads with errors), or
                // every occupied one when all_kmers is requested — in slot order, like the counts above
                for (u32 m = a.report_all ? occm : valm; m; m &= m - 1) {
                    const int j = __ffs(m) - 1;
                    const int i = threadIdx.x + j * P2T;
                    u64 klo = W == 1 ? reinterpret_cast<u64*>(keys)[i] : reinterpret_cast<u64*>(keys)[2 * i];
                    u64 khi = W == 1 ? 0 : reinterpret_cast<u64*>(keys)[2 * i + 1];
                    if (a.report_all) { a.all_lo[pa] = klo; if (W == 2) a.all_hi[pa] = khi; pa++; }
                    if ((valm >> j) & 1u) {
                        u32 v = vals[i];
                        u32 c = v >> 8;
                        c = c > 65535u ? 65535u : c;
                        a.out_lo[pv] = klo;
                        if (W == 2) a.out_hi[pv] = khi;
                        a.out_val[pv] = (v & 0xffu) | (c << 8);
                        pv++;
                    }
                }
                if (!a.report_all) pa += na;
                if (threadIdx.x == 0 && s_sp_cnt) {
                    u32 c = s_sp_cnt > 65535u ? 65535u : s_sp_cnt;
                    if (a.report_all) { a.all_lo[pa] = ~0ull; if (W == 2) a.all_hi[pa] = ~0ull; }
                    if (c >= a.min_obs) {
                        a.out_lo[pv] = ~0ull;
                        if (W == 2) a.out_hi[pv] = ~0ull;
                        a.out_val[pv] = (s_sp_exts & 0xffu) | (c << 8);
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void unpack_vals_kernel(const u32* __restrict__ val, u8* __restrict__ exts, u16* __restrict__ counts, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 v = val[i];
    exts[i] = (u8)(v & 0xffu);
    counts[i] = (u16)(v >> 8);
}

__global__ void __launch_bounds__(256) count_input_kmers_kernel(const u32* __restrict__ length, u64 n, int k, u64* out, u32* max_len) {
    __shared__ u64 s_v[8];
    __shared__ u32 s_l[8];
    u64 v = 0;
    u32 L = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u32 l = length[i];
        L = max(L, l);
        v += l >= (u32)k ? l - k + 1 : 0;
    }
    for (int o = 16; o; o >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, o);
        L = max(L, __shfl_down_sync(0xffffffffu, L, o));
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = v; s_l[threadIdx.x >> 5] = L; }
    __syncthreads();
    if (threadIdx.x == 0) {   // one atomic per CTA (same-address atomics serialise)
        u64 tv = 0;
        u32 tl = 0;
        for (int w = 0; w < 8; w++) { tv += s_v[w]; tl = max(tl, s_l[w]); }
        if (tv) atomicAdd(out, tv);
        atomicMax(max_len, tl);
    }
}

__global__ void item_count_kernel(const u32* __restrict__ length, u32 uniform_len, u64 n, int k, u32* cnt) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 L = uniform_len ? uniform_len : length[i];
    cnt[i] = L >= (u32)k ? (L - k + 1 + CHUNK - 1) / CHUNK : 0;
}
__global__ void item_fill_kernel(const u32* __restrict__ cnt, const u64* __restrict__ off, u64 n, u32* item_seq, u32* item_j0) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 c = cnt[i];
    u64 o = off[i];
    for (u32 t = 0; t < c; t++) { item_seq[o + t] = (u32)i; item_j0[o + t] = t * CHUNK; }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
// MSP plan shared by every rank of a sharded run: minimizer length and number of buckets from the
// TOTAL number of input k-mers.
void plan_filter(const Ctx* c, int k, u64 N, int* p_out, int* bbits_out) {
    const int W = k <= 32 ? 1 : 2;
    u64 target = c->target_bucket_occ > 0 ? (u64)c->target_bucket_occ : 0;
    if (!target) {
        // k-mer occurrences per bucket such that the bucket's DISTINCT k-mers fit the shared-memory table
        // (8192 slots for one-word keys, 4096 for two-word keys; longer k-mers are also more often distinct)
        const u64 tmax = W == 1 ? 2 * (u64)P2Cfg<1>::CAP : 4096;
        target = N / ((u64)c->sm_count * 8);
        if (target < tmax / 8) target = tmax / 8;
        if (target > tmax) target = tmax;
    }
    int bbits = 0;
    while (bbits < 20 && (N >> (bbits + 1)) >= target) bbits++;
    // Minimizer length: 12 up to 2^16 buckets, one more base per further bucket bit.  A bucket is the union of the
    // minimizers whose hash ends in its bits; with p fixed, finer bucketing leaves only a handful of (very unevenly
    // used) minimizers per bucket and the bucket sizes spread out: 7 141 table-overflow splits at 2^19 buckets with
    // p = 12 against 94 at 2^16.  4^p / 2 canonical p-mers >= 128 per bucket keeps the spread of the 2^16 case.
    int p = c->msp_p > 0 ? c->msp_p : 12 + (bbits > 16 ? bbits - 16 : 0);
    if (p > k - 3) p = k - 3;  // window of K-p+1 >= 4 p-mers (register-tiled window minimum)
    if (p > 16) p = 16;
    if (p < 1) p = 1;
    if (k - p > 63) p = k - 63;
    *p_out = p;
    *bbits_out = bbits;
}

static int count_input(Ctx* c, int k, const SeqSet* s, u64* N_out, u32* max_len_out) {
    u64 N = 0;
    u32 max_len = s->max_len;
    if (s->uniform_len) {
        N = s->uniform_len >= (u32)k ? (u64)(s->uniform_len - k + 1) * s->n_seqs : 0;
        max_len = s->uniform_len;
    } else if (s->n_seqs) {
        DBuf<u64> tmp;
        TRY(tmp.alloc(c, 2));
        TRY(tmp.zero());
        count_input_kmers_kernel<<<(u32)std::min<u64>(grid_for(s->n_seqs, 256), (u64)c->sm_count * 8), 256, 0, c->stream>>>(s->length, s->n_seqs, k, tmp.p, (u32*)(tmp.p + 1));
        TRY(check_launch(c, "count_input_kmers"));
        u64 h[2];
        TRY(read_u64(c, tmp.p, h, 2));
        N = h[0];
        max_len = (u32)h[1];
    }
    *N_out = N;
    *max_len_out = max_len;
    return DBG_OK;
}

int count_input_kmers_dev(Ctx* c, int k, const SeqSet* s, u64* N_out) {
    u32 max_len = 0;
    return count_input(c, k, s, N_out, &max_len);
}

// Output of the partition stage: super-k-mer records grouped by MSP bucket.
struct PartOut {
    DBuf<u64> rec;         // n_rec records of RW words, bucket-contiguous
    DBuf<u64> bucket_off;  // NB + 1 exclusive offsets (device)
    DBuf<u32> bucket_count;
    u64 n_rec = 0;
};

// ---- P1 + P1b: sequences -> bucket-contiguous super-k-mer records ----
template <int W>
static int partition_stage(Ctx* c, int k, const SeqSet* s, int stranded, u64 N, u32 max_len, int p, int bbits, u32 bk_lo,
                           u32 bk_span, bool pool_out, PartOut& po) {
    constexpr int RW = RecLayout<W>::WORDS;
    KP kp = make_kp(k);
    cudaStream_t st = c->stream;
    const u32 NB = 1u << bbits;
    const bool use_tiles = s->contiguous && s->total_end > s->base0;
    const int maxk = rec_max_kmers(RW, k);
    // work items (general kernel only)
    DBuf<u32> item_seq, item_j0;
    u64 n_items = s->n_seqs;
    bool chunked = !use_tiles && max_len >= (u32)k && (max_len - k + 1) > (u32)CHUNK;
    if (chunked) {
        DBuf<u32> cnt;
        DBuf<u64> off, tot;
        TRY(cnt.alloc(c, s->n_seqs));
        TRY(off.alloc(c, s->n_seqs));
        TRY(tot.alloc(c, 1));
        item_count_kernel<<<grid_for(s->n_seqs, 256), 256, 0, st>>>(s->length, s->uniform_len, s->n_seqs, k, cnt.p);
        TRY(check_launch(c, "item_count"));
        TRY(exclusive_scan_u32_to_u64(c, cnt.p, off.p, s->n_seqs, tot.p));
        TRY(read_u64(c, tot.p, &n_items));
        TRY(item_seq.alloc(c, n_items));
        TRY(item_j0.alloc(c, n_items));
        item_fill_kernel<<<grid_for(s->n_seqs, 256), 256, 0, st>>>(cnt.p, off.p, s->n_seqs, item_seq.p, item_j0.p);
        TRY(check_launch(c, "item_fill"));
    }
    DBuf<u32> bucket_fill;
    DBuf<u64> ctr;
    if (pool_out) { TRY(po.bucket_count.alloc_pool(c, NB)); TRY(po.bucket_off.alloc_pool(c, (u64)NB + 1)); }
    else { TRY(po.bucket_count.alloc(c, NB)); TRY(po.bucket_off.alloc(c, (u64)NB + 1)); }
    TRY(bucket_fill.alloc(c, NB));
    TRY(ctr.alloc(c, 8));
    TileArgs ta;
    ta.base0 = s->base0; ta.total_end = s->total_end; ta.tile0 = 0; ta.tstride = 1;
    const u64 bp = (u64)WT_CW * tile_owned(k, p);   // positions per block of the tile kernel
    ta.n_tiles = use_tiles ? (s->total_end - s->base0 + bp - 1) / bp : 0;
    SeqSet* sm = const_cast<SeqSet*>(s);
    if (!use_tiles) TRY(seqset_ready(c, sm));
    u32 grid1 = use_tiles ? (u32)std::min<u64>(ta.n_tiles, (u64)c->sm_count * WT_CTAS)
                          : (u32)std::min<u64>((n_items + P1_WARPS - 1) / P1_WARPS, (u64)c->sm_count * 6);
    u64 n_warps = (u64)grid1 * (use_tiles ? WT_CW : P1_WARPS);
    // records expected in this pass: ~1/10 of its k-mer occurrences; N/4 leaves room, the retry below covers the rest
    u64 capacity = (u64)((double)N * bk_span / NB) / 4 + n_warps * WCHUNK + 1024;
    DBuf<u64> stage_rec;
    DBuf<u32> stage_bucket;
    u64 n_slots = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        TRY(stage_rec.alloc(c, capacity * RW));
        TRY(stage_bucket.alloc(c, capacity));
        TRY(stage_bucket.fill_ff());
        TRY(po.bucket_count.zero());
        TRY(ctr.zero());
        P1Args a;
        a.words = s->words; a.n_words = s->n_words; a.start = s->start; a.length = s->length; a.seq_exts = s->seq_exts;
        a.n_seqs = s->n_seqs; a.uniform_len = s->uniform_len;
        a.item_seq = chunked ? item_seq.p : nullptr; a.item_j0 = chunked ? item_j0.p : nullptr; a.n_items = n_items;
        a.p = p; a.stranded = stranded; a.bucket_mask = NB - 1; a.maxk = maxk; a.bk_lo = bk_lo; a.bk_span = bk_span;
        a.rec = stage_rec.p; a.rec_bucket = stage_bucket.p; a.capacity = capacity;
        a.cursor = ctr.p; a.bucket_count = po.bucket_count.p; a.overflow = (u32*)(ctr.p + 1);
        a.mode = 0; a.bucket_start = nullptr; a.bucket_cap = nullptr; a.bucket_fill = nullptr;
        CU(c, cudaEventRecord(c->ev[8], st));
        if (use_tiles && sm->n_pending > 0) {
            // pipelined upload: one launch per arrived chunk, covering the tiles whose bases (plus halo) are on the device
            u64 done = 0;
            const int np = sm->n_pending;
            for (int ci = 0; ci < np; ci++) {
                CU(c, cudaStreamWaitEvent(st, sm->pend_ev[ci], 0));
                u64 bases_ok = sm->pend_words_end[ci] * 32;
                // a block touches bases up to its end + K + 2 (flank); 192 covers that for every K <= 64
                u64 upto = ci == np - 1 ? ta.n_tiles : (bases_ok > s->base0 + bp + 192 ? (bases_ok - 192 - s->base0) / bp : 0);
                if (upto > ta.n_tiles) upto = ta.n_tiles;
                if (upto > done) {
                    TileArgs tc = ta;
                    tc.tile0 = done; tc.n_tiles = upto;
                    u32 g = (u32)std::min<u64>(upto - done, (u64)c->sm_count * WT_CTAS);
                    msp_tile_kernel<W><<<g, WT_THREADS, 0, st>>>(kp, a, tc);
                    TRY(check_launch(c, "msp_partition"));
                    done = upto;
                }
            }
            sm->n_pending = 0;
        } else {
            if (use_tiles) msp_tile_kernel<W><<<grid1, WT_THREADS, 0, st>>>(kp, a, ta);
            else msp_partition_kernel<W><<<grid1, P1_THREADS, 0, st>>>(kp, a);
            TRY(check_launch(c, "msp_partition"));
        }
        CU(c, cudaEventRecord(c->ev[9], st));
        u64 h[2];
        TRY(read_u64(c, ctr.p, h, 2));
        n_slots = h[0];
        if (!(u32)h[1]) break;
        if (attempt == 1) DBG_SET_ERR(c, DBG_E_INTERNAL, "record staging overflow after retry (%llu slots)", (unsigned long long)h[0]);
        capacity = h[0] + 1024;  // the cursor says exactly how much was needed
    }
    TRY(exclusive_scan_u32_to_u64(c, po.bucket_count.p, po.bucket_off.p, NB, po.bucket_off.p + NB));
    TRY(read_u64(c, po.bucket_off.p + NB, &po.n_rec));
    if (pool_out) TRY(po.rec.alloc_pool(c, po.n_rec * RW)); else TRY(po.rec.alloc(c, po.n_rec * RW));
    TRY(bucket_fill.zero());
    if (n_slots > capacity) n_slots = capacity;
    scatter_records_kernel<RW><<<grid_for(n_slots, 1024), 256, 0, st>>>(stage_rec.p, stage_bucket.p, n_slots, po.bucket_off.p,
                                                                        bucket_fill.p, po.rec.p);
    TRY(check_launch(c, "scatter_records"));
    return DBG_OK;
}

// ---- P1 direct: sequences -> per-bucket regions without staging or scatter (contiguous layouts, one pass) ----
// A sampling pass over ~1/16 of the tiles (bucket histogram only) sizes every bucket's region (estimate + 6 sigma of the
// thinning noise + slack); the main pass then takes each record's slot from its bucket's own cursor and writes it in
// place.  Nothing is read back here: the caller checks ctr[1] (a region, or the total, was too small -> staging path) and
// ctr[2] (records stored) together with the counting stage's counters.
struct DirectOut {
    DBuf<u64> rec, bucket_start, ctr;   // ctr: [0] sum of capacities, [1] overflow flag, [2] records stored
    DBuf<u32> cap, fill, cnt;
    u64 rec_bound = 0;
};

template <int W>
static int partition_direct(Ctx* c, int k, const SeqSet* s, int stranded, u64 N, int p, int bbits, DirectOut& d) {
    constexpr int RW = RecLayout<W>::WORDS;
    KP kp = make_kp(k);
    cudaStream_t st = c->stream;
    const u32 NB = 1u << bbits;
    SeqSet* sm = const_cast<SeqSet*>(s);
    TileArgs ta;
    ta.base0 = s->base0; ta.total_end = s->total_end; ta.tile0 = 0; ta.tstride = 1;
    const u64 bp = (u64)WT_CW * tile_owned(k, p);   // positions per block of the tile kernel
    ta.n_tiles = (s->total_end - s->base0 + bp - 1) / bp;
    // expected records: one per (K-p+2)/2 k-mers (window of K-p+1 p-mers); 3.5x covers read ends, length caps and the
    // regions' slack — a larger need raises the overflow flag and the caller falls back to staging
    d.rec_bound = (u64)(3.5 * 2.0 * (double)N / (double)(k - p + 2)) + (u64)NB * 64 + 4096;
    TRY(d.rec.alloc(c, d.rec_bound * RW));
    TRY(d.bucket_start.alloc(c, (u64)NB + 1));
    TRY(d.cap.alloc(c, NB)); TRY(d.fill.alloc(c, NB)); TRY(d.cnt.alloc(c, NB)); TRY(d.ctr.alloc(c, 4));
    TRY(d.cnt.zero()); TRY(d.fill.zero()); TRY(d.ctr.zero());
    P1Args a;
    a.words = s->words; a.n_words = s->n_words; a.start = s->start; a.length = s->length; a.seq_exts = s->seq_exts;
    a.n_seqs = s->n_seqs; a.uniform_len = s->uniform_len;
    a.item_seq = nullptr; a.item_j0 = nullptr; a.n_items = 0;
    a.p = p; a.stranded = stranded; a.bucket_mask = NB - 1; a.maxk = rec_max_kmers(RW, k); a.bk_lo = 0; a.bk_span = NB;
    a.rec = d.rec.p; a.rec_bucket = nullptr; a.capacity = d.rec_bound;
    a.cursor = nullptr; a.bucket_count = d.cnt.p; a.overflow = (u32*)(d.ctr.p + 1);
    a.bucket_start = d.bucket_start.p; a.bucket_cap = d.cap.p; a.bucket_fill = d.fill.p;
    // ---- sampling pass (with a pipelined upload: over the first chunk, as soon as it has arrived) ----
    const int np = sm->n_pending;
    auto tiles_upto = [&](int ci) -> u64 {
        if (ci == np - 1) return ta.n_tiles;
        u64 bases_ok = sm->pend_words_end[ci] * 32;
        u64 upto = bases_ok > s->base0 + bp + 192 ? (bases_ok - 192 - s->base0) / bp : 0;
        return std::min<u64>(upto, ta.n_tiles);
    };
    u64 t_s = ta.n_tiles;
    if (np > 0) {
        CU(c, cudaStreamWaitEvent(st, sm->pend_ev[0], 0));
        t_s = tiles_upto(0);
    }
    // thinning: every 16th tile when the buckets are large; small buckets (many buckets per k-mer: multi-GPU jobs) need a denser
    // sample, or the capacity of a bucket the sample happened to miss falls below its true size (measured at 2^19 buckets of
    // ~140 records, 1 : 16: a few of the 524 288 regions overflowed on most calls and the whole partition was redone by the
    // staging path).  >= ~32 sampled records per bucket keeps that below 1e-9 per bucket.
    const double rec_per_bucket = 2.0 * (double)N / (double)(k - p + 2) / (double)NB;
    const u64 thin = (u64)std::min(16.0, std::max(1.0, rec_per_bucket / 32.0));
    const u32 stride = (u32)std::max<u64>(1, t_s * thin / ta.n_tiles);
    const u64 n_sampled = (t_s + stride - 1) / stride;
    const float scale = (float)((double)ta.n_tiles / (double)std::max<u64>(n_sampled, 1));
    {
        TileArgs tc = ta;
        tc.n_tiles = t_s; tc.tstride = stride;
        a.mode = 1;
        msp_tile_kernel<W><<<(u32)std::min<u64>(std::max<u64>(n_sampled, 1), (u64)c->sm_count * WT_CTAS), WT_THREADS, 0, st>>>(kp, a, tc);
        TRY(check_launch(c, "msp_partition_sample"));
    }
    bucket_caps_kernel<<<grid_for(NB, 256), 256, 0, st>>>(d.cnt.p, NB, scale, d.cap.p);
    TRY(check_launch(c, "bucket_caps"));
    TRY(exclusive_scan_u32_to_u64(c, d.cap.p, d.bucket_start.p, NB, d.ctr.p));
    check_total_kernel<<<1, 1, 0, st>>>(d.ctr.p, d.rec_bound, (u32*)(d.ctr.p + 1));
    TRY(check_launch(c, "check_total"));
    // ---- main pass: records straight into their bucket's region (ev[8..9] time this launch alone) ----
    a.mode = 2;
    CU(c, cudaEventRecord(c->ev[8], st));
    if (np > 0) {
        u64 done = 0;
        for (int ci = 0; ci < np; ci++) {
            CU(c, cudaStreamWaitEvent(st, sm->pend_ev[ci], 0));
            const u64 upto = tiles_upto(ci);
            if (upto > done) {
                TileArgs tc = ta;
                tc.tile0 = done; tc.n_tiles = upto;
                msp_tile_kernel<W><<<(u32)std::min<u64>(upto - done, (u64)c->sm_count * WT_CTAS), WT_THREADS, 0, st>>>(kp, a, tc);
                TRY(check_launch(c, "msp_partition"));
                done = upto;
            }
        }
        sm->n_pending = 0;
    } else {
        msp_tile_kernel<W><<<(u32)std::min<u64>(ta.n_tiles, (u64)c->sm_count * WT_CTAS), WT_THREADS, 0, st>>>(kp, a, ta);
        TRY(check_launch(c, "msp_partition"));
    }
    CU(c, cudaEventRecord(c->ev[9], st));
    bucket_fill_final_kernel<<<(u32)std::min<u64>(grid_for(NB, 256), (u64)c->sm_count * 4), 256, 0, st>>>(d.fill.p, d.cap.p, d.bucket_start.p, d.rec_bound, NB, d.cnt.p, d.ctr.p + 2);
    TRY(check_launch(c, "bucket_fill_final"));
    return DBG_OK;
}

// Valid / distinct k-mers emitted by the counting stage (unordered), kept across the passes of one filter call.
struct CountOut {
    DBuf<u64> v_lo, v_hi, a_lo, a_hi, ctr;
    DBuf<u32> v_val;
    u64 cap_valid = 0, cap_all = 0;
    u64 n_valid = 0, n_all = 0, n_splits = 0, n_rec_distinct = 0;
};

template <int W>
static int count_alloc(Ctx* c, u64 N, u32 min_obs, int report_all, bool exact_bound, CountOut& co) {
    // Valid k-mers are a small fraction of the occurrences on real coverage (V/N ~ 0.03 at 50x): size the output by
    // an estimate and fall back to the exact bound (every valid k-mer needs >= min_obs occurrences) if it overflows.
    const u64 bound_valid = min_obs > 1 ? N / min_obs + 1 : N;
    co.cap_valid = exact_bound ? bound_valid
                               : std::min<u64>(bound_valid, c->valid_est_div ? N / c->valid_est_div + 16 : N / 8 + (1u << 20));
    co.cap_all = report_all ? N : 0;
    TRY(co.ctr.alloc(c, 8));
    TRY(co.ctr.zero());
    TRY(co.v_lo.alloc(c, co.cap_valid));
    TRY(co.v_val.alloc(c, co.cap_valid));
    if (W == 2) TRY(co.v_hi.alloc(c, co.cap_valid));
    if (report_all) {
        TRY(co.a_lo.alloc(c, co.cap_all));
        if (W == 2) TRY(co.a_hi.alloc(c, co.cap_all));
    }
    return DBG_OK;
}

// ---- P2: bucket-contiguous records -> (unordered) valid k-mers appended to co.  Returns DBG_OK and sets
// *overflow when the valid / all buffers were too small (the caller retries with the exact bound). ----
template <int W>
static int count_stage(Ctx* c, int k, u64* rec, u64 n_rec, const u64* bucket_start, const u32* bucket_cnt, u32 NB, u32 min_obs,
                       int stranded, int report_all, CountOut& co, bool* overflow) {
    KP kp = make_kp(k);
    cudaStream_t st = c->stream;
    *overflow = false;
    DBuf<u32> mult, dedup_cnt;
    const bool dedup = c->dedup >= (W == 1 ? 1 : 2);   // 32-byte records dedupe only ~1.4x (K=63): opt-in (dedup=2)
    if (dedup) { TRY(mult.alloc(c, n_rec)); TRY(dedup_cnt.alloc(c, NB)); }
    // per-pass counters: bucket queue [0], splits [3], error [4], distinct records [5]; [1], [2] = output cursors persist
    CU(c, cudaMemsetAsync(co.ctr.p, 0, 8, st));
    CU(c, cudaMemsetAsync(co.ctr.p + 3, 0, 24, st));
    P2Args a;
    a.rec = rec; a.mult = dedup ? mult.p : nullptr; a.mult_ready = 0; a.dedup_cnt = dedup_cnt.p;
    a.bucket_start = bucket_start; a.bucket_cnt = bucket_cnt; a.n_buckets = NB;
    a.min_obs = min_obs; a.stranded = stranded; a.report_all = report_all;
    // k-mers per task: 8 keeps ~one task per thread and chunk for both record sizes (32-byte records hold up to 59 k-mers; tasks of 16
    // left half of the 1024 threads of a K > 32 CTA without work: buckets of ~350 records); records longer than 8 tasks fall back to 16
    a.task_len = rec_max_kmers(RecLayout<W>::WORDS, k) <= (W == 1 ? 32 : 64) ? 8 : 16;
    a.out_lo = co.v_lo.p; a.out_hi = co.v_hi.p; a.out_val = co.v_val.p; a.cap_valid = co.cap_valid;
    a.all_lo = co.a_lo.p; a.all_hi = co.a_hi.p; a.cap_all = co.cap_all;
    a.counters = co.ctr.p;
    size_t smem = (sizeof(Kmer<W>) + 4) * P2Cfg<W>::CAP;
    CU(c, cudaFuncSetAttribute(count_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    u32 grid2 = (u32)std::min<u64>(NB, (u64)c->sm_count * P2Cfg<W>::CTAS);
    CU(c, cudaEventRecord(c->ev[10], st));
    count_kernel<W><<<grid2, P2Cfg<W>::THREADS, smem, st>>>(kp, a);
    TRY(check_launch(c, "count_kernel"));
    CU(c, cudaEventRecord(c->ev[11], st));
    u64 h[6];
    TRY(read_u64(c, co.ctr.p, h, 6));
    if (h[4] == 2) { *overflow = true; return DBG_OK; }
    if (h[4]) DBG_SET_ERR(c, DBG_E_INTERNAL, "count_kernel failed (code %llu)", (unsigned long long)h[4]);
    co.n_valid = h[1];
    co.n_all = h[2];
    co.n_splits += h[3];
    if (dedup) co.n_rec_distinct += h[5];
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]);
    c->stats.ms_k_count += ms;
    return DBG_OK;
}

// ---- P3: ascending order (src/filter.rs:205-219) ----
template <int W>
static int sort_stage(Ctx* c, int k, int report_all, CountOut& co, Table* t) {
    cudaStream_t st = c->stream;
    dbg_stats& S = c->stats;
    const u64 V = co.n_valid, U = co.n_all;
    S.n_valid = V;
    S.n_distinct = U;
    S.n_bucket_splits = co.n_splits;
    S.n_records_distinct = co.n_rec_distinct;
    t->n = V;
    if (V) {
        DBuf<u64> b_lo, b_hi;
        DBuf<u32> b_val;
        TRY(b_lo.alloc_pool(c, V));
        TRY(b_val.alloc(c, V));
        if (W == 2) TRY(b_hi.alloc_pool(c, V));
        u64 *rlo, *rhi;
        u32* rval;
        TRY(radix_sort_pairs(c, W, 2 * k, V, co.v_lo.p, co.v_hi.p, co.v_val.p, b_lo.p, b_hi.p, b_val.p, &rlo, &rhi, &rval));
        // keep right-sized arrays: take the V-sized buffer when the result landed there, else copy out of the bound-sized one
        DBuf<u8> d_exts;
        DBuf<u16> d_counts;
        TRY(d_exts.alloc_pool(c, V));
        TRY(d_counts.alloc_pool(c, V));
        unpack_vals_kernel<<<grid_for(V, 256), 256, 0, st>>>(rval, d_exts.p, d_counts.p, V);
        TRY(check_launch(c, "unpack_vals"));
        if (rlo != b_lo.p) {
            CU(c, cudaMemcpyAsync(b_lo.p, rlo, V * 8, cudaMemcpyDeviceToDevice, st));
            if (W == 2) CU(c, cudaMemcpyAsync(b_hi.p, rhi, V * 8, cudaMemcpyDeviceToDevice, st));
        }
        t->lo = b_lo.take();
        if (W == 2) t->hi = b_hi.take();
        t->exts = d_exts.take();
        t->counts = d_counts.take();
    }
    if (report_all && U) {
        // all_kmers: every distinct k-mer, ascending (src/filter.rs:210-212)
        DBuf<u64> b_lo, b_hi;
        DBuf<u32> dummy_a, dummy_b;
        TRY(b_lo.alloc_pool(c, U));
        if (W == 2) TRY(b_hi.alloc_pool(c, U));
        TRY(dummy_a.alloc(c, U));
        TRY(dummy_b.alloc(c, U));
        u64 *rlo, *rhi;
        u32* rval;
        TRY(radix_sort_pairs(c, W, 2 * k, U, co.a_lo.p, co.a_hi.p, dummy_a.p, b_lo.p, b_hi.p, dummy_b.p, &rlo, &rhi, &rval));
        if (rlo != b_lo.p) {
            CU(c, cudaMemcpyAsync(b_lo.p, rlo, U * 8, cudaMemcpyDeviceToDevice, st));
            if (W == 2) CU(c, cudaMemcpyAsync(b_hi.p, rhi, U * 8, cudaMemcpyDeviceToDevice, st));
        }
        t->all_lo = b_lo.take();
        if (W == 2) t->all_hi = b_hi.take();
        t->n_all = U;
    }
    CU(c, cudaEventRecord(c->ev[3], st));
    return DBG_OK;
}

// Scratch bytes one pass needs per input k-mer occurrence handled in it (staging 5 + records 2 + dedup 0.5 + slack),
// plus the job-wide valid-k-mer buffers (~1.5 B per occurrence).  Used by the pass planner only.
static int plan_passes(Ctx* c, u64 N, u64 mem_gb, u32 NB) {
    u64 budget = c->mem_budget_bytes;
    if (!budget) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        budget = (u64)((double)(fr + c->arena_size) * 0.6);        // what the arena may grow to
        if (mem_gb) budget = std::min<u64>(budget, mem_gb * 1000000000ull);  // the reference's memory_size (filter.rs:151-158)
    }
    u64 fixed = (u64)(1.5 * (double)N), per = (u64)(8.0 * (double)N);
    if (budget <= fixed + (64u << 20)) return (int)std::min<u64>(NB, 64);
    u64 passes = (per + (budget - fixed) - 1) / (budget - fixed);
    if (passes < 1) passes = 1;
    if (passes > NB) passes = NB;
    return (int)passes;
}

template <int W>
static int filter_impl(Ctx* c, int k, const SeqSet* s, u32 min_obs, int stranded, int report_all, u64 mem_gb, Table** out) {
    cudaStream_t st = c->stream;
    dbg_stats& S = c->stats;
    TRY(arena_begin(c));
    CU(c, cudaEventRecord(c->ev[0], st));
    u64 N = 0;
    u32 max_len = 0;
    TRY(count_input(c, k, s, &N, &max_len));
    Table* t = &(new dbg_kmer_table())->t;
    t->ctx = c;
    t->k = k;
    t->n_input = N;
    S.n_seqs = s->n_seqs;
    S.n_input_kmers = N;
    *out = t;
    if (N == 0) {
        S.n_records = S.n_buckets = S.n_distinct = S.n_valid = 0;
        return DBG_OK;
    }
    int p, bbits;
    plan_filter(c, k, N, &p, &bbits);
    const u32 NB = 1u << bbits;
    S.msp_p = p; S.bucket_bits = bbits; S.n_buckets = NB;
    // Pass planner: when one pass over all buckets would not fit the scratch budget, the buckets are split into ranges
    // and the reads are re-scanned once per range — the reference's own scheme (filter.rs:151-203: "slices" of the 256
    // prefix buckets sized by memory_size), with the same property: the number of passes never changes the result.
    const int n_pass = plan_passes(c, N, mem_gb, NB);
    S.n_passes = n_pass;
    float ms_part = 0, ms_cnt = 0;
    // direct partition (no staging, no scatter pass): contiguous layouts, one pass, enough tiles for the sampling pass
    const u64 n_tiles_all = (s->contiguous && s->total_end > s->base0) ? (s->total_end - s->base0 + TP - 1) / TP : 0;
    // A pipelined upload can only sample its FIRST chunk (the main pass must start before the rest has arrived).  Reads as a
    // sequencer writes them are in no genomic order, so the first eighth is as good a sample as every 16th tile; position-sorted
    // input is not, and there a mispredicted region costs a second partition + count: the context remembers a failed attempt and
    // keeps such callers on the staging path afterwards.
    // (Two-word keys used to keep staging: their buckets hold ~340 records, too few for a 1 : 16 sample; the sampling density now
    // follows the bucket size.)
    const bool pipelined = s->n_pending > 0;
    bool use_direct = c->direct_partition && (W == 1 || W2_DIRECT) && n_pass == 1 && n_tiles_all >= c->direct_min_tiles &&
                      (!pipelined || !c->pipelined_direct_failed);
    S.direct_partition = 0;
    for (int attempt = 0;; attempt++) {
        c->arena_off = 0;
        S.ms_k_count = 0; S.ms_k_partition = 0; S.n_records = 0;
        ms_part = ms_cnt = 0;
        CountOut co;
        TRY(count_alloc<W>(c, N, min_obs, report_all, attempt > 0, co));
        const u64 mark = c->arena_off;
        bool overflow = false, direct_failed = false;
        for (int pass = 0; pass < n_pass && !overflow; pass++) {
            const u32 lo = (u32)((u64)NB * pass / n_pass), hi = (u32)((u64)NB * (pass + 1) / n_pass);
            if (use_direct) {
                DirectOut d;
                CU(c, cudaEventRecord(c->ev[4], st));
                TRY(partition_direct<W>(c, k, s, stranded, N, p, bbits, d));
                CU(c, cudaEventRecord(c->ev[5], st));
                TRY(count_stage<W>(c, k, d.rec.p, d.rec_bound, d.bucket_start.p, d.cnt.p, NB, min_obs, stranded, report_all, co, &overflow));
                CU(c, cudaEventRecord(c->ev[6], st));
                u64 h[3];
                TRY(read_u64(c, d.ctr.p, h, 3));
                if ((u32)h[1]) { direct_failed = true; break; }   // a region (or the total) was too small: redo through staging
                S.n_records += h[2];
                S.direct_partition = 1;
            } else {
                PartOut po;
                CU(c, cudaEventRecord(c->ev[4], st));
                TRY(partition_stage<W>(c, k, s, stranded, N, max_len, p, bbits, lo, hi - lo, false, po));
                S.n_records += po.n_rec;
                CU(c, cudaEventRecord(c->ev[5], st));
                TRY(count_stage<W>(c, k, po.rec.p, po.n_rec, po.bucket_off.p, po.bucket_count.p, NB, min_obs, stranded, report_all, co, &overflow));
                CU(c, cudaEventRecord(c->ev[6], st));
                TRY(sync(c));
            }
            {
                float a1 = 0, a2 = 0, a3 = 0;
                cudaEventElapsedTime(&a1, c->ev[4], c->ev[5]);
                cudaEventElapsedTime(&a2, c->ev[5], c->ev[6]);
                cudaEventElapsedTime(&a3, c->ev[8], c->ev[9]);
                ms_part += a1; ms_cnt += a2; S.ms_k_partition += a3;
            }
            c->arena_off = mark;  // every per-pass buffer is gone
        }
        if (direct_failed) {
            if (pipelined) c->pipelined_direct_failed = 1;
            use_direct = false; attempt--; continue;
        }
        if (!overflow) {
            CU(c, cudaEventRecord(c->ev[2], st));
            TRY(sort_stage<W>(c, k, report_all, co, t));
            break;
        }
        if (attempt == 1) DBG_SET_ERR(c, DBG_E_INTERNAL, "valid k-mer buffer overflow with the exact bound");
    }
    TRY(sync(c));
    S.ms_partition = ms_part;
    S.ms_count = ms_cnt;
    cudaEventElapsedTime(&S.ms_sort, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&S.ms_filter_total, c->ev[0], c->ev[3]);
    S.gpu_launches = c->launches;
    return DBG_OK;
}

int filter_kmers_dev(Ctx* c, int k, const SeqSet* s, u32 min_obs, int stranded, int report_all, u64 mem_gb,
                     Table** out) {
    *out = nullptr;
    if (k < 4 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [4,64] (filter::bucket needs k >= 4, src/filter.rs:18-23)", k);
    if (!s) DBG_SET_ERR(c, DBG_E_BADARG, "null seqset");
    int rc = k <= 32 ? filter_impl<1>(c, k, s, min_obs, stranded, report_all, mem_gb, out)
                     : filter_impl<2>(c, k, s, min_obs, stranded, report_all, mem_gb, out);
    if (rc != DBG_OK && *out) { free_table(*out); *out = nullptr; }
    return rc;
}

// ================================================================================================
// Multi-GPU building blocks: the counting stage shards by MSP bucket.  Every rank
// partitions its own reads with the SAME plan (p, bucket bits from the total k-mer count), ships each
// bucket range to its owning rank (one all-to-all of 16/32-byte super-k-mer records, done by the caller
// with NCCL on these device buffers), and counts the buckets it owns.  All occurrences of a canonical
// k-mer share a bucket, so per-rank tables are disjoint and their union is the unsharded table.
// ================================================================================================
// bucket regions with slack (direct partition) -> bucket-contiguous records: one warp per bucket, coalesced 16-byte copies
template <int RW>
__global__ void __launch_bounds__(256) compact_regions_kernel(const u64* __restrict__ src, const u64* __restrict__ reg_start, const u32* __restrict__ cnt,
                                                               const u64* __restrict__ dst_off, u32 nb, u64* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    for (u64 b = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nb; b += ((u64)gridDim.x * blockDim.x) >> 5) {
        const u64 s0 = reg_start[b], d0 = dst_off[b];
        const u32 n = cnt[b];
        for (u32 i = lane; i < n; i += 32) {
            const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(src + (s0 + i) * RW);
            ulonglong2* dp = reinterpret_cast<ulonglong2*>(dst + (d0 + i) * RW);
            dp[0] = sp[0];
            if (RW == 4) dp[1] = sp[1];
        }
    }
}

int partition_reads_dev(Ctx* c, int k, const SeqSet* s, int stranded, int p, int bbits, Partition** out) {
    *out = nullptr;
    if (k < 4 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [4,64]", k);
    if (bbits < 0 || bbits > 20 || p < 1 || p > 16 || p > k - 3 + (k < 4 ? 0 : 0) || k - p > 63)
        DBG_SET_ERR(c, DBG_E_BADARG, "bad MSP plan p=%d bucket_bits=%d for k=%d", p, bbits, k);
    TRY(arena_begin(c));
    u64 N = 0;
    u32 max_len = 0;
    TRY(count_input(c, k, s, &N, &max_len));
    Partition* P = &(new dbg_partition())->p;
    P->ctx = c; P->k = k; P->p = p; P->bbits = bbits; P->n_input = N; P->rec_words = k <= 32 ? 2 : 4;
    PartOut po;
    int rc = DBG_OK;
    bool done = false;
    // Large contiguous device-resident inputs: the direct partition (sampling pass + records straight into per-bucket
    // regions) followed by ONE streaming pass that closes the gaps between the regions — 2 x 16 bytes per record moved
    // coalesced, against the staging path's scatter of every record through a bucket-cursor atomic.
    const u64 n_tiles_all = (s->contiguous && s->total_end > s->base0) ? (s->total_end - s->base0 + TP - 1) / TP : 0;
    if (N && k <= 32 && c->direct_partition && n_tiles_all >= c->direct_min_tiles && s->n_pending == 0) {
        const u32 NB = 1u << bbits;
        DirectOut d;
        rc = partition_direct<1>(c, k, s, stranded, N, p, bbits, d);
        if (rc == DBG_OK) rc = po.bucket_count.alloc_pool(c, NB);
        if (rc == DBG_OK) rc = po.bucket_off.alloc_pool(c, (u64)NB + 1);
        if (rc == DBG_OK) rc = exclusive_scan_u32_to_u64(c, d.cnt.p, po.bucket_off.p, NB, po.bucket_off.p + NB);
        u64 h[3] = {0, 0, 0};
        if (rc == DBG_OK) rc = read_u64(c, d.ctr.p, h, 3);   // [1] overflow flag, [2] records stored
        if (rc == DBG_OK && !(u32)h[1]) {
            po.n_rec = h[2];
            rc = po.rec.alloc_pool(c, po.n_rec * 2);
            if (rc == DBG_OK) {
                CU(c, cudaMemcpyAsync(po.bucket_count.p, d.cnt.p, (u64)NB * 4, cudaMemcpyDeviceToDevice, c->stream));
                compact_regions_kernel<2><<<(u32)std::min<u64>(((u64)NB + 7) / 8, (u64)c->sm_count * 32), 256, 0, c->stream>>>(
                    d.rec.p, d.bucket_start.p, d.cnt.p, po.bucket_off.p, NB, po.rec.p);
                rc = check_launch(c, "compact_regions");
            }
            done = rc == DBG_OK;
        }
        if (rc != DBG_OK) { delete reinterpret_cast<dbg_partition*>(P); return rc; }
        if (!done) { po.bucket_count.release(); po.bucket_off.release(); }   // a region overflowed: staging path below
    }
    if (done) {
    } else if (N) rc = k <= 32 ? partition_stage<1>(c, k, s, stranded, N, max_len, p, bbits, 0, 1u << bbits, true, po)
                        : partition_stage<2>(c, k, s, stranded, N, max_len, p, bbits, 0, 1u << bbits, true, po);
    else {
        rc = po.bucket_count.alloc_pool(c, 1u << bbits);
        if (rc == DBG_OK) rc = po.bucket_count.zero();
    }
    if (rc != DBG_OK) { delete reinterpret_cast<dbg_partition*>(P); return rc; }
    rc = sync(c);
    if (rc != DBG_OK) { delete reinterpret_cast<dbg_partition*>(P); return rc; }
    if (N) cudaEventElapsedTime(&c->stats.ms_k_partition, c->ev[8], c->ev[9]);
    P->n_rec = po.n_rec;
    P->rec = po.rec.p ? po.rec.take() : nullptr;
    P->bucket_count = po.bucket_count.take();
    if (po.bucket_off.p) P->bucket_off = po.bucket_off.take();
    c->stats.n_records = po.n_rec;
    c->stats.n_input_kmers = N;
    *out = P;
    return DBG_OK;
}

// ---- fused compaction + exchange (multi.cu): the partition's bucket regions are NOT made contiguous on the sending rank;
// one kernel copies every bucket's records straight to their final, bucket-contiguous position in the OWNING rank's
// receive window over peer memory (NVLink stores), so that neither a local gap-closing pass, nor a send buffer, nor a
// per-source merge on the receiver is needed. ----
int partition_regions_dev(Ctx* c, int k, const SeqSet* s, int stranded, int p, int bbits, PartRegions** out) {
    *out = nullptr;
    if (k < 4 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [4,64]", k);
    if (bbits < 0 || bbits > 20 || p < 1 || p > 16 || p > k - 3 || k - p > 63)
        DBG_SET_ERR(c, DBG_E_BADARG, "bad MSP plan p=%d bucket_bits=%d for k=%d", p, bbits, k);
    TRY(arena_begin(c));
    u64 N = 0;
    u32 max_len = 0;
    TRY(count_input(c, k, s, &N, &max_len));
    const u32 NB = 1u << bbits;
    PartRegions* R = new PartRegions();
    R->ctx = c; R->k = k; R->p = p; R->bbits = bbits; R->rec_words = k <= 32 ? 2 : 4; R->n_input = N;
    struct Guard { PartRegions* r; ~Guard() { if (r) free_part_regions(r); } } guard{R};
    const u64 n_tiles_all = (s->contiguous && s->total_end > s->base0) ? (s->total_end - s->base0 + TP - 1) / TP : 0;
    const bool pipelined = s->n_pending > 0;   // (first-chunk sampling, remembered fallback: see filter_impl)
    if (N && k <= 32 && c->direct_partition && n_tiles_all >= c->direct_min_tiles && (!pipelined || !c->pipelined_direct_failed)) {
        DirectOut* d = new DirectOut();
        R->holder = d; R->direct = 1;
        TRY(partition_direct<1>(c, k, s, stranded, N, p, bbits, *d));
        u64 h[3] = {0, 0, 0};
        TRY(read_u64(c, d->ctr.p, h, 3));   // [1] overflow flag, [2] records stored
        if (!(u32)h[1]) {
            R->rec = d->rec.p; R->start = d->bucket_start.p; R->cnt = d->cnt.p; R->n_rec = h[2];
        } else {   // a region overflowed: staging path below
            if (getenv("DBG_MULTI_TRACE")) {   // which regions, and by how much
                std::vector<u32> hf(NB), hc(NB);
                std::vector<u64> hs(NB + 1);
                cudaMemcpy(hf.data(), d->fill.p, (u64)NB * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(hc.data(), d->cap.p, (u64)NB * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(hs.data(), d->bucket_start.p, ((u64)NB + 1) * 8, cudaMemcpyDeviceToHost);
                u64 nover = 0, sumcap = 0, sumfill = 0;
                for (u32 b = 0; b < NB; b++) {
                    sumcap += hc[b]; sumfill += hf[b];
                    if (hf[b] > hc[b] && nover++ < 8) fprintf(stderr, "[dbg multi] region %u: fill %u > cap %u\n", b, hf[b], hc[b]);
                }
                fprintf(stderr, "[dbg multi] direct partition overflow: %llu regions over, sum cap %llu (bound %llu), records %llu, flag %llu\n",
                        (unsigned long long)nover, (unsigned long long)sumcap, (unsigned long long)d->rec_bound, (unsigned long long)sumfill, (unsigned long long)h[1]);
            }
            delete d;
            R->holder = nullptr; R->direct = 0;
            if (pipelined) c->pipelined_direct_failed = 1;
        }
    }
    if (!R->rec) {
        PartOut* po = new PartOut();
        R->holder = po; R->direct = 0;
        if (N) {
            TRY(k <= 32 ? partition_stage<1>(c, k, s, stranded, N, max_len, p, bbits, 0, NB, true, *po)
                        : partition_stage<2>(c, k, s, stranded, N, max_len, p, bbits, 0, NB, true, *po));
        } else {
            TRY(po->bucket_count.alloc_pool(c, NB)); TRY(po->bucket_count.zero());
            TRY(po->bucket_off.alloc_pool(c, (u64)NB + 1)); TRY(po->bucket_off.zero());
            TRY(po->rec.alloc_pool(c, 1));
        }
        R->rec = po->rec.p; R->start = po->bucket_off.p; R->cnt = po->bucket_count.p; R->n_rec = po->n_rec;
    }
    TRY(sync(c));
    if (N) cudaEventElapsedTime(&c->stats.ms_k_partition, c->ev[8], c->ev[9]);
    if (getenv("DBG_MULTI_TRACE")) fprintf(stderr, "[dbg multi] partition: direct=%d records=%llu buckets=2^%d p=%d\n", R->direct, (unsigned long long)R->n_rec, bbits, p);
    c->stats.n_records = R->n_rec;
    c->stats.n_input_kmers = N;
    guard.r = nullptr;
    *out = R;
    return DBG_OK;
}
void free_part_regions(PartRegions* R) {
    if (!R) return;
    if (R->holder) { if (R->direct) delete static_cast<DirectOut*>(R->holder); else delete static_cast<PartOut*>(R->holder); }
    delete R;
}

// tot[b] = records of bucket b over all ranks, pre[b] = records of bucket b on the ranks below `me`
__global__ void bucket_totals_kernel(const u32* __restrict__ all_cnt, int P, int me, u32 nb, u32* __restrict__ tot, u32* __restrict__ pre) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    u32 t = 0, q = 0;
    for (int r = 0; r < P; r++) { const u32 v = all_cnt[(u64)r * nb + b]; if (r < me) q += v; t += v; }
    tot[b] = t; pre[b] = q;
}
int bucket_totals_dev(Ctx* c, const u32* d_all_cnt, int P, int me, u32 nb, u32* d_tot, u32* d_pre) {
    bucket_totals_kernel<<<grid_for(nb, 256), 256, 0, c->stream>>>(d_all_cnt, P, me, nb, d_tot, d_pre);
    return check_launch(c, "bucket_totals");
}
// the owned bucket range [b0, b0 + n) in local coordinates: off[i] = goff[b0 + i] - goff[b0], cnt[i] = tot[b0 + i]
__global__ void local_buckets_kernel(const u64* __restrict__ goff, const u32* __restrict__ tot, u32 b0, u32 n, u64* __restrict__ off,
                                     u32* __restrict__ cnt) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    off[i] = goff[b0 + i] - goff[b0];
    if (i < n) cnt[i] = tot[b0 + i];
}
int local_buckets_dev(Ctx* c, const u64* d_goff, const u32* d_tot, u32 b0, u32 n, u64* d_off, u32* d_cnt) {
    local_buckets_kernel<<<grid_for((u64)n + 1, 256), 256, 0, c->stream>>>(d_goff, d_tot, b0, n, d_off, d_cnt);
    return check_launch(c, "local_buckets");
}

// One warp per bucket: coalesced 16-byte copies from the bucket's region to base[owner] + (goff[b] - goff[bound[owner]] + pre[b]);
// per destination: k-mer occurrences and records shipped (sums[r], sums[P + r]).
template <int RW>
__global__ void __launch_bounds__(256) scatter_buckets_kernel(const u64* __restrict__ src, const u64* __restrict__ src_start,
                                                               const u32* __restrict__ cnt, const u64* __restrict__ goff,
                                                               const u32* __restrict__ pre, ScatterDst D, u32 nb, u64* __restrict__ sums) {
    __shared__ unsigned long long s_km[DBG_MAX_RANKS], s_rc[DBG_MAX_RANKS];
    if (threadIdx.x < DBG_MAX_RANKS) { s_km[threadIdx.x] = 0; s_rc[threadIdx.x] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // Work item j -> destination (j + me) mod P, bucket j / P of that destination's range: at any moment a sender writes to ALL
    // destinations, and the senders are rotated against each other.  (Sweeping the buckets in order sends the whole GPU — and every
    // other GPU at the same time — to one destination after the other: 8 senders queueing on one NVLink ingress, the other 7 idle.)
    u64 max_n = 0;
    for (int r = 0; r < D.P; r++) max_n = max(max_n, D.bound[r + 1] - D.bound[r]);
    const u64 n_items = max_n * (u64)D.P;
    for (u64 j = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_items; j += ((u64)gridDim.x * blockDim.x) >> 5) {
        const int r = (int)((j + (u64)D.me) % (u64)D.P);
        const u64 bi = j / (u64)D.P;
        if (bi >= D.bound[r + 1] - D.bound[r]) continue;
        const u64 b = D.bound[r] + bi;
        const u32 n = cnt[b];
        if (!n) continue;
        const u64 s0 = src_start[b], d0 = goff[b] - goff[D.bound[r]] + pre[b];
        u64* dst = D.base[r];
        u32 km = 0;
        // four records per lane in flight: the stores are posted, the loads in front of them are what a warp waits for
        for (u32 i0 = 0; i0 < n; i0 += 128) {
            ulonglong2 v[4][RW / 2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const u32 i = i0 + 32 * u + lane;
                if (i < n) {
                    const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(src + (s0 + i) * RW);
                    v[u][0] = sp[0];
                    if (RW == 4) v[u][RW / 2 - 1] = sp[1];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const u32 i = i0 + 32 * u + lane;
                if (i < n) {
                    ulonglong2* dp = reinterpret_cast<ulonglong2*>(dst + (d0 + i) * RW);
                    dp[0] = v[u][0];
                    if (RW == 4) dp[1] = v[u][RW / 2 - 1];
                    km += (u32)(v[u][RW / 2 - 1].y >> 8) & 63u;
                }
            }
        }
        for (int o = 16; o; o >>= 1) km += __shfl_down_sync(0xffffffffu, km, o);
        if (lane == 0) { atomicAdd(&s_km[r], (unsigned long long)km); atomicAdd(&s_rc[r], (unsigned long long)n); }
    }
    __syncthreads();
    if (threadIdx.x < D.P) {
        if (s_km[threadIdx.x]) atomicAdd(&sums[threadIdx.x], (u64)s_km[threadIdx.x]);
        if (s_rc[threadIdx.x]) atomicAdd(&sums[D.P + threadIdx.x], (u64)s_rc[threadIdx.x]);
    }
}
int scatter_buckets_dev(Ctx* c, const PartRegions* R, const u64* d_goff, const u32* d_pre, const ScatterDst& D, u64* d_sums) {
    const u32 nb = 1u << R->bbits;
    const u32 grid = (u32)std::min<u64>(((u64)nb + 7) / 8, (u64)c->sm_count * 32);
    if (R->rec_words == 2) scatter_buckets_kernel<2><<<grid, 256, 0, c->stream>>>(R->rec, R->start, R->cnt, d_goff, d_pre, D, nb, d_sums);
    else scatter_buckets_kernel<4><<<grid, 256, 0, c->stream>>>(R->rec, R->start, R->cnt, d_goff, d_pre, D, nb, d_sums);
    return check_launch(c, "scatter_buckets");
}

// Counting + sorting of bucket-contiguous records that are already in place (the receive window of the fused exchange).
// d_off / d_cnt: device arrays of the n_local owned buckets (not arena memory).  n_kmers_local = k-mer occurrences held by the
// records: the valid-k-mer buffer is sized by the exact bound, so the in-place deduplication never has to be undone.
int filter_from_bucketed_dev(Ctx* c, int k, u64* d_records, u64 n_records, const u64* d_off, const u32* d_cnt, u32 n_local,
                             u64 n_kmers_local, u64 n_input_total, u32 min_obs, int stranded, int report_all, Table** out) {
    *out = nullptr;
    if (k < 4 || k > 64 || !n_local) DBG_SET_ERR(c, DBG_E_BADARG, "bad arguments");
    cudaStream_t st = c->stream;
    dbg_stats& S = c->stats;
    TRY(arena_begin(c));
    CU(c, cudaEventRecord(c->ev[1], st));
    Table* t = &(new dbg_kmer_table())->t;
    t->ctx = c; t->k = k; t->n_input = n_input_total;
    *out = t;
    S.n_records = n_records;
    S.n_input_kmers = n_kmers_local;
    S.ms_k_count = 0;
    if (n_records == 0) return DBG_OK;
    int rc = DBG_OK;
    {
        CountOut co;
        bool overflow = false;
        if (k <= 32) {
            rc = count_alloc<1>(c, n_kmers_local, min_obs, report_all, true, co);
            if (rc == DBG_OK) rc = count_stage<1>(c, k, d_records, n_records, d_off, d_cnt, n_local, min_obs, stranded, report_all, co, &overflow);
            if (rc == DBG_OK && !overflow) { cudaEventRecord(c->ev[2], st); rc = sort_stage<1>(c, k, report_all, co, t); }
        } else {
            rc = count_alloc<2>(c, n_kmers_local, min_obs, report_all, true, co);
            if (rc == DBG_OK) rc = count_stage<2>(c, k, d_records, n_records, d_off, d_cnt, n_local, min_obs, stranded, report_all, co, &overflow);
            if (rc == DBG_OK && !overflow) { cudaEventRecord(c->ev[2], st); rc = sort_stage<2>(c, k, report_all, co, t); }
        }
        if (rc == DBG_OK && overflow) { c->err = "valid k-mer buffer overflow with the exact bound"; rc = DBG_E_INTERNAL; }
    }
    if (rc != DBG_OK) { free_table(t); *out = nullptr; return rc; }
    TRY(sync(c));
    cudaEventElapsedTime(&S.ms_count, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&S.ms_sort, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&S.ms_k_count, c->ev[10], c->ev[11]);
    S.gpu_launches = c->launches;
    return DBG_OK;
}

// (defined above partition_reads_dev's first use)
void free_partition(Partition* P) {
    if (!P) return;
    cudaStream_t st = P->ctx->stream;
    if (P->rec) cudaFreeAsync(P->rec, st);
    if (P->bucket_count) cudaFreeAsync(P->bucket_count, st);
    if (P->bucket_off) cudaFreeAsync(P->bucket_off, st);
    delete reinterpret_cast<dbg_partition*>(P);
}

// One warp per (source, bucket) group: copy the group's records from the received per-source runs into the
// bucket-contiguous layout.
template <int RW>
__global__ void merge_runs_kernel(const u64* __restrict__ src, u64* __restrict__ dst, const u64* __restrict__ g_src,
                                  const u64* __restrict__ g_dst, const u32* __restrict__ g_cnt, u64 n_groups) {
    u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    u64 so = g_src[g], d0 = g_dst[g];
    u32 cnt = g_cnt[g];
    for (u32 i = lane; i < cnt; i += 32) {
        const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(src + (so + i) * RW);
        ulonglong2* dp = reinterpret_cast<ulonglong2*>(dst + (d0 + i) * RW);
        dp[0] = sp[0];
        if (RW == 4) dp[1] = sp[1];
    }
}

template <int RW>
__global__ void __launch_bounds__(256) sum_record_kmers_kernel(const u64* __restrict__ rec, u64 n, u64* out) {
    // grid-stride + one atomic per CTA: same-address L2 atomics serialise, a per-warp atomic would dominate the kernel
    __shared__ u64 s_w[8];
    u64 v = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        v += (rec[i * RW + RW - 1] >> 8) & 63ull;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        if (t) atomicAdd(out, t);
    }
}

// d_records: n_src runs back to back, run s holding the records of this rank's n_local buckets in bucket order;
// h_counts[s * n_local + b] = records of local bucket b in run s.
int filter_from_records_dev(Ctx* c, int k, const u64* d_records, u64 n_records, const u32* h_counts, u32 n_src, u32 n_local,
                            u64 n_input_total, u32 min_obs, int stranded, int report_all, Table** out) {
    *out = nullptr;
    if (k < 4 || k > 64 || !n_src || !n_local || !h_counts) DBG_SET_ERR(c, DBG_E_BADARG, "bad arguments");
    const int RW = k <= 32 ? 2 : 4;
    cudaStream_t st = c->stream;
    dbg_stats& S = c->stats;
    TRY(arena_begin(c));
    CU(c, cudaEventRecord(c->ev[1], st));
    // host: bucket offsets of the merged layout and per-group source/destination offsets
    const u64 n_groups = (u64)n_src * n_local;
    std::vector<u64> h_off(n_local + 1, 0), g_src(n_groups), g_dst(n_groups);
    for (u32 b = 0; b < n_local; b++) {
        u64 t = 0;
        for (u32 sidx = 0; sidx < n_src; sidx++) t += h_counts[(u64)sidx * n_local + b];
        h_off[b + 1] = h_off[b] + t;
    }
    if (h_off[n_local] != n_records) DBG_SET_ERR(c, DBG_E_BADARG, "counts sum to %llu, n_records is %llu", (unsigned long long)h_off[n_local], (unsigned long long)n_records);
    {
        u64 run = 0;
        std::vector<u64> fill(n_local, 0);
        for (u32 sidx = 0; sidx < n_src; sidx++)
            for (u32 b = 0; b < n_local; b++) {
                u64 g = (u64)sidx * n_local + b;
                g_src[g] = run;
                g_dst[g] = h_off[b] + fill[b];
                fill[b] += h_counts[g];
                run += h_counts[g];
            }
    }
    Table* t = &(new dbg_kmer_table())->t;
    t->ctx = c; t->k = k; t->n_input = n_input_total;
    *out = t;
    S.n_records = n_records;
    if (n_records == 0) return DBG_OK;
    DBuf<u64> d_off, d_gsrc, d_gdst, merged;
    DBuf<u32> d_gcnt, d_bcnt;
    std::vector<u32> h_bcnt(n_local);
    for (u32 b = 0; b < n_local; b++) h_bcnt[b] = (u32)(h_off[b + 1] - h_off[b]);
    TRY(d_bcnt.alloc(c, n_local));
    CU(c, cudaMemcpyAsync(d_bcnt.p, h_bcnt.data(), (u64)n_local * 4, cudaMemcpyHostToDevice, st));
    TRY(d_off.alloc(c, n_local + 1)); TRY(d_gsrc.alloc(c, n_groups)); TRY(d_gdst.alloc(c, n_groups)); TRY(d_gcnt.alloc(c, n_groups));
    TRY(merged.alloc(c, n_records * RW));
    CU(c, cudaMemcpyAsync(d_off.p, h_off.data(), (n_local + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(d_gsrc.p, g_src.data(), n_groups * 8, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(d_gdst.p, g_dst.data(), n_groups * 8, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(d_gcnt.p, h_counts, n_groups * 4, cudaMemcpyHostToDevice, st));
    if (RW == 2) merge_runs_kernel<2><<<grid_for(n_groups * 32, 256), 256, 0, st>>>(d_records, merged.p, d_gsrc.p, d_gdst.p, d_gcnt.p, n_groups);
    else merge_runs_kernel<4><<<grid_for(n_groups * 32, 256), 256, 0, st>>>(d_records, merged.p, d_gsrc.p, d_gdst.p, d_gcnt.p, n_groups);
    TRY(check_launch(c, "merge_runs"));
    // k-mer occurrences actually present on this rank (capacity bounds of the counting stage)
    DBuf<u64> d_sum;
    TRY(d_sum.alloc(c, 1));
    TRY(d_sum.zero());
    const u32 sgrid = (u32)std::min<u64>(grid_for(n_records, 256), (u64)c->sm_count * 8);
    if (RW == 2) sum_record_kmers_kernel<2><<<sgrid, 256, 0, st>>>(merged.p, n_records, d_sum.p);
    else sum_record_kmers_kernel<4><<<sgrid, 256, 0, st>>>(merged.p, n_records, d_sum.p);
    TRY(check_launch(c, "sum_record_kmers"));
    u64 N_local_bound = 0;
    TRY(read_u64(c, d_sum.p, &N_local_bound));  // also: host vectors may go out of scope after this sync
    S.n_input_kmers = N_local_bound;
    int rc = DBG_OK;
    S.ms_k_count = 0;
    for (int attempt = 0; attempt < 2 && rc == DBG_OK; attempt++) {
        const u64 mark = c->arena_off;
        CountOut co;
        bool overflow = false;
        if (k <= 32) {
            rc = count_alloc<1>(c, N_local_bound, min_obs, report_all, attempt > 0, co);
            if (rc == DBG_OK) rc = count_stage<1>(c, k, merged.p, n_records, d_off.p, d_bcnt.p, n_local, min_obs, stranded, report_all, co, &overflow);
            if (rc == DBG_OK && !overflow) { cudaEventRecord(c->ev[2], st); rc = sort_stage<1>(c, k, report_all, co, t); }
        } else {
            rc = count_alloc<2>(c, N_local_bound, min_obs, report_all, attempt > 0, co);
            if (rc == DBG_OK) rc = count_stage<2>(c, k, merged.p, n_records, d_off.p, d_bcnt.p, n_local, min_obs, stranded, report_all, co, &overflow);
            if (rc == DBG_OK && !overflow) { cudaEventRecord(c->ev[2], st); rc = sort_stage<2>(c, k, report_all, co, t); }
        }
        if (!overflow) break;
        if (attempt == 1) { c->err = "valid k-mer buffer overflow with the exact bound"; rc = DBG_E_INTERNAL; }
        c->arena_off = mark;
        // the records were deduplicated in place by the first attempt: not reusable as input -> merge again
        if (RW == 2) merge_runs_kernel<2><<<grid_for(n_groups * 32, 256), 256, 0, st>>>(d_records, merged.p, d_gsrc.p, d_gdst.p, d_gcnt.p, n_groups);
        else merge_runs_kernel<4><<<grid_for(n_groups * 32, 256), 256, 0, st>>>(d_records, merged.p, d_gsrc.p, d_gdst.p, d_gcnt.p, n_groups);
        c->launches++;
    }
    if (rc != DBG_OK) { free_table(t); *out = nullptr; return rc; }
    TRY(sync(c));
    cudaEventElapsedTime(&S.ms_count, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&S.ms_sort, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&S.ms_k_count, c->ev[10], c->ev[11]);
    S.gpu_launches = c->launches;
    return DBG_OK;
}

// ------------------------------------------------------------------------------------------------
// msp::msp_sequence bucket of every k-mer under the reference's DEFAULT (identity) permutation:
// bucket(k-mer) = min_rc(argmin p-mer).to_u64() (src/msp.rs:115-117, 305-311).  With rc = !stranded and an
// injective score this is a pure function of the k-mer: the smallest (canonical) p-mer value in its window.
// ------------------------------------------------------------------------------------------------
__global__ void msp_bucket_kernel(const u64* __restrict__ words, u64 n_words, const u64* __restrict__ start,
                                  const u32* __restrict__ length, u32 uniform_len, u64 n_seqs, const u64* __restrict__ koff,
                                  u64 n_out, int k, int p, int stranded, u32* __restrict__ out) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    u64 lo = 0, hi = n_seqs;  // last sequence with koff[i] <= t
    while (lo < hi) { u64 m = (lo + hi) >> 1; if (koff[m] <= t) lo = m + 1; else hi = m; }
    u64 si = lo - 1;
    u64 st = uniform_len ? si * (u64)uniform_len : start[si];
    u64 j = t - koff[si];
    u32 best = 0xffffffffu, bestv = 0;
    for (int q = 0; q <= k - p; q++) {
        u64 b = st + j + q;
        u64 wi = b >> 5;
        int sh = (int)(b & 31) * 2;
        u64 h = wi < n_words ? words[wi] : 0, l = (wi + 1) < n_words ? words[wi + 1] : 0;
        u64 v = sh ? (h << sh) | (l >> (64 - sh)) : h;
        u32 x = (u32)(v >> (64 - 2 * p));
        u32 r = (~rev2_32(x)) >> (32 - 2 * p);
        u32 canon = x < r ? x : r;
        u32 score = stranded ? x : canon;  // rc = !stranded (msp.rs:305-311)
        if (score < best) { best = score; bestv = canon; }   // bucket() canonicalises regardless (msp.rs:115-117)
    }
    out[t] = bestv;
}

__global__ void kmer_count_kernel(const u32* __restrict__ length, u32 uniform_len, u64 n, int k, u32* cnt) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 L = uniform_len ? uniform_len : length[i];
    cnt[i] = L >= (u32)k ? L - k + 1 : 0;
}

int msp_kmer_buckets_dev(Ctx* c, int k, int p, const SeqSet* s, int stranded, u32* h_out, u64 n_out) {
    if (p < 1 || p > 16 || p > k || k > 64 || k < 2) DBG_SET_ERR(c, DBG_E_BADARG, "need 1 <= p <= min(k,16), 2 <= k <= 64");
    TRY(arena_begin(c));
    cudaStream_t st = c->stream;
    if (!s->n_seqs) {
        if (n_out) DBG_SET_ERR(c, DBG_E_BADARG, "n_out does not match the number of k-mers (0)");
        return DBG_OK;
    }
    DBuf<u32> cnt, d_out;
    DBuf<u64> koff, tot;
    TRY(cnt.alloc(c, s->n_seqs)); TRY(koff.alloc(c, s->n_seqs)); TRY(tot.alloc(c, 1));
    kmer_count_kernel<<<grid_for(s->n_seqs, 256), 256, 0, st>>>(s->length, s->uniform_len, s->n_seqs, k, cnt.p);
    TRY(check_launch(c, "kmer_count"));
    TRY(exclusive_scan_u32_to_u64(c, cnt.p, koff.p, s->n_seqs, tot.p));
    u64 total = 0;
    TRY(read_u64(c, tot.p, &total));
    if (total != n_out) DBG_SET_ERR(c, DBG_E_BADARG, "n_out=%llu but the sequences hold %llu k-mers", (unsigned long long)n_out, (unsigned long long)total);
    if (!total) return DBG_OK;
    TRY(d_out.alloc(c, total));
    msp_bucket_kernel<<<grid_for(total, 256), 256, 0, st>>>(s->words, s->n_words, s->start, s->length, s->uniform_len, s->n_seqs,
                                                            koff.p, total, k, p, stranded, d_out.p);
    TRY(check_launch(c, "msp_bucket"));
    CU(c, cudaMemcpyAsync(h_out, d_out.p, total * 4, cudaMemcpyDeviceToHost, st));
    return sync(c);
}

// ------------------------------------------------------------------------------------------------
// msp::msp_sequence (src/msp.rs:279-324): the MSP INTERVALS of every sequence exactly as Scanner::scan (:207-276) cuts
// them — boundaries depend on which occurrence of the minimizer the scan is holding (ties: find_min keeps the largest
// position, MinPos::cmp :127-140; an entering p-mer replaces the held one only when STRICTLY smaller, :244-246), i.e. on
// scan history, so one thread replays the scan of one sequence.  Two launches: count, then write at scanned offsets.
// Reads are short; a genome-long sequence would serialise on its thread (the per-k-mer bucket of dbg_msp_kmer_buckets is
// the history-free form the partition kernels use).
// ------------------------------------------------------------------------------------------------
struct MspSeqArgs {
    const u64* words; u64 n_words; const u64* start; const u32* length; u32 uniform_len; u64 n_seqs;
    int k, p, rc; const u32* perm;
    const u64* off;   // nullptr: count pass (cnt[i] = intervals of sequence i)
    u32* cnt; u32* o_seq; u32* o_start; u32* o_len; u32* o_bucket; u8* o_exts;
};
__device__ __forceinline__ u32 msp_pmer_at(const MspSeqArgs& a, u64 st, u32 pos) {   // p-mer starting at base pos of the sequence
    const u64 b = st + pos;
    const u64 wi = b >> 5;
    const int sh = (int)(b & 31) * 2;
    const u64 h = wi < a.n_words ? a.words[wi] : 0, l = (wi + 1) < a.n_words ? a.words[wi + 1] : 0;
    const u64 v = sh ? (h << sh) | (l >> (64 - sh)) : h;
    return (u32)(v >> (64 - 2 * a.p));
}
__device__ __forceinline__ u32 msp_score(const MspSeqArgs& a, u32 x) {   // msp.rs:305-311
    u32 v = a.perm ? a.perm[x] : x;
    if (a.rc) {
        const u32 r = (~rev2_32(x)) >> (32 - 2 * a.p);
        const u32 vr = a.perm ? a.perm[r] : r;
        v = v < vr ? v : vr;
    }
    return v;
}
__global__ void msp_sequence_kernel(MspSeqArgs a) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_seqs) return;
    const u32 m = a.uniform_len ? a.uniform_len : a.length[i];
    const u64 st = a.uniform_len ? i * (u64)a.uniform_len : a.start[i];
    const int k = a.k, p = a.p;
    if (m < (u32)k) { if (!a.off) a.cnt[i] = 0; return; }   // msp.rs:294-296
    u64 o = a.off ? a.off[i] : 0;
    u32 n_iv = 0;
    // held minimizer
    u32 mpos = 0, mval = 0, mkmer = 0;
    auto find_min = [&](u32 lo, u32 hi) {   // msp.rs:218-228: ties -> the LARGEST position
        mpos = lo; mkmer = msp_pmer_at(a, st, lo); mval = msp_score(a, mkmer);
        for (u32 q = lo + 1; q <= hi; q++) {
            const u32 x = msp_pmer_at(a, st, q), v = msp_score(a, x);
            if (v <= mval) { mval = v; mpos = q; mkmer = x; }
        }
    };
    auto emit = [&](u32 s0, u32 len, u32 kmer) {
        if (a.off) {
            const u32 r = (~rev2_32(kmer)) >> (32 - 2 * p);
            a.o_seq[o] = (u32)i; a.o_start[o] = s0; a.o_len[o] = len;
            a.o_bucket[o] = kmer < r ? kmer : r;                                          // MspIntervalP::bucket, msp.rs:115-117
            const u32 le = s0 > 0 ? 1u << base_at(a.words, st + s0 - 1) : 0u;             // Exts::from_slice_bounds, lib.rs:645-660
            const u32 re = s0 + len < m ? 1u << base_at(a.words, st + s0 + len) : 0u;
            a.o_exts[o] = (u8)((re << 4) | le);
            o++;
        }
        n_iv++;
    };
    find_min(0, (u32)(k - p));
    u32 cur_start = 0, cur_kmer = mkmer;
    for (u32 j = 1; j + k <= m; j++) {
        const u32 epos = j + (u32)(k - p);               // end_pos always corresponds to j + k - p (msp.rs:238-239)
        bool fresh = false;
        if (j > mpos) { find_min(j, epos); fresh = true; }                                    // the minimizer left the window (:241-243)
        else {
            const u32 x = msp_pmer_at(a, st, epos), v = msp_score(a, x);
            if (v < mval) { mval = v; mpos = epos; mkmer = x; fresh = true; }                 // strictly smaller p-mer entered (:244-246)
        }
        if (fresh) {
            emit(cur_start, j + (u32)k - 1 - cur_start, cur_kmer);                            // :253-263
            cur_start = j; cur_kmer = mkmer;
        }
    }
    emit(cur_start, m - cur_start, cur_kmer);                                                 // :266-273
    if (!a.off) a.cnt[i] = n_iv;
}

int msp_sequence_dev(Ctx* c, int k, int p, const SeqSet* s, int rc, const u32* h_perm, u64 cap, u64* n_out, u32* h_seq, u32* h_start,
                     u32* h_len, u32* h_bucket, u8* h_exts) {
    if (p < 1 || p > 16 || p > k || k < 2) DBG_SET_ERR(c, DBG_E_BADARG, "need 1 <= p <= min(k, 16), k >= 2");
    if (h_perm && p > 12) DBG_SET_ERR(c, DBG_E_BADARG, "a permutation table is supported up to p = 12 (4^p entries)");
    *n_out = 0;
    if (!s->n_seqs) return DBG_OK;
    TRY(arena_begin(c));
    cudaStream_t st = c->stream;
    DBuf<u32> cnt, perm;
    DBuf<u64> off, tot;
    TRY(cnt.alloc(c, s->n_seqs)); TRY(off.alloc(c, s->n_seqs)); TRY(tot.alloc(c, 1));
    if (h_perm) {
        TRY(perm.alloc(c, 1ull << (2 * p)));
        CU(c, cudaMemcpyAsync(perm.p, h_perm, sizeof(u32) << (2 * p), cudaMemcpyHostToDevice, st));
    }
    MspSeqArgs a;
    a.words = s->words; a.n_words = s->n_words; a.start = s->start; a.length = s->length; a.uniform_len = s->uniform_len; a.n_seqs = s->n_seqs;
    a.k = k; a.p = p; a.rc = rc; a.perm = h_perm ? perm.p : nullptr;
    a.off = nullptr; a.cnt = cnt.p; a.o_seq = a.o_start = a.o_len = a.o_bucket = nullptr; a.o_exts = nullptr;
    msp_sequence_kernel<<<grid_for(s->n_seqs, 128), 128, 0, st>>>(a);
    TRY(check_launch(c, "msp_sequence_count"));
    TRY(exclusive_scan_u32_to_u64(c, cnt.p, off.p, s->n_seqs, tot.p));
    u64 total = 0;
    TRY(read_u64(c, tot.p, &total));
    *n_out = total;
    if (cap < total || !total) return DBG_OK;   // size query (or nothing to write)
    if (!h_seq || !h_start || !h_len || !h_bucket || !h_exts) DBG_SET_ERR(c, DBG_E_BADARG, "null output array");
    DBuf<u32> d_seq, d_start, d_len, d_bucket;
    DBuf<u8> d_exts;
    TRY(d_seq.alloc(c, total)); TRY(d_start.alloc(c, total)); TRY(d_len.alloc(c, total)); TRY(d_bucket.alloc(c, total)); TRY(d_exts.alloc(c, total));
    a.off = off.p; a.o_seq = d_seq.p; a.o_start = d_start.p; a.o_len = d_len.p; a.o_bucket = d_bucket.p; a.o_exts = d_exts.p;
    msp_sequence_kernel<<<grid_for(s->n_seqs, 128), 128, 0, st>>>(a);
    TRY(check_launch(c, "msp_sequence"));
    CU(c, cudaMemcpyAsync(h_seq, d_seq.p, total * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(h_start, d_start.p, total * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(h_len, d_len.p, total * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(h_bucket, d_bucket.p, total * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(h_exts, d_exts.p, total, cudaMemcpyDeviceToHost, st));
    return sync(c);
}

// ------------------------------------------------------------------------------------------------
// CountFilterSet<u8> (src/filter.rs:68-101): per k-mer the SORTED, DEDUPLICATED set of the data values (labels / colours, one
// per input sequence) of its observations; valid iff the number of observations >= min_kmer_obs.  Labels < 64, so the set
// is a 64-bit mask (bit c = label c seen) — exactly the information of the reference's Vec<u8> after sort() + dedup().
// Built from the CountFilter path: sequences are grouped by label, every group is counted on its own (CountFilter(1): all its
// distinct k-mers with Exts and per-label counts), the per-label tables are concatenated, sorted by k-mer and reduced per
// k-mer (Exts OR, counts summed, label bits OR).  A per-label count saturates at 65535 like every CountFilter count, so
// thresholds above 65535 are not supported here.
// ------------------------------------------------------------------------------------------------
__global__ void gather_subset_kernel(const u64* __restrict__ start, const u32* __restrict__ length, const u8* __restrict__ exts, u32 uniform_len,
                                     const u32* __restrict__ idx, u64 n, u64* __restrict__ o_start, u32* __restrict__ o_length, u8* __restrict__ o_exts) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 j = idx[i];
    o_start[i] = uniform_len ? (u64)j * uniform_len : start[j];
    o_length[i] = uniform_len ? uniform_len : length[j];
    if (o_exts) o_exts[i] = exts[j];
}
__global__ void pack_label_vals_kernel(const u8* __restrict__ exts, const u16* __restrict__ counts, u32 label, u64 n, u32* __restrict__ val) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) val[i] = (u32)exts[i] | ((u32)counts[i] << 8) | (label << 24);
}
// thread at the head of every run of equal keys reduces the run (<= 64 entries: one per label)
template <int W>
__global__ void colorset_reduce_kernel(const u64* __restrict__ lo, const u64* __restrict__ hi, const u32* __restrict__ val, u64 n, u32 min_obs,
                                       const u64* __restrict__ pos /* nullptr: flag pass */, u32* __restrict__ flag, u64* __restrict__ o_lo,
                                       u64* __restrict__ o_hi, u8* __restrict__ o_exts, u16* __restrict__ o_counts, u64* __restrict__ o_colors) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = i == 0 || lo[i] != lo[i - 1] || (W == 2 && hi[i] != hi[i - 1]);
    if (!head) { if (!pos) flag[i] = 0; return; }
    u32 e = 0;
    u64 nobs = 0, colors = 0;
    for (u64 j = i; j < n && lo[j] == lo[i] && (W == 1 || hi[j] == hi[i]); j++) {
        const u32 v = val[j];
        e |= v & 0xffu;
        nobs += (v >> 8) & 0xffffu;
        colors |= 1ull << (v >> 24);
    }
    const bool valid = nobs >= min_obs;
    if (!pos) { flag[i] = valid ? 1u : 0u; return; }
    if (!valid) return;
    const u64 o = pos[i];
    o_lo[o] = lo[i];
    if (W == 2) o_hi[o] = hi[i];
    o_exts[o] = (u8)e;
    o_counts[o] = (u16)(nobs > 65535 ? 65535 : nobs);
    o_colors[o] = colors;
}

template <int W>
static int colorset_merge(Ctx* c, int k, std::vector<Table*>& parts, const std::vector<u32>& labels, u32 min_obs, Table* t) {
    cudaStream_t st = c->stream;
    u64 total = 0;
    for (Table* p : parts) total += p->n;
    t->n = 0;
    if (!total) return DBG_OK;
    TRY(arena_begin(c));
    DBuf<u64> a_lo, a_hi, b_lo, b_hi, pos, tot;
    DBuf<u32> a_v, b_v, flag;
    TRY(a_lo.alloc(c, total)); TRY(b_lo.alloc(c, total)); TRY(a_v.alloc(c, total)); TRY(b_v.alloc(c, total));
    if (W == 2) { TRY(a_hi.alloc(c, total)); TRY(b_hi.alloc(c, total)); }
    u64 off = 0;
    for (size_t q = 0; q < parts.size(); q++) {
        Table* p = parts[q];
        if (!p->n) continue;
        CU(c, cudaMemcpyAsync(a_lo.p + off, p->lo, p->n * 8, cudaMemcpyDeviceToDevice, st));
        if (W == 2) CU(c, cudaMemcpyAsync(a_hi.p + off, p->hi, p->n * 8, cudaMemcpyDeviceToDevice, st));
        pack_label_vals_kernel<<<grid_for(p->n, 256), 256, 0, st>>>(p->exts, p->counts, labels[q], p->n, a_v.p + off);
        TRY(check_launch(c, "pack_label_vals"));
        off += p->n;
    }
    u64 *rlo, *rhi;
    u32* rv;
    TRY(radix_sort_pairs(c, W, 2 * k, total, a_lo.p, a_hi.p, a_v.p, b_lo.p, b_hi.p, b_v.p, &rlo, &rhi, &rv));
    TRY(flag.alloc(c, total)); TRY(pos.alloc(c, total)); TRY(tot.alloc(c, 1));
    colorset_reduce_kernel<W><<<grid_for(total, 256), 256, 0, st>>>(rlo, rhi, rv, total, min_obs, nullptr, flag.p, nullptr, nullptr, nullptr, nullptr, nullptr);
    TRY(check_launch(c, "colorset_flag"));
    TRY(exclusive_scan_u32_to_u64(c, flag.p, pos.p, total, tot.p));
    u64 V = 0;
    TRY(read_u64(c, tot.p, &V));
    t->n = V;
    if (!V) return DBG_OK;
    DBuf<u64> o_lo, o_hi, o_col;
    DBuf<u8> o_e;
    DBuf<u16> o_c;
    TRY(o_lo.alloc_pool(c, V)); TRY(o_col.alloc_pool(c, V)); TRY(o_e.alloc_pool(c, V)); TRY(o_c.alloc_pool(c, V));
    if (W == 2) TRY(o_hi.alloc_pool(c, V));
    colorset_reduce_kernel<W><<<grid_for(total, 256), 256, 0, st>>>(rlo, rhi, rv, total, min_obs, pos.p, nullptr, o_lo.p, o_hi.p, o_e.p, o_c.p, o_col.p);
    TRY(check_launch(c, "colorset_reduce"));
    TRY(sync(c));
    t->lo = o_lo.take(); if (W == 2) t->hi = o_hi.take();
    t->exts = o_e.take(); t->counts = o_c.take(); t->colors = o_col.take();
    return DBG_OK;
}

int filter_kmers_colorset_dev(Ctx* c, int k, const SeqSet* s, const u8* h_labels, u32 min_obs, int stranded, u64 mem_gb, Table** out) {
    *out = nullptr;
    if (k < 4 || k > 64) DBG_SET_ERR(c, DBG_E_BADARG, "k=%d outside [4,64]", k);
    if (!s || (s->n_seqs && !h_labels)) DBG_SET_ERR(c, DBG_E_BADARG, "null argument");
    if (min_obs > 65535) DBG_SET_ERR(c, DBG_E_BADARG, "CountFilterSet thresholds above 65535 are not supported (per-label counts saturate)");
    if (s->n_seqs >= (1ull << 32)) DBG_SET_ERR(c, DBG_E_BADARG, "too many sequences");
    std::vector<std::vector<u32>> groups(64);
    for (u64 i = 0; i < s->n_seqs; i++) {
        if (h_labels[i] >= 64) DBG_SET_ERR(c, DBG_E_BADARG, "label %u of sequence %llu: labels must be < 64", (unsigned)h_labels[i], (unsigned long long)i);
        groups[h_labels[i]].push_back((u32)i);
    }
    SeqSet* sm = const_cast<SeqSet*>(s);
    TRY(seqset_ready(c, sm));
    std::vector<Table*> parts;
    std::vector<u32> labels;
    auto cleanup = [&]() { for (Table* p : parts) free_table(p); };
    u64 n_input = 0;
    for (u32 lab = 0; lab < 64; lab++) {
        const std::vector<u32>& g = groups[lab];
        if (g.empty()) continue;
        DBuf<u32> d_idx, d_len;
        DBuf<u64> d_start;
        DBuf<u8> d_exts;
        int rc = d_idx.alloc_pool(c, g.size());
        if (rc == DBG_OK) rc = d_len.alloc_pool(c, g.size());
        if (rc == DBG_OK) rc = d_start.alloc_pool(c, g.size());
        if (rc == DBG_OK && s->seq_exts) rc = d_exts.alloc_pool(c, g.size());
        if (rc != DBG_OK) { cleanup(); return rc; }
        cudaMemcpyAsync(d_idx.p, g.data(), g.size() * 4, cudaMemcpyHostToDevice, c->stream);
        gather_subset_kernel<<<grid_for(g.size(), 256), 256, 0, c->stream>>>(s->start, s->length, s->seq_exts, s->uniform_len, d_idx.p, g.size(),
                                                                             d_start.p, d_len.p, s->seq_exts ? d_exts.p : nullptr);
        rc = check_launch(c, "gather_subset");
        if (rc == DBG_OK) rc = sync(c);   // g.data() is host memory
        if (rc != DBG_OK) { cleanup(); return rc; }
        SeqSet sub;
        memset(sub.pend_ev, 0, sizeof(sub.pend_ev));
        sub.ctx = c; sub.words = s->words; sub.n_words = s->n_words; sub.start = d_start.p; sub.length = d_len.p;
        sub.seq_exts = s->seq_exts ? d_exts.p : nullptr; sub.n_seqs = g.size(); sub.uniform_len = 0;
        sub.max_len = s->uniform_len ? s->uniform_len : s->max_len; sub.contiguous = false; sub.owned = false;
        Table* tp = nullptr;
        rc = filter_kmers_dev(c, k, &sub, 1, stranded, 0, mem_gb, &tp);   // CountFilter(1): every distinct k-mer of this label
        if (rc != DBG_OK) { cleanup(); return rc; }
        n_input += tp->n_input;
        parts.push_back(tp);
        labels.push_back(lab);
    }
    Table* t = &(new dbg_kmer_table())->t;
    t->ctx = c; t->k = k; t->n_input = n_input;
    int rc = k <= 32 ? colorset_merge<1>(c, k, parts, labels, min_obs, t) : colorset_merge<2>(c, k, parts, labels, min_obs, t);
    cleanup();
    if (rc != DBG_OK) { free_table(t); return rc; }
    c->stats.n_valid = t->n; c->stats.n_input_kmers = n_input;
    *out = t;
    return DBG_OK;
}

}  // namespace dbg
