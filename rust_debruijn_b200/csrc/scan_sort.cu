// Device-wide exclusive scan and stable LSD radix sort of (k-mer, payload) pairs.
//
// Role on the path: filter_kmers must hand back valid k-mers in ASCENDING order
// (src/filter.rs:205-219: buckets are key prefixes visited in order, each sorted with sort_by_key).
// The counting stage emits distinct k-mers bucket by MSP bucket (unordered across buckets), so the
// V valid records (V ~ 0.03 N on 50x noisy reads) are sorted here — the N-sized sort the reference
// does per bucket (src/filter.rs:206) never happens on the device.
#include "common.cuh"

namespace dbg {

// ------------------------------------------------------------------------------------------------
// exclusive scan: 3 kernels (tile sums, scan of tile sums, apply)
// ------------------------------------------------------------------------------------------------
static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u64 block_exclusive_scan(u64 v, u64* s_warp, u64& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u64 w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        u64 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    block_total = s_warp[32];
    u64 r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const TIn* __restrict__ in, u64* __restrict__ partial, u64 n) {
    __shared__ u64 s_warp[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += (u64)in[base + i];
    u64 tot;
    block_exclusive_scan(s, s_warp, tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_partials(u64* partial, u64 nb, u64* d_total) {
    __shared__ u64 s_warp[33];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (u64 b0 = 0; b0 < nb; b0 += 1024) {
        u64 i = b0 + threadIdx.x;
        u64 v = i < nb ? partial[i] : 0;
        u64 tot;
        u64 ex = block_exclusive_scan(v, s_warp, tot);
        u64 carry = s_carry;
        if (i < nb) partial[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && d_total) *d_total = s_carry;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const TIn* __restrict__ in, u64* __restrict__ out,
                                                           const u64* __restrict__ partial, u64 n) {
    __shared__ u64 s_warp[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u64 v[SCAN_ITEMS];
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? (u64)in[base + i] : 0;
        s += v[i];
    }
    u64 tot;
    u64 ex = block_exclusive_scan(s, s_warp, tot) + partial[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

template <typename TIn>
static int exclusive_scan_impl(Ctx* c, const TIn* in, u64* out, u64 n, u64* d_total) {
    if (n == 0) {
        if (d_total) CU(c, cudaMemsetAsync(d_total, 0, 8, c->stream));
        return DBG_OK;
    }
    u64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    DBuf<u64> partial;
    TRY(partial.alloc(c, nb));
    scan_tile_sums<TIn><<<(u32)nb, SCAN_THREADS, 0, c->stream>>>(in, partial.p, n);
    TRY(check_launch(c, "scan_tile_sums"));
    scan_partials<<<1, 1024, 0, c->stream>>>(partial.p, nb, d_total);
    TRY(check_launch(c, "scan_partials"));
    scan_apply<TIn><<<(u32)nb, SCAN_THREADS, 0, c->stream>>>(in, out, partial.p, n);
    TRY(check_launch(c, "scan_apply"));
    return DBG_OK;
}
int exclusive_scan_u32_to_u64(Ctx* c, const u32* in, u64* out, u64 n, u64* d_total) {
    return exclusive_scan_impl<u32>(c, in, out, n, d_total);
}
int exclusive_scan_u64(Ctx* c, const u64* in, u64* out, u64 n, u64* d_total) {
    return exclusive_scan_impl<u64>(c, in, out, n, d_total);
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits.  Per pass: upsweep (per-tile digit counts, digit-major), one exclusive
// scan over the 256 x ntiles count matrix (gives every (digit, tile) its global base directly), and a
// downsweep that ranks with warp match_any (stable) and scatters.
// ------------------------------------------------------------------------------------------------
static const int RS_THREADS = 256;
static const int RS_WARPS = RS_THREADS / 32;
static const int RS_ITEMS = 8;                       // rounds per warp
static const int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per CTA
static const int RS_WARP_SEG = 32 * RS_ITEMS;        // contiguous keys per warp

__device__ __forceinline__ u32 digit_of(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 i, int d) {
    return d < 8 ? (u32)(lo[i] >> (8 * d)) & 0xffu : (u32)(hi[i] >> (8 * (d - 8))) & 0xffu;
}

__global__ void __launch_bounds__(RS_THREADS) rs_upsweep(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 n,
                                                         u32 ntiles, int d, u32* __restrict__ counts) {
    __shared__ u32 hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 wbase = (u64)blockIdx.x * RS_TILE + (u64)warp * RS_WARP_SEG;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        u64 i = wbase + r * 32 + lane;
        u32 dg = i < n ? digit_of(lo, hi, i, d) : 0xffffffffu;
        u32 peers = __match_any_sync(0xffffffffu, dg);
        if (dg != 0xffffffffu && (peers & ((1u << lane) - 1)) == 0) atomicAdd(&hist[dg], __popc(peers));
    }
    __syncthreads();
    counts[(u64)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

template <int W>
__global__ void __launch_bounds__(RS_THREADS) rs_downsweep(const u64* __restrict__ lo, const u64* __restrict__ hi,
                                                           const u32* __restrict__ val, u64* __restrict__ olo,
                                                           u64* __restrict__ ohi, u32* __restrict__ oval, u64 n,
                                                           u32 ntiles, int d, const u64* __restrict__ base) {
    __shared__ u64 woff[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&woff[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1;
    u64 wbase = (u64)blockIdx.x * RS_TILE + (u64)warp * RS_WARP_SEG;
    u64 klo[RS_ITEMS], khi[RS_ITEMS];
    u32 kv[RS_ITEMS], rank[RS_ITEMS], dgt[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        u64 i = wbase + r * 32 + lane;
        bool act = i < n;
        klo[r] = act ? lo[i] : 0;
        khi[r] = (W == 2 && act) ? hi[i] : 0;
        kv[r] = act ? val[i] : 0;
        u32 dg = act ? (d < 8 ? (u32)(klo[r] >> (8 * d)) & 0xffu : (u32)(khi[r] >> (8 * (d - 8))) & 0xffu) : 0xffffffffu;
        dgt[r] = dg;
        u32 peers = __match_any_sync(0xffffffffu, dg);
        u32 before = act ? (u32)woff[warp][dg] : 0;
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
        if (act && (peers & lt) == 0) woff[warp][dg] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // digit = threadIdx.x: exclusive prefix over the warps of this tile + global base
        u64 run = base[(u64)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            u64 cnt = woff[w][threadIdx.x];
            woff[w][threadIdx.x] = run;
            run += cnt;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        if (dgt[r] != 0xffffffffu) {
            u64 pos = woff[warp][dgt[r]] + rank[r];
            olo[pos] = klo[r];
            if (W == 2) ohi[pos] = khi[r];
            oval[pos] = kv[r];
        }
    }
}

int radix_sort_pairs(Ctx* c, int W, int key_bits, u64 n, u64* lo_a, u64* hi_a, u32* val_a, u64* lo_b, u64* hi_b,
                     u32* val_b, u64** res_lo, u64** res_hi, u32** res_val) {
    *res_lo = lo_a; *res_hi = hi_a; *res_val = val_a;
    if (n <= 1) return DBG_OK;
    int passes = (key_bits + 7) / 8;
    u32 ntiles = (u32)((n + RS_TILE - 1) / RS_TILE);
    DBuf<u32> counts;
    DBuf<u64> base;
    TRY(counts.alloc(c, (u64)256 * ntiles));
    TRY(base.alloc(c, (u64)256 * ntiles));
    u64 *slo = lo_a, *shi = hi_a, *dlo = lo_b, *dhi = hi_b;
    u32 *sv = val_a, *dv = val_b;
    for (int d = 0; d < passes; d++) {
        rs_upsweep<<<ntiles, RS_THREADS, 0, c->stream>>>(slo, shi, n, ntiles, d, counts.p);
        TRY(check_launch(c, "rs_upsweep"));
        TRY(exclusive_scan_u32_to_u64(c, counts.p, base.p, (u64)256 * ntiles, nullptr));
        if (W == 1)
            rs_downsweep<1><<<ntiles, RS_THREADS, 0, c->stream>>>(slo, shi, sv, dlo, dhi, dv, n, ntiles, d, base.p);
        else
            rs_downsweep<2><<<ntiles, RS_THREADS, 0, c->stream>>>(slo, shi, sv, dlo, dhi, dv, n, ntiles, d, base.p);
        TRY(check_launch(c, "rs_downsweep"));
        std::swap(slo, dlo);
        std::swap(shi, dhi);
        std::swap(sv, dv);
    }
    *res_lo = slo; *res_hi = shi; *res_val = sv;
    return DBG_OK;
}

}  // namespace dbg
