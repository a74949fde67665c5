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
// LSD radix sort, 8-bit digits, "onesweep" form: one kernel computes the digit histograms of every
// pass (they do not depend on the order), then each pass is a single kernel: tiles are taken in
// order from an atomic counter, ranked with warp match_any (stable), their per-digit counts chained
// through a decoupled look-back, and the tile is scattered through shared memory so that the global
// stores of each digit run are contiguous.  2 x 12 bytes per key per pass instead of 3 reads + 1 write.
// ------------------------------------------------------------------------------------------------
static const int OS_THREADS = 512;
static const int OS_WARPS = OS_THREADS / 32;
static const int OS_ITEMS = 8;
static const int OS_TILE = OS_THREADS * OS_ITEMS;  // 4096 keys per CTA
static const int OS_SEG = 32 * OS_ITEMS;           // contiguous keys per warp
static const u64 OS_FLAG_AGG = 1ull << 62, OS_FLAG_PREFIX = 2ull << 62, OS_VAL_MASK = (1ull << 62) - 1;

__global__ void __launch_bounds__(512) rs_hist_kernel(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 n, int d0, int passes,
                                                      u64* __restrict__ ghist) {
    __shared__ u32 h[16 * 256];
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 l = lo[i], hh = passes > 8 ? hi[i] : 0;
        for (int d = d0; d < passes; d++) {
            u32 dg = d < 8 ? (u32)(l >> (8 * d)) & 0xffu : (u32)(hh >> (8 * (d - 8))) & 0xffu;
            atomicAdd(&h[d * 256 + dg], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x)
        if (h[i]) atomicAdd(&ghist[i], (u64)h[i]);
}

__global__ void rs_hist_scan_kernel(u64* ghist, int passes) {  // one warp per pass: exclusive scan of 256 counters
    int d = blockIdx.x, lane = threadIdx.x;
    if (d >= passes) return;
    u64 carry = 0;
    for (int c = 0; c < 8; c++) {
        u64 v = ghist[d * 256 + c * 32 + lane], inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        ghist[d * 256 + c * 32 + lane] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// resident CTAs per SM: 3 for one-word keys (40 registers, a few spilled words; 67 KB of shared memory each), 2 for two-word keys
// (measured: 2.39 -> 2.11 ms at K=31 with 3; K=63: 2.9 ms with 2, 3.1 ms with 3)
template <int W>
__global__ void __launch_bounds__(OS_THREADS, W == 1 ? 3 : 2) rs_onesweep(const u64* __restrict__ lo, const u64* __restrict__ hi,
                                                          const u32* __restrict__ val, u64* __restrict__ olo,
                                                          u64* __restrict__ ohi, u32* __restrict__ oval, u64 n, int d,
                                                          const u64* __restrict__ gbase, u64* status, u32* tile_counter) {
    extern __shared__ __align__(16) unsigned char os_smem[];
    u64* s_lo = reinterpret_cast<u64*>(os_smem);
    u64* s_hi = s_lo + (W == 2 ? OS_TILE : 0);
    u32* s_val = reinterpret_cast<u32*>(s_hi + OS_TILE);
    __shared__ u32 whist[OS_WARPS][256];
    __shared__ u32 dstart[256];
    __shared__ u64 gpos[256];
    __shared__ u32 s_wsum[8];
    __shared__ u32 s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < OS_WARPS * 256; i += OS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1;
    const u64 tbase = (u64)tile * OS_TILE, wbase = tbase + (u64)warp * OS_SEG;
    u64 klo[OS_ITEMS], khi[OS_ITEMS];
    u32 kv[OS_ITEMS], rank[OS_ITEMS], dgt[OS_ITEMS];
    // all loads first: the ranking rounds below are separated by __syncwarp (a memory fence the loads cannot be hoisted over), so
    // loading inside them paid one global-memory latency per item instead of one per thread
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        const u64 i = wbase + r * 32 + lane;
        const bool act = i < n;
        klo[r] = act ? lo[i] : 0;
        khi[r] = (W == 2 && act) ? hi[i] : 0;
        kv[r] = act ? val[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        u64 i = wbase + r * 32 + lane;
        bool act = i < n;
        u32 dg = act ? (d < 8 ? (u32)(klo[r] >> (8 * d)) & 0xffu : (u32)(khi[r] >> (8 * (d - 8))) & 0xffu) : 0xffffffffu;
        dgt[r] = dg;
        u32 peers = __match_any_sync(0xffffffffu, dg);
        u32 before = act ? whist[warp][dg] : 0;
        rank[r] = before + __popc(peers & lt);
        __syncwarp();
        if (act && (peers & lt) == 0) whist[warp][dg] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // digit totals of the tile (thread = digit), per-warp exclusive offsets, tile-local digit starts
    u32 cnt = 0;
    if (threadIdx.x < 256) {
#pragma unroll
        for (int w = 0; w < OS_WARPS; w++) { u32 c = whist[w][threadIdx.x]; whist[w][threadIdx.x] = cnt; cnt += c; }
        // publish the aggregate right away so later tiles can move on
        u64 st = (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG) | cnt;
        *reinterpret_cast<volatile u64*>(&status[(u64)tile * 256 + threadIdx.x]) = st;
        u32 inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_wsum[warp] = inc;
        dstart[threadIdx.x] = inc - cnt;  // warp-local exclusive; fixed up after the barrier
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        u32 add = 0;
        for (int w = 0; w < warp; w++) add += s_wsum[w];
        u32 ds = dstart[threadIdx.x] + add;
        // decoupled look-back over the preceding tiles for this digit
        u64 excl = 0;
        if (tile > 0) {
            for (long long tt = (long long)tile - 1; tt >= 0; tt--) {
                u64 v;
                do { v = *reinterpret_cast<volatile u64*>(&status[(u64)tt * 256 + threadIdx.x]); } while ((v >> 62) == 0);
                excl += v & OS_VAL_MASK;
                if ((v >> 62) == 2) break;
            }
            *reinterpret_cast<volatile u64*>(&status[(u64)tile * 256 + threadIdx.x]) = OS_FLAG_PREFIX | (excl + cnt);
        }
        __syncwarp();
        dstart[threadIdx.x] = ds;
        gpos[threadIdx.x] = gbase[threadIdx.x] + excl - ds;
    }
    __syncthreads();
    // stage the tile in digit order
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        if (dgt[r] != 0xffffffffu) {
            u32 li = dstart[dgt[r]] + whist[warp][dgt[r]] + rank[r];
            s_lo[li] = klo[r];
            if (W == 2) s_hi[li] = khi[r];
            s_val[li] = kv[r];
        }
    }
    __syncthreads();
    const u32 tcount = (u32)min((u64)OS_TILE, n - tbase);
    for (u32 i = threadIdx.x; i < tcount; i += OS_THREADS) {
        u64 l = s_lo[i], h = W == 2 ? s_hi[i] : 0;
        u32 dg = d < 8 ? (u32)(l >> (8 * d)) & 0xffu : (u32)(h >> (8 * (d - 8))) & 0xffu;
        u64 pos = gpos[dg] + i;
        olo[pos] = l;
        if (W == 2) ohi[pos] = h;
        oval[pos] = s_val[i];
    }
}

// After sorting on the top digits only: keys that share the sorted prefix form (rare, short) runs that are still in
// their original relative order.  The thread at a run start insertion-sorts the run on the full key.  Runs longer
// than RS_FIX_MAX raise a flag (the caller then sorts on every digit instead).
static const int RS_FIX_MAX = 48;
template <int W>
__global__ void rs_fixup_kernel(u64* __restrict__ lo, u64* __restrict__ hi, u32* __restrict__ val, u64 n, int shift, u32* flag) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto same = [&](u64 x, u64 y) -> bool {   // do keys x and y share the sorted prefix (key >> shift)?
        if (W == 1) return (lo[x] >> shift) == (lo[y] >> shift);
        if (shift >= 64) return (hi[x] >> (shift - 64)) == (hi[y] >> (shift - 64));
        return hi[x] == hi[y] && (lo[x] >> shift) == (lo[y] >> shift);
    };
    if (i > 0 && same(i - 1, i)) return;           // not a run start
    if (i + 1 >= n || !same(i, i + 1)) return;     // run of one
    u64 j = i + 1;
    while (j < n && same(i, j) && j - i <= (u64)RS_FIX_MAX) j++;
    if (j - i > (u64)RS_FIX_MAX) { atomicExch(flag, 1u); return; }
    for (u64 a = i + 1; a < j; a++) {
        u64 kl = lo[a], kh = W == 2 ? hi[a] : 0;
        u32 kv = val[a];
        u64 b = a;
        while (b > i) {
            u64 pl = lo[b - 1], ph = W == 2 ? hi[b - 1] : 0;
            bool greater = W == 2 ? (ph > kh || (ph == kh && pl > kl)) : pl > kl;
            if (!greater) break;
            lo[b] = pl;
            if (W == 2) hi[b] = ph;
            val[b] = val[b - 1];
            b--;
        }
        lo[b] = kl;
        if (W == 2) hi[b] = kh;
        val[b] = kv;
    }
}

int radix_sort_pairs(Ctx* c, int W, int key_bits, u64 n, u64* lo_a, u64* hi_a, u32* val_a, u64* lo_b, u64* hi_b,
                     u32* val_b, u64** res_lo, u64** res_hi, u32** res_val) {
    *res_lo = lo_a; *res_hi = hi_a; *res_val = val_a;
    if (n <= 1) return DBG_OK;
    const int passes = (key_bits + 7) / 8;
    // digits to sort on: enough top bits that equal prefixes are rare (>= log2(n) + 4 effective bits)
    int need_bits = 4;
    while ((1ull << (need_bits - 4)) < n && need_bits < 128) need_bits++;
    int npass = passes;
    for (int q = 1; q <= passes; q++) {
        int eff = 8 * q - (8 * passes - key_bits);
        if (eff >= need_bits) { npass = q; break; }
    }
    u32 ntiles = (u32)((n + OS_TILE - 1) / OS_TILE);
    DBuf<u64> ghist, status;
    DBuf<u32> tctr;
    TRY(ghist.alloc(c, 16 * 256));
    TRY(status.alloc(c, (u64)ntiles * 256));
    TRY(tctr.alloc(c, 32));
    size_t smem = (size_t)OS_TILE * (W == 2 ? 20 : 12);
    if (W == 1) CU(c, cudaFuncSetAttribute(rs_onesweep<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(c, cudaFuncSetAttribute(rs_onesweep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    u64 *slo = lo_a, *shi = hi_a, *dlo = lo_b, *dhi = hi_b;
    u32 *sv = val_a, *dv = val_b;
    for (int attempt = 0; attempt < 2; attempt++) {
        const int d0 = passes - npass;
        TRY(ghist.zero());
        TRY(tctr.zero());
        u32 hgrid = (u32)std::min<u64>((n + 511) / 512, (u64)c->sm_count * 4);
        rs_hist_kernel<<<hgrid, 512, 0, c->stream>>>(slo, shi, n, d0, passes, ghist.p);
        TRY(check_launch(c, "rs_hist"));
        rs_hist_scan_kernel<<<passes, 32, 0, c->stream>>>(ghist.p, passes);
        TRY(check_launch(c, "rs_hist_scan"));
        for (int d = d0; d < passes; d++) {
            TRY(status.zero());
            if (W == 1)
                rs_onesweep<1><<<ntiles, OS_THREADS, smem, c->stream>>>(slo, shi, sv, dlo, dhi, dv, n, d, ghist.p + d * 256, status.p, tctr.p + d);
            else
                rs_onesweep<2><<<ntiles, OS_THREADS, smem, c->stream>>>(slo, shi, sv, dlo, dhi, dv, n, d, ghist.p + d * 256, status.p, tctr.p + d);
            TRY(check_launch(c, "rs_onesweep"));
            std::swap(slo, dlo);
            std::swap(shi, dhi);
            std::swap(sv, dv);
        }
        if (d0 == 0) break;  // sorted on every digit
        // fix the runs that share the sorted prefix
        CU(c, cudaMemsetAsync(tctr.p + 31, 0, 4, c->stream));
        const int shift = 8 * d0;  // prefix = key >> shift
        if (W == 1) rs_fixup_kernel<1><<<grid_for(n, 256), 256, 0, c->stream>>>(slo, shi, sv, n, shift, tctr.p + 31);
        else rs_fixup_kernel<2><<<grid_for(n, 256), 256, 0, c->stream>>>(slo, shi, sv, n, shift, tctr.p + 31);
        TRY(check_launch(c, "rs_fixup"));
        u32 flag = 0;
        CU(c, cudaMemcpyAsync(c->h_scratch, tctr.p + 31, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(c, spin_sync(c->stream));
        flag = *(u32*)c->h_scratch;
        if (!flag) break;
        npass = passes;  // long equal-prefix runs (low-complexity data): sort on every digit (LSD passes are stable, so
                         // re-sorting the partially sorted data on all digits is still correct)
    }
    *res_lo = slo; *res_hi = shi; *res_val = sv;
    return DBG_OK;
}

}  // namespace dbg
