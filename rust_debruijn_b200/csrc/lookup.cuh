// k-mer -> index lookup over a SORTED k-mer array (stands in for BoomHashMap2::get_key_id, src/filter.rs:228,
// src/compression.rs:410): prefix LUT + binary search.  Shared by compress.cu and shard_compress.cu.
#pragma once
#include "common.cuh"

namespace dbg {

static const u32 NIL = 0xffffffffu;   // no successor / finished chain
static const u32 NIL2 = 0xfffffffeu;  // finished and mirrored into the other ping-pong buffer

template <int W>
__device__ __forceinline__ Kmer<W> load_key(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 i) {
    Kmer<W> k;
    if constexpr (W == 1) { k.lo = lo[i]; }
    else { k.lo = lo[i]; k.hi = hi[i]; }
    return k;
}

// ---- S3: k-mer -> index lookup.  The table arrives sorted (filter.rs:205-219), so instead of building a hash
// table (the reference's BoomHashMap2) the lookup is a prefix LUT (first index of every top-LB-bit prefix,
// built with a histogram + exclusive scan) followed by a binary search over the ~8 keys sharing the prefix. ----
template <int W>
__device__ __forceinline__ u32 key_prefix(Kmer<W> k, int shift) {  // top bits of the 2K-bit key: key >> shift
    if constexpr (W == 1) { return (u32)(k.lo >> shift); }
    else { return shift >= 64 ? (u32)(k.hi >> (shift - 64)) : (u32)((k.hi << (64 - shift)) | (k.lo >> shift)); }
}

template <int W>
__global__ void lut_hist_kernel(const u64* __restrict__ lo, const u64* __restrict__ hi, u64 n, int shift, u32* __restrict__ cnt) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 pfx = 0xffffffffu;
    if (i < n) pfx = key_prefix<W>(load_key<W>(lo, hi, i), shift);
    // sorted input: lanes of a warp mostly share a prefix -> one atomic per distinct prefix in the warp
    u32 peers = __match_any_sync(0xffffffffu, pfx);
    if (i < n && (peers & ((1u << (threadIdx.x & 31)) - 1)) == 0) atomicAdd(&cnt[pfx], (u32)__popc(peers));
}

template <int W>
__device__ __forceinline__ u32 table_find(const u64* __restrict__ lo, const u64* __restrict__ hi, const u64* __restrict__ lut,
                                          int shift, Kmer<W> key) {
    u32 pfx = key_prefix<W>(key, shift);
    u64 a = lut[pfx], b = lut[pfx + 1];
    while (a < b) {
        u64 m = (a + b) >> 1;
        Kmer<W> km = load_key<W>(lo, hi, m);
        if (km == key) return (u32)m;
        if (km < key) a = m + 1; else b = m;
    }
    return NIL;
}

#ifndef LUT_KEYS_LOG2
#define LUT_KEYS_LOG2 4   // ~2^4 keys per prefix: the binary search ends inside one 128-byte line
#endif
template <int W>
static inline int build_prefix_lut(Ctx* c, int k, const u64* lo, const u64* hi, u64 n, DBuf<u32>& cnt, DBuf<u64>& lut, int* shift_out) {
    int lb = 8;
    while ((1ull << (lb + LUT_KEYS_LOG2)) <= n && lb < 24) lb++;
    if (lb > 2 * k) lb = 2 * k;
    const int shift = 2 * k - lb;
    const u64 n_pfx = 1ull << lb;
    TRY(cnt.alloc(c, n_pfx));
    TRY(lut.alloc(c, n_pfx + 1));
    TRY(cnt.zero());
    if (n) {
        lut_hist_kernel<W><<<grid_for(n, 256), 256, 0, c->stream>>>(lo, hi, n, shift, cnt.p);
        TRY(check_launch(c, "lut_hist"));
    }
    TRY(exclusive_scan_u32_to_u64(c, cnt.p, lut.p, n_pfx, lut.p + n_pfx));
    *shift_out = shift;
    return DBG_OK;
}


}  // namespace dbg
