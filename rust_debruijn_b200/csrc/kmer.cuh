// 2-bit k-mer arithmetic, Exts and super-k-mer record layout shared by every kernel (host + device).
//
// Restates (never copies) the arithmetic of the reference crate:
//   k-mer integer layout, base i at bits 2(K-1-i)            src/kmer.rs:429-437,515-518
//   extend_right / extend_left                                src/kmer.rs:469-487
//   rc = ~reverse_by_twos(x) >> 2(32W-K)                      src/kmer.rs:104-165,620-634
//   min_rc_flip (equality -> flipped branch)                  src/lib.rs:224-231
//   is_palindrome (even K only)                               src/lib.rs:244-246
//   Exts bit layout / rc / complement                         src/lib.rs:569-601,729-748
// B200 notes: reverse_by_twos is done with the BREV instruction (bit reversal) plus one
// swap-adjacent-bits step instead of the reference's 5/6 mask-shift stages; the canonical k-mer in
// the extraction loops is kept rolling (fwd and rc both updated per base) so no per-k-mer reversal
// is ever needed there.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define HD __host__ __device__ __forceinline__

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

namespace dbg {

// ---- reverse the order of 2-bit units ---------------------------------------------------------
HD u64 rev2_64(u64 x) {
#ifdef __CUDA_ARCH__
    x = __brevll(x);
#else
    x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
    x = ((x & 0x3333333333333333ull) << 2) | ((x >> 2) & 0x3333333333333333ull);
    x = ((x & 0x0F0F0F0F0F0F0F0Full) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0Full);
    x = ((x & 0x00FF00FF00FF00FFull) << 8) | ((x >> 8) & 0x00FF00FF00FF00FFull);
    x = ((x & 0x0000FFFF0000FFFFull) << 16) | ((x >> 16) & 0x0000FFFF0000FFFFull);
    x = (x << 32) | (x >> 32);
#endif
    // full bit reversal also swapped the two bits inside every base: swap them back
    return ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
}
HD u32 rev2_32(u32 x) {
#ifdef __CUDA_ARCH__
    x = __brev(x);
#else
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    x = ((x & 0x33333333u) << 2) | ((x >> 2) & 0x33333333u);
    x = ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
    x = ((x & 0x00FF00FFu) << 8) | ((x >> 8) & 0x00FF00FFu);
    x = (x << 16) | (x >> 16);
#endif
    return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}

// ---- k-mer of W 64-bit words (W=1: K<=32, W=2: K<=64), right-aligned, unused top bits zero -----
template <int W>
struct Kmer;

template <>
struct Kmer<1> {
    u64 lo;
    HD bool operator==(const Kmer& o) const { return lo == o.lo; }
    HD bool operator<(const Kmer& o) const { return lo < o.lo; }
};
template <>
struct __align__(16) Kmer<2> {
    u64 lo, hi;
    HD bool operator==(const Kmer& o) const { return lo == o.lo && hi == o.hi; }
    HD bool operator<(const Kmer& o) const { return hi < o.hi || (hi == o.hi && lo < o.lo); }
};

// Per-K constants (computed on the host once, passed by value to kernels).
struct KP {
    int k;
    int top_shift;  // bit position of base 0 inside its word: 2(k-1) mod 64
    int rc_shift;   // 64W - 2k
    u64 mask_lo, mask_hi;
};
inline KP make_kp(int k) {
    KP p;
    p.k = k;
    if (k <= 32) {
        p.top_shift = 2 * (k - 1);
        p.rc_shift = 64 - 2 * k;
        p.mask_lo = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
        p.mask_hi = 0;
    } else {
        p.top_shift = 2 * (k - 1) - 64;
        p.rc_shift = 128 - 2 * k;
        p.mask_lo = ~0ull;
        p.mask_hi = (k == 64) ? ~0ull : ((1ull << (2 * k - 64)) - 1);
    }
    return p;
}

template <int W>
struct Ops;

template <>
struct Ops<1> {
    typedef Kmer<1> K;
    static HD K zero() { return K{0}; }
    static HD K ext_right(const KP& p, K x, u32 v) { return K{((x.lo << 2) & p.mask_lo) | (u64)v}; }
    static HD K ext_left(const KP& p, K x, u32 v) { return K{(x.lo >> 2) | ((u64)v << p.top_shift)}; }
    static HD K rc(const KP& p, K x) { return K{(~rev2_64(x.lo)) >> p.rc_shift}; }
    // rolling reverse complement: rc(x.ext_right(v)) == roll_rc(rc(x), v)
    static HD K roll_rc(const KP& p, K r, u32 v) { return K{(r.lo >> 2) | ((u64)(3u - v) << p.top_shift)}; }
    static HD u32 first_base(const KP& p, K x) { return (u32)(x.lo >> p.top_shift) & 3u; }
    static HD u32 last_base(const KP&, K x) { return (u32)x.lo & 3u; }
    static HD u32 hash32(K x) {  // cheap 32-bit mix for the shared-memory table (slot = high bits, class = low bits)
        u32 h = (u32)x.lo * 0x9E3779B1u + (u32)(x.lo >> 32) * 0x85EBCA6Bu;
        h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
        return h;
    }
    static HD u64 mix(K x) {  // bijective 64-bit finaliser (slot selection in the HBM table)
        u64 h = x.lo;
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        return h;
    }
};

template <>
struct Ops<2> {
    typedef Kmer<2> K;
    static HD K zero() { return K{0, 0}; }
    static HD K ext_right(const KP& p, K x, u32 v) {
        return K{(x.lo << 2) | (u64)v, ((x.hi << 2) | (x.lo >> 62)) & p.mask_hi};
    }
    static HD K ext_left(const KP& p, K x, u32 v) {
        return K{(x.lo >> 2) | (x.hi << 62), (x.hi >> 2) | ((u64)v << p.top_shift)};
    }
    static HD K rc(const KP& p, K x) {
        u64 nh = ~rev2_64(x.lo), nl = ~rev2_64(x.hi);  // 128-bit reverse swaps the words
        int s = p.rc_shift;                             // 0..62
        if (s == 0) return K{nl, nh};
        return K{(nl >> s) | (nh << (64 - s)), nh >> s};
    }
    static HD K roll_rc(const KP& p, K r, u32 v) {
        return K{(r.lo >> 2) | (r.hi << 62), (r.hi >> 2) | ((u64)(3u - v) << p.top_shift)};
    }
    static HD u32 first_base(const KP& p, K x) { return (u32)(x.hi >> p.top_shift) & 3u; }
    static HD u32 last_base(const KP&, K x) { return (u32)x.lo & 3u; }
    static HD u32 hash32(K x) {
        u32 h = (u32)x.lo * 0x9E3779B1u + (u32)(x.lo >> 32) * 0x85EBCA6Bu + (u32)x.hi * 0xC2B2AE35u + (u32)(x.hi >> 32) * 0x27D4EB2Fu;
        h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
        return h;
    }
    static HD u64 mix(K x) {
        u64 h = x.lo ^ (x.hi * 0x9E3779B97F4A7C15ull);
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        h += x.hi;  // (lo,hi) -> h stays injective in lo for fixed hi; good enough for slot/class bits
        h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
        return h;
    }
};

template <int W>
HD bool is_palindrome(const KP& p, Kmer<W> x) {
    return (p.k % 2 == 0) && (x == Ops<W>::rc(p, x));
}

// ---- Exts (src/lib.rs:577-749): bits 0..3 = left A,C,G,T ; bits 4..7 = right A,C,G,T ------------
HD u32 exts_complement(u32 v) {
    u32 r = ((v & 0x55u) << 1) | ((v >> 1) & 0x55u);
    return ((r & 0x33u) << 2) | ((r >> 2) & 0x33u);
}
HD u32 exts_rc(u32 v) { return exts_complement(((v & 0xfu) << 4) | (v >> 4)); }
HD u32 exts_side(u32 v, int dir) { return dir ? (v >> 4) & 0xfu : v & 0xfu; }  // dir: 0 = Left, 1 = Right
HD int popc4(u32 nib) { return (int)((0x4332322132212110ull >> (4 * nib)) & 7); }
HD int unique_base(u32 nib) { return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : 3; }

// ---- minimum-substring partitioning (src/msp.rs): score of a p-mer (p <= 16, right-aligned in 32 bits) and MSP bucket
// of a k-mer.  score = a hash permutation of the CANONICAL p-mer when unstranded (min(perm[p], perm[rc p]), msp.rs:305-311
// with a bijective perm), of the p-mer itself when stranded; bucket(k-mer) = low bits of the smallest score among its
// K - p + 1 p-mers: a pure, strand-symmetric function of the k-mer, so every occurrence of a canonical k-mer lands in
// the same bucket whatever read, strand or rank it was seen on. ----
HD u32 pmer_score(u32 x, int p, bool stranded) {
    if (!stranded) {
        u32 r = (~rev2_32(x)) >> (32 - 2 * p);
        x = x < r ? x : r;
    }
    x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13;
    return x;
}
// Smallest p-mer score of the k-mer WITHOUT its first p-mer (drop_first) and WITHOUT its last p-mer (drop_last): the two
// neighbours of a k-mer share all but one p-mer with it, so their buckets cost one more score each instead of K - p + 1.
template <int W>
HD void kmer_min_scores_shared(const KP& kp, Kmer<W> x, int p, bool stranded, u32& drop_first, u32& drop_last) {
    const u32 pmask = p == 16 ? 0xffffffffu : ((1u << (2 * p)) - 1);
    u32 a = 0xffffffffu, b = 0xffffffffu;   // a: p-mers 0 .. w-2, b: p-mers 1 .. w-1
    const int w1 = kp.k - p;                 // index of the last p-mer
    for (int t = w1; t >= 0; t--) {          // from the last p-mer towards the first, shifting right
        const u32 s = pmer_score((u32)x.lo & pmask, p, stranded);
        if (t != w1) a = s < a ? s : a;
        if (t != 0) b = s < b ? s : b;
        if constexpr (W == 1) { x.lo >>= 2; }
        else { x.lo = (x.lo >> 2) | (x.hi << 62); x.hi >>= 2; }
    }
    drop_first = b;
    drop_last = a;
}
template <int W>
HD u32 kmer_first_pmer(const KP& kp, Kmer<W> x, int p) {   // the p-mer at bases 0 .. p-1
    const u32 pmask = p == 16 ? 0xffffffffu : ((1u << (2 * p)) - 1);
    const int sh = 2 * (kp.k - p);
    if constexpr (W == 1) { return (u32)(x.lo >> sh) & pmask; }
    else { return (u32)(sh >= 64 ? x.hi >> (sh - 64) : (x.hi << (64 - sh)) | (x.lo >> sh)) & pmask; }
}
template <int W>
HD u32 kmer_min_score(const KP& kp, Kmer<W> x, int p, bool stranded) {
    const u32 pmask = p == 16 ? 0xffffffffu : ((1u << (2 * p)) - 1);
    u32 best = 0xffffffffu;
    // p-mer t (t = 0 .. K-p) = bases t .. t+p-1 = key >> 2(K-p-t); walk from the last p-mer towards the first by shifting right
    for (int t = kp.k - p; t >= 0; t--) {
        const u32 s = pmer_score((u32)x.lo & pmask, p, stranded);
        best = s < best ? s : best;
        if constexpr (W == 1) { x.lo >>= 2; }
        else { x.lo = (x.lo >> 2) | (x.hi << 62); x.hi >>= 2; }
    }
    return best;
}

// ---- DnaString word layout (src/dna_string.rs:383-399): base b of the concatenation sits in word
// b/32 at bits 62-2(b%32) --------------------------------------------------------------------------
// 64 bits starting at base offset `b` (bases b..b+31 left-aligned); words[] must be readable at
// index (b>>5)+1 whenever (b&31) != 0.
HD u64 bases64(const u64* words, u64 b) {
    u64 wi = b >> 5;
    int sh = (int)(b & 31) * 2;
    u64 hi = words[wi];
    if (sh == 0) return hi;
    return (hi << sh) | (words[wi + 1] >> (64 - sh));
}
HD u32 base_at(const u64* words, u64 b) { return (u32)(words[b >> 5] >> (62 - 2 * (b & 31))) & 3u; }

// ---- super-k-mer record ---------------------------------------------------------------------------
// One record = a run of n consecutive k-mers of one sequence that share an MSP bucket, stored as its
// n+K-1 bases (DnaString bit order, left-aligned from word 0) plus the 4-bit left/right extension
// masks of the run (flanking read base, or the sequence-level Exts nibble at a sequence end —
// KmerExtsIter, src/lib.rs:820-830).  W=1: 16 bytes (<= 57 bases), W=2: 32 bytes (<= 121 bases).
// Low 14 bits of the last word: [13:8] n (1..63), [7:4] right mask, [3:0] left mask.
template <int W>
struct RecLayout;
template <>
struct RecLayout<1> {
    static const int WORDS = 2;
    static const int MAX_BASES = 57;
};
template <>
struct RecLayout<2> {
    static const int WORDS = 4;
    static const int MAX_BASES = 121;
};
HD int rec_max_kmers(int rec_words, int k) {
    int mb = rec_words == 2 ? 57 : 121;
    int m = mb - k + 1;
    return m > 63 ? 63 : m;
}

}  // namespace dbg
