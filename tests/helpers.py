"""Shared helpers for the parity tests: random contig generators in the style of the reference's
src/test.rs:19-132 (seeded here, un-seeded there), ASCII<->2-bit, canonical graph form."""
import numpy as np

A2B = {"A": 0, "C": 1, "G": 2, "T": 3}


def enc(s):
    return np.array([A2B[c] for c in s], dtype=np.uint8)


def dec(b):
    return "".join("ACGT"[int(x)] for x in b)


def random_dna(rng, n):  # test.rs:19-25
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def rc_bases(b):
    return (3 - b[::-1]).astype(np.uint8)


def simple_random_contigs(rng):  # test.rs:58-95
    p1, p2, pc, p3, p4 = (random_dna(rng, n) for n in (40, 30, 100, 30, 40))
    c1 = np.concatenate([p1, pc, p3])
    c2 = np.concatenate([p2, pc, p4])
    pal1 = random_dna(rng, 33)
    c3 = np.concatenate([random_dna(rng, 30), pal1, rc_bases(pal1), random_dna(rng, 50)])
    return [c1, c2, c3]


def random_contigs(rng):  # test.rs:98-132
    nchunks = max(5, int(rng.gamma(0.6, 25.0)))
    chunks = [random_dna(rng, max(10, int(rng.gamma(1.5, 200.0)))) for _ in range(nchunks)]
    nchrom = max(4, int(rng.gamma(0.6, 25.0)))
    chroms = []
    for _ in range(nchrom):
        cc = max(4, int(rng.gamma(0.6, 25.0)))
        chroms.append(np.concatenate([chunks[int(rng.integers(0, nchunks))] for _ in range(cc)]))
    return chroms


def small_k_contigs(rng, n_ctg=6, lo=8, hi=60, alphabet=4):
    """Short random contigs over a reduced alphabet: at K=4..7 these are dense in cycles, hairpins,
    palindromes and branch points (the cases SURVEY.md §7 'hard parts' lists)."""
    return [rng.integers(0, alphabet, size=int(rng.integers(lo, hi)), dtype=np.uint8) for _ in range(n_ctg)]


def kmers_of(bases, k):
    """python ints, base 0 most significant (src/kmer.rs:429-437)."""
    out = []
    if len(bases) < k:
        return out
    x = 0
    mask = (1 << (2 * k)) - 1
    for i, b in enumerate(bases):
        x = ((x << 2) | int(b)) & mask
        if i >= k - 1:
            out.append(x)
    return out


def rc_int(x, k):
    r = 0
    for _ in range(k):
        r = (r << 2) | (3 - (x & 3))
        x >>= 2
    return r


def canon(x, k):
    return min(x, rc_int(x, k))


def node_bases(orc, g, i):
    return orc.unpack_bases(g["words"], int(g["start"][i]), int(g["length"][i]))


def assert_tables_equal(a, b):
    for f in ("lo", "hi", "exts", "counts", "all_lo", "all_hi"):
        assert np.array_equal(a[f], b[f]), f"table field {f} differs"


def assert_graphs_equal(a, b):
    assert a["n_nodes"] == b["n_nodes"] and a["n_bases"] == b["n_bases"], (a["n_nodes"], b["n_nodes"], a["n_bases"], b["n_bases"])
    for f in ("words", "start", "length", "exts", "data"):
        assert np.array_equal(a[f], b[f]), f"graph field {f} differs"


def brute_filter_colorset(orc, k, seqs, labels, min_obs, stranded=False, seq_exts=None):
    """filter_kmers with CountFilterSet<u8> (src/filter.rs:68-101, 138-231) by brute force over python ints: returns ascending
    lists (kmers, exts, label sets) of the valid k-mers.  KmerExtsIter (lib.rs:812-841), min_rc_flip (:224-231), Exts::rc (:746)."""
    L = orc.lib()
    acc = {}
    mask = (1 << (2 * k)) - 1
    for si, (sq, lab) in enumerate(zip(seqs, labels)):
        n = len(sq)
        sx = int(seq_exts[si]) if seq_exts is not None else 0
        x = 0
        for i, b in enumerate(sq):
            x = ((x << 2) | int(b)) & mask
            if i < k - 1:
                continue
            j = i - k + 1
            left = (sx & 0xf) if j == 0 else 1 << int(sq[j - 1])
            right = ((sx >> 4) & 0xf) if j == n - k else 1 << int(sq[j + k])
            e = left | (right << 4)
            r = rc_int(x, k)
            key = x
            if not stranded and not x < r:
                key, e = r, L.orc_exts_rc(e)
            a = acc.setdefault(key, [0, 0, set()])
            a[0] += 1
            a[1] |= e
            a[2].add(int(lab))
    keys = sorted(kk for kk, a in acc.items() if a[0] >= min_obs)
    return keys, [acc[kk][1] for kk in keys], [sorted(acc[kk][2]) for kk in keys], [min(acc[kk][0], 65535) for kk in keys]


def canon_form(orc, g, with_data=True):
    """Order- and strand-free form of a BaseGraph: sorted (min(seq, rc seq), Exts [rc'd with it], data)."""
    out = []
    L = orc.lib()
    for i in range(int(g["n_nodes"])):
        b = node_bases(orc, g, i)
        r = (3 - b[::-1]).astype(np.uint8)
        e = int(g["exts"][i])
        if bytes(r) < bytes(b):
            b, e = r, L.orc_exts_rc(e)
        out.append((bytes(b), e, int(g["data"][i]) if with_data else 0))
    return sorted(out)


def msp_shard_graphs(orc, k, p, contigs, stranded=False, twice=True, reduce_op=0):
    """The reference's shard_asm loop (src/test.rs:433-470): msp_sequence(rc = true) -> per-shard filter_kmers(CountFilter(2)), every
    substring pushed twice -> per-shard compress_kmers_with_hash.  Returns the shard BaseGraphs in ascending bucket order."""
    shards = {}
    for c in contigs:
        iv = orc.msp_scan(k, p, c, rc=True)
        for st, ln, b, e in zip(iv["start"], iv["len"], iv["bucket"], iv["exts"]):
            shards.setdefault(int(b), []).append((c[int(st):int(st) + int(ln)], int(e)))
    graphs = []
    for b in sorted(shards):
        items = shards[b] * (2 if twice else 1)
        w2, s2, l2 = orc.seqset_from_lists([x[0] for x in items])
        t = orc.filter_kmers(k, w2, s2, l2, seq_exts=np.array([x[1] for x in items], np.uint8), min_obs=2 if twice else 1, stranded=stranded)
        g = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded, reduce_op=reduce_op)
        assert g["error"] == 0
        graphs.append(g)
    return graphs
