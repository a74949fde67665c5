"""CPU tests pinning the oracle (oracle/oracle.cpp) before anything is compared against it.

Sources of truth, in order: (1) known-answer vectors held by the reference's own tests/doctests,
(2) the properties the reference's integration tests assert (src/test.rs P1-P6, src/msp.rs (1)-(4)),
(3) SURVEY.md Appendix B anchors from an independent restatement."""
import json
import os

import numpy as np
import pytest

from helpers import (assert_graphs_equal, canon, canon_form, dec, enc, kmers_of, msp_shard_graphs, node_bases, random_contigs,
                     random_dna, rc_int, simple_random_contigs, small_k_contigs)

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- (1) KATs from the reference ----------------------------------------------------------------
def test_kmer_doctest_kat(orc):
    """src/kmer.rs:14-34 (Kmer16 doctest) + SURVEY §7.1 K31 KAT."""
    L = orc.lib()
    k1 = kmers_of(enc("ACGTACGTACGTACGT"), 16)[0]
    assert L.orc_kmer_rc(16, L.orc_kmer_rc(16, k1)) == k1
    assert L.orc_kmer_extend_left(16, k1, 3) == kmers_of(enc("TACGTACGTACGTACG"), 16)[0]
    ks = sorted(kmers_of(enc("TACGTACGTACGTACGTT"), 16))
    assert ks == [kmers_of(enc(s), 16)[0] for s in ("ACGTACGTACGTACGT", "CGTACGTACGTACGTT", "TACGTACGTACGTACG")]
    s31 = ("ACGT" * 8)[:31]
    x = kmers_of(enc(s31), 31)[0]
    assert x == 0x06c6c6c6c6c6c6c6
    assert L.orc_kmer_rc(31, x) == 0x1b1b1b1b1b1b1b1b == kmers_of(enc(("CGTA" * 8)[:31]), 31)[0]


@pytest.mark.parametrize("k", [4, 5, 6, 8, 10, 12, 14, 15, 16, 20, 24, 31, 32])
def test_kmer_ops_property(orc, k):
    """src/kmer.rs:848-905: rc∘rc = id, rc base identity, extend_left/right."""
    L = orc.lib()
    rng = np.random.default_rng(k)
    for _ in range(300):
        b = random_dna(rng, k)
        x = kmers_of(b, k)[0]
        r = L.orc_kmer_rc(k, x)
        assert r == rc_int(x, k) and L.orc_kmer_rc(k, r) == x
        v = int(rng.integers(0, 4))
        assert L.orc_kmer_extend_right(k, x, v) == kmers_of(np.append(b[1:], v), k)[0]
        assert L.orc_kmer_extend_left(k, x, v) == kmers_of(np.insert(b[:-1], 0, v), k)[0]


@pytest.mark.parametrize("k", [33, 40, 48, 63, 64])
def test_kmer_rc128(orc, k):
    import ctypes as C
    L = orc.lib()
    rng = np.random.default_rng(k)
    for _ in range(100):
        x = kmers_of(random_dna(rng, k), k)[0]
        i = np.array([x & (2**64 - 1), x >> 64], dtype=np.uint64)
        o = np.zeros(2, dtype=np.uint64)
        L.orc_kmer_rc128(k, i.ctypes.data_as(C.POINTER(C.c_uint64)), o.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert (int(o[1]) << 64) | int(o[0]) == rc_int(x, k)


def test_exts_kat(orc):
    """src/lib.rs:729-748; SURVEY §7.1: Exts(0x12).rc() == 0x48."""
    L = orc.lib()
    assert L.orc_exts_rc(0x12) == 0x48
    for v in range(256):
        assert L.orc_exts_rc(L.orc_exts_rc(v)) == v
        c = L.orc_exts_complement(v)
        # complement reverses the 4 bits inside each nibble (A<->T, C<->G)
        for nib in (0, 4):
            for i in range(4):
                assert ((v >> (nib + i)) & 1) == ((c >> (nib + 3 - i)) & 1)


def test_dna_string_layout_kat(orc):
    """Bases of src/dna_string.rs:937-951 test_push_bytes ([2,0,0,0,0,1,1,0]) in the storage layout
    of :383-399: first base in bits 63..62 of word 0 => top 16 bits 10000000 00010100."""
    w = orc.pack_bases(np.array([2, 0, 0, 0, 0, 1, 1, 0], dtype=np.uint8))
    assert int(w[0]) >> 48 == 0x8014
    b = np.arange(70, dtype=np.uint8) % 4
    w = orc.pack_bases(b)
    assert len(w) == 3 and np.array_equal(orc.unpack_bases(w, 0, 70), b)


DNA142 = ("TGCATTAGAAAACTCCTTGCCTGTCAGCCCGACAGGTAGAAACTCATTAATCCACACATTGA"
          "CTCTATTTCAGGTAAATATGACGTCAACTCCTGCATGTTGAAGGCAGTGAGTGGCTGAAACAGCATCAAGGCGTGAAGGC")


def test_kmers_142(orc):
    """src/dna_string.rs:1061-1068 test_kmers string; model-derived check from SURVEY App. B:
    K=31, min_obs 1, unstranded => V=112, one node == input, exts 0, data 112."""
    w, s, l = orc.seqset_from_lists([enc(DNA142)])
    t = orc.filter_kmers(31, w, s, l, min_obs=1)
    assert len(t["lo"]) == 112
    assert int(t["lo"][0]) == 0x001d7e5ed25612b2 and int(t["exts"][0]) == 0x14
    assert set(int(x) for x in t["lo"]) == set(canon(x, 31) for x in kmers_of(enc(DNA142), 31))
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], t["counts"])
    assert g["n_nodes"] == 1 and dec(node_bases(orc, g, 0)) == DNA142
    assert int(g["exts"][0]) == 0 and int(g["data"][0]) == 112


def test_degen_seq_asm(orc):
    """src/test.rs:169-180 input; expected output per SURVEY App. B model."""
    ctg = enc("AAAAATAAAATAAAATAAAATAAAATAAAATAAAATAAAATAAAA")
    w, s, l = orc.seqset_from_lists([ctg, ctg])
    t = orc.filter_kmers(31, w, s, l, min_obs=2)
    assert len(t["lo"]) == 6
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], t["counts"])
    assert g["n_nodes"] == 2
    assert dec(node_bases(orc, g, 0)) == "AAAAATAAAATAAAATAAAATAAAATAAAAT" and g["exts"][0] == 0x10 and g["data"][0] == 2
    assert dec(node_bases(orc, g, 1)) == "AAAATAAAATAAAATAAAATAAAATAAAATAAAAT" and g["exts"][1] == 0x19 and g["data"][1] == 28


# ---- (3) anchors --------------------------------------------------------------------------------
with open(os.path.join(HERE, "golden", "anchors.json")) as f:
    ANCHORS = json.load(f)["rows"]


@pytest.mark.parametrize("row", ANCHORS, ids=lambda r: f"R{r[0]}-k{r[1]}-{'s' if r[2] else 'u'}-{'noisy' if r[3] else 'clean'}")
def test_anchor(orc, row):
    R, k, stranded, noisy, mo, N, U, V, M, Lb, mx, xv, mv, fn = row
    w, s, l = orc.synth_reads(R, 1, orc.ERR_THR_NOISY if noisy else 0)
    t = orc.filter_kmers(k, w, s, l, min_obs=mo, stranded=stranded, report_all=True)
    assert (t["n_input"], len(t["all_lo"]), len(t["lo"])) == (N, U, V)
    assert orc.xor_valid(t) == int(xv, 16) and orc.mix_valid(t) == int(mv, 16)
    g = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded)
    assert g["error"] == 0
    assert (g["n_nodes"], g["n_bases"], int(g["length"].max())) == (M, Lb, mx)
    assert orc.fnv_nodes(g) == int(fn, 16)


def test_first_bases(orc):
    assert dec(orc.unpack_bases(orc.synth_reads(1000)[0], 0, 20)) == "GGTTCAGGTAACCTTCATTA"
    assert dec(orc.unpack_bases(orc.synth_reads(10000)[0], 0, 20)) == "TTAAGGCGCTAATTGATTCC"


# ---- (2) properties of the reference's integration tests ---------------------------------------
def _check_graph_props(orc, k, contigs, t, g, stranded=False):
    """P1, P3, P4 of src/test.rs:351-355, 388-413."""
    cf = (lambda x: x) if stranded else (lambda x: canon(x, k))
    kmer_set = set(cf(x) for c in contigs for x in kmers_of(c, k))
    assert set(int(x) for x in t["lo"]) == kmer_set                              # P1
    allc = set()
    for i in range(g["n_nodes"]):
        b = node_bases(orc, g, i)
        ks = kmers_of(b, k)
        cs = set(cf(x) for x in ks)
        assert cs <= kmer_set and len(cs) == len(ks)
        assert not (cs & allc)
        allc |= cs
        e = int(g["exts"][i])
        mask = (1 << (2 * k)) - 1
        for base in range(4):                                                     # P4
            if e & (1 << base):
                assert cf((ks[0] >> 2) | (base << (2 * (k - 1)))) in kmer_set
            if e & (1 << (4 + base)):
                assert cf(((ks[-1] << 2) & mask) | base) in kmer_set
    assert allc == kmer_set                                                       # P3


@pytest.mark.parametrize("k", [31, 32])
def test_reassemble_contigs(orc, k):
    """src/test.rs:299-414 (unsharded variant: whole contigs, each given twice, CountFilter(2))."""
    rng = np.random.default_rng(100 + k)
    for it in range(6):
        contigs = simple_random_contigs(rng) if it == 0 else random_contigs(rng)
        contigs = [c for c in contigs if len(c) >= k]
        w, s, l = orc.seqset_from_lists(contigs + contigs)
        t = orc.filter_kmers(k, w, s, l, min_obs=2)
        g = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], reduce_op=orc.SAT_ADD)
        assert g["error"] == 0
        _check_graph_props(orc, k, contigs, t, g)
        # P2: every k-mer reachable via Exts, none invented (test.rs:357-381)
        kmer_set = set(int(x) for x in t["lo"])
        ext_set = set()
        mask = (1 << (2 * k)) - 1
        for x, e in zip(t["lo"], t["exts"]):
            x, e = int(x), int(e)
            for base in range(4):
                if e & (1 << base):
                    ext_set.add(canon((x >> 2) | (base << (2 * (k - 1))), k))
                if e & (1 << (4 + base)):
                    ext_set.add(canon(((x << 2) & mask) | base, k))
        assert ext_set <= kmer_set
        # memory_size only changes the number of passes (filter.rs:151-168)
        t1 = orc.filter_kmers(k, w, s, l, min_obs=2, memory_gb=1)
        assert np.array_equal(t1["lo"], t["lo"]) and np.array_equal(t1["exts"], t["exts"])


@pytest.mark.parametrize("k,stranded", [(4, False), (5, False), (6, False), (7, False), (5, True), (6, True)])
def test_small_k_cycles_palindromes(orc, k, stranded):
    rng = np.random.default_rng(k * 7 + stranded)
    for it in range(60):
        contigs = small_k_contigs(rng, alphabet=2 if it % 3 == 0 else 4)
        contigs = [c for c in contigs if len(c) >= k]
        if not contigs:
            continue
        w, s, l = orc.seqset_from_lists(contigs)
        t = orc.filter_kmers(k, w, s, l, min_obs=1, stranded=stranded)
        g = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded)
        assert g["error"] == 0
        _check_graph_props(orc, k, contigs, t, g, stranded)


def test_seed_order_changes_only_strand_and_cycle_breaks(orc):
    """canonical form of the graph is invariant to seed (boomphf slot) order for
    non-cyclic components."""
    rng = np.random.default_rng(5)
    k = 31
    contigs = [c for c in random_contigs(rng) if len(c) >= k]
    w, s, l = orc.seqset_from_lists(contigs)
    t = orc.filter_kmers(k, w, s, l, min_obs=1)
    g0 = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"])

    def canon_form(g):
        out = []
        L = orc.lib()
        for i in range(g["n_nodes"]):
            b = node_bases(orc, g, i)
            r = (3 - b[::-1]).astype(np.uint8)
            e = int(g["exts"][i])
            if bytes(r) < bytes(b):
                b, e = r, L.orc_exts_rc(e)
            out.append((bytes(b), e, int(g["data"][i])))
        return sorted(out)

    perm = rng.permutation(len(t["lo"])).astype(np.uint32)
    g1 = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], seed_order=perm)
    assert canon_form(g0) == canon_form(g1)


def test_threads_variant_identical(orc):
    w, s, l = orc.synth_reads(2000, 1, orc.ERR_THR_NOISY)
    a = orc.filter_kmers(31, w, s, l, min_obs=2, report_all=True)
    b = orc.filter_kmers(31, w, s, l, min_obs=2, report_all=True, threads=4)
    for f in ("lo", "exts", "counts", "all_lo"):
        assert np.array_equal(a[f], b[f])


def test_edge_cases(orc):
    """Empty input, reads shorter than K (lib.rs:783,813), min_obs=0, sequence-level exts (lib.rs:820-830)."""
    e = np.zeros(0, np.uint64)
    t = orc.filter_kmers(31, e, e, np.zeros(0, np.uint32))
    assert len(t["lo"]) == 0
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], t["counts"])
    assert g["n_nodes"] == 0 and g["n_bases"] == 0
    w, s, l = orc.seqset_from_lists([enc("ACGT"), enc("A" * 30)])
    assert len(orc.filter_kmers(31, w, s, l)["lo"]) == 0
    # a length-K sequence takes both nibbles of seq_exts verbatim
    seq = enc("ACGTTGCATGCATCGATCGATCGTAGCTAGA")
    w, s, l = orc.seqset_from_lists([seq])
    t = orc.filter_kmers(31, w, s, l, seq_exts=np.array([0x5a], np.uint8), stranded=True)
    assert len(t["lo"]) == 1 and int(t["exts"][0]) == 0x5a
    t = orc.filter_kmers(31, w, s, l, seq_exts=np.array([0x5a], np.uint8), min_obs=0)
    x = kmers_of(seq, 31)[0]
    assert int(t["lo"][0]) == canon(x, 31)
    assert int(t["exts"][0]) == (0x5a if x < rc_int(x, 31) else orc.lib().orc_exts_rc(0x5a))


def test_count_saturation(orc):
    """filter.rs:57: counts saturate at 65535."""
    seq = enc("ACGTTGCATGCATCGATCGATCGTAGCTAGA")
    w, s, l = orc.seqset_from_lists([seq] * 70000)
    t = orc.filter_kmers(31, w, s, l, min_obs=1)
    assert len(t["lo"]) == 1 and int(t["counts"][0]) == 65535


# ---- MSP properties — src/msp.rs:404-485 --------------------------------------------------------
@pytest.mark.parametrize("k,p", [(16, 5), (31, 6), (31, 8), (35, 5), (50, 8), (63, 12)])
def test_msp_scanner_properties(orc, k, p):
    rng = np.random.default_rng(k * 100 + p)
    for it in range(30):
        m = int(rng.integers(k, 6 * k))
        seq = random_dna(rng, m) if it % 5 else np.zeros(m, np.uint8)  # incl. DnaString::blank (msp.rs:517-528)
        iv = orc.msp_scan(k, p, seq, rc=False)  # score = p.to_u64() as in test_new_slicer (msp.rs:494)
        pm = kmers_of(seq, p)
        covered = np.zeros(m - k + 1, bool)
        for st, ln, mp, mz in zip(iv["start"], iv["len"], iv["minpos"], iv["minimizer"]):
            st, ln, mp, mz = int(st), int(ln), int(mp), int(mz)
            assert not covered[st:st + ln - k + 1].any()
            covered[st:st + ln - k + 1] = True                      # (1)
            assert p <= ln <= 2 * k - p                              # (2)
            assert pm[mp] == mz and st <= mp <= st + ln - p          # (3)
            assert min(pm[st:st + ln - p + 1]) >= mz
        assert covered.all()
        for i in range(len(iv["start"]) - 1):                        # (4) right-maximal
            st, ln, mp, mz = (int(iv[f][i]) for f in ("start", "len", "minpos", "minimizer"))
            nk = st + ln - k + 1
            assert pm[st + ln - p + 1] < mz or not (nk <= mp)


def test_msp_sharded_filter_equals_unsharded(orc):
    """The reference's own sharded flow (src/test.rs:433-456): per-bucket filter_kmers over MSP
    substrings with from_slice_bounds exts reproduces the unsharded (k-mer, exts, count) set."""
    rng = np.random.default_rng(9)
    k, p = 31, 6
    contigs = [c for c in random_contigs(rng) if len(c) >= k]
    w, s, l = orc.seqset_from_lists(contigs)
    ref = orc.filter_kmers(k, w, s, l, min_obs=1)
    shards = {}
    for c in contigs:
        iv = orc.msp_scan(k, p, c, rc=True)
        for st, ln, b, e in zip(iv["start"], iv["len"], iv["bucket"], iv["exts"]):
            shards.setdefault(int(b), []).append((c[int(st):int(st) + int(ln)], int(e)))
    got = {}
    for b, items in shards.items():
        w2, s2, l2 = orc.seqset_from_lists([x[0] for x in items])
        t = orc.filter_kmers(k, w2, s2, l2, seq_exts=np.array([x[1] for x in items], np.uint8), min_obs=1)
        for x, e, c in zip(t["lo"], t["exts"], t["counts"]):
            assert int(x) not in got
            got[int(x)] = (int(e), int(c))
    assert got == {int(x): (int(e), int(c)) for x, e, c in zip(ref["lo"], ref["exts"], ref["counts"])}


def test_from_acgt_bytes_matches_reference_kats(orc):
    """DnaString::from_acgt_bytes (dna_string.rs:224-250): same words as from_dna_string on the reference's own test
    strings (test_from_dna_string, dna_string.rs:954-975), case-insensitive, non-ACGT -> A (lib.rs:65-73;
    bitops_avx2.rs test_invalid_bases :196-215 checks exactly this validity rule)."""
    dna = ("TGCATTAGAAAACTCCTTGCCTGTCAGCCCGACAGGTAGAAACTCATTAATCCACACATTGA"
           "CTCTATTTCAGGTAAATATGACGTCAACTCCTGCATGTTGAAGGCAGTGAGTGGCTGAAACAGCATCAAGGCGTGAAGGC")   # dna_string.rs:1062
    w, st, ln, bad = orc.from_acgt_bytes([dna.encode()])
    assert bad == 0 and list(ln) == [142] and np.array_equal(w, orc.pack_bases(enc(dna)))
    w2, _, _, bad2 = orc.from_acgt_bytes([dna.lower().encode()])
    assert bad2 == 0 and np.array_equal(w2, w)
    w3, _, _, bad3 = orc.from_acgt_bytes([b"ACGTNNacgtXy-\x00\xff"])
    assert bad3 == 7 and np.array_equal(w3, orc.pack_bases(enc("ACGTAAACGTAAAAA")))
    # doctest string of dna_string.rs:19: k-mer 0 of slice(10, 40) is CACGTATGACAGATAG
    s = "ACAGCAGCAGCACGTATGACAGATAGTGACAGCAGTTTGTGACCGCAAGAGCAGTAATATGATG"
    w4, _, _, _ = orc.from_acgt_bytes([s.encode()])
    assert np.array_equal(orc.unpack_bases(w4, 10, 16), enc("CACGTATGACAGATAG"))
    # several sequences are appended bit-contiguously (PackedDnaStringSet::add, dna_string.rs:811-821)
    w5, st5, ln5, _ = orc.from_acgt_bytes([b"ACG", b"", b"TTTTT", b"g"])
    assert list(st5) == [0, 3, 3, 8] and list(ln5) == [3, 0, 5, 1]
    assert np.array_equal(w5, orc.pack_bases(enc("ACGTTTTTG")))


def test_remove_censored_exts_properties(orc):
    """filter::remove_censored_exts[_sharded] (filter.rs:238-306; the reference has no test of its own for them): every kept
    extension points at a valid k-mer, nothing is added, the operation is idempotent; the sharded variant keeps exactly the
    extensions whose target was never seen in the shard (here: the ones injected through sequence-level Exts)."""
    rng = np.random.default_rng(9)
    w, st, ln = orc.synth_reads(1500, 1, orc.ERR_THR_NOISY)
    for k, stranded in ((31, False), (32, True), (63, False)):
        sx = rng.integers(0, 256, size=len(st)).astype(np.uint8)
        t = orc.filter_kmers(k, w, st, ln, seq_exts=sx, min_obs=2, stranded=stranded, report_all=True)
        e1 = orc.remove_censored_exts(k, t, stranded=stranded)
        e2 = orc.remove_censored_exts(k, t, stranded=stranded, sharded=True)
        assert np.all(e1 & ~t["exts"] == 0) and np.all(e2 & ~t["exts"] == 0) and np.all(e1 & ~e2 == 0)
        assert (e1 != t["exts"]).any() and (e2 != e1).any()
        t1 = dict(t, exts=e1)
        assert np.array_equal(orc.remove_censored_exts(k, t1, stranded=stranded), e1)
        # with pruned Exts nothing points outside the table any more: compression merges the paths the dangling bits split
        g0 = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded)
        g1 = orc.compress_kmers(k, t["lo"], t["hi"], e1, t["counts"], stranded=stranded)
        assert g1["error"] == 0 and g1["n_nodes"] < g0["n_nodes"] and g1["n_bases"] - g1["n_nodes"] * (k - 1) == len(t["lo"])


def test_graph_edges_and_is_compressed(orc):
    """BaseGraph::finish + find_edges / find_link (graph.rs:116-142, 223-291) restated, pinned by the property the reference's
    own tests assert after compress_kmers: `is_compressed(&spec) == None` (test.rs:249-254, 268-274) — on tables where every
    k-mer is valid (the reference's case: each contig twice / CountFilter(1)) and on censored tables after
    remove_censored_exts.  Edges are reciprocal (odd K): the target's edge list on the incoming side leads back."""
    from helpers import random_contigs
    rng = np.random.default_rng(5)
    cases = []
    for k, stranded in ((31, False), (32, False), (6, False), (5, True), (33, True)):
        contigs = [c for c in random_contigs(rng) if len(c) >= k]
        ss = orc.seqset_from_lists(contigs + contigs)
        t = orc.filter_kmers(k, *ss, min_obs=2, stranded=stranded)
        cases.append((k, stranded, t, t["exts"]))
    w, st, ln = orc.synth_reads(1500, 1, orc.ERR_THR_NOISY)
    t = orc.filter_kmers(31, w, st, ln, min_obs=2)
    cases.append((31, False, t, orc.remove_censored_exts(31, t)))
    for k, stranded, t, exts in cases:
        g = orc.compress_kmers(k, t["lo"], t["hi"], exts, t["counts"], stranded=stranded)
        assert g["error"] == 0
        target, flags, pair = orc.graph_edges(k, g, stranded=stranded)
        assert pair is None, (k, stranded, pair)
        m = g["n_nodes"]
        if k % 2 == 0:
            continue   # even K: palindromic k-mers (flipped-branch quirk, lib.rs:224-231) make find_link's answers one-sided
        for n in range(0, m, max(1, m // 300)):
            for d in range(2):
                for i in range(4):
                    if target[n, d, i] != 0xffffffff:
                        nxt, inc = int(target[n, d, i]), int(flags[n, d, i] & 1)
                        assert n in [int(x) for x in target[nxt, inc]], "edge without a way back"
    # the uncensored noisy table is NOT compressed in is_compressed's sense: dangling Exts keep paths apart (SURVEY §4 caveat)
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], t["counts"])
    assert orc.graph_edges(31, g)[2] is not None


def test_graph_fix_exts_properties(orc):
    """DebruijnGraph::fix_exts (graph.rs:337-377): afterwards every node extension resolves to a node (no dangling bits),
    nothing is added, idempotent; with valid_nodes, extensions into unmarked nodes go too."""
    w, st, ln = orc.synth_reads(1500, 1, orc.ERR_THR_NOISY)
    t = orc.filter_kmers(31, w, st, ln, min_obs=2)
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], t["counts"])
    ne = orc.graph_fix_exts(31, g)
    assert np.all(ne & ~g["exts"] == 0) and (ne != g["exts"]).any()
    g2 = dict(g, exts=ne)
    target, _, _ = orc.graph_edges(31, g2)
    popc = np.array([bin(int(x)).count("1") for x in ne])
    assert np.array_equal((target != 0xffffffff).sum(axis=(1, 2)), popc)
    assert np.array_equal(orc.graph_fix_exts(31, g2), ne)
    vn = np.arange(g["n_nodes"]) % 3 != 0
    nv = orc.graph_fix_exts(31, g, valid_nodes=vn)
    assert np.all(nv & ~ne == 0) and (nv != ne).any()


def test_scmap_compress_properties(orc):
    """ScmapCompress (compression.rs:66-98): join_test = data equality.  Every node's k-mers carry one data value (its own
    data), two k-mers of a node never differ, and with all data equal the result is SimpleCompress's."""
    w, st, ln = orc.synth_reads(1500, 1, orc.ERR_THR_NOISY)
    t = orc.filter_kmers(31, w, st, ln, min_obs=2)
    counts = (t["counts"] % 3).astype(np.uint16)           # few distinct values: long equal-data runs and many breaks
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], counts, reduce_op=orc.SCMAP)
    g0 = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], counts, reduce_op=orc.MAX)
    assert g["error"] == 0 and g["n_nodes"] > g0["n_nodes"] and g["n_bases"] - g["n_nodes"] * 30 == len(t["lo"])
    # data of every k-mer of a node equals the node's data: look each node k-mer up in the table
    idx = {int(k_): i for i, k_ in enumerate(t["lo"])}
    mask = (1 << 62) - 1
    for n in range(0, g["n_nodes"], 7):
        b = orc.unpack_bases(g["words"], int(g["start"][n]), int(g["length"][n]))
        x = 0
        for j, base in enumerate(b):
            x = ((x << 2) | int(base)) & mask
            if j >= 30:
                r = 0
                for q in range(31):
                    r |= (3 - ((x >> (2 * q)) & 3)) << (2 * (30 - q))
                assert counts[idx[min(x, r)]] == g["data"][n]
    same = np.full(len(counts), 5, np.uint16)
    ga = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], same, reduce_op=orc.SCMAP)
    gb = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], same, reduce_op=orc.MAX)
    assert all(np.array_equal(ga[f], gb[f]) for f in ("words", "start", "length", "exts", "data"))


def test_write_gfa_format(orc):
    """DebruijnGraph::write_gfa (graph.rs:538-614): header, one S line per node in node order, every link reported once
    (left edges with target >= node, right edges with target > node), overlap K-1."""
    seq = "ACGTTGCATGCATCGATCGATCGTAGCTAGAGGATCCATTAGC"
    a, b = seq + "A" + "GGTCAAT" * 5, seq + "C" + "TTGACAT" * 5      # a fork after the shared prefix
    ss = orc.seqset_from_lists([enc(a), enc(a), enc(b), enc(b)])
    t = orc.filter_kmers(31, *ss, min_obs=2)
    g = orc.compress_kmers(31, t["lo"], t["hi"], t["exts"], t["counts"])
    text = orc.write_gfa(31, g)
    lines = text.splitlines()
    assert lines[0] == "H\tVN:Z:debruijn-rs"
    s_lines = [l for l in lines if l.startswith("S\t")]
    l_lines = [l for l in lines if l.startswith("L\t")]
    assert len(s_lines) == g["n_nodes"] == 3 and [int(l.split("\t")[1]) for l in s_lines] == [0, 1, 2]
    assert len(l_lines) == 2 and all(l.endswith("\t30M") for l in l_lines)
    for l in l_lines:
        f = l.split("\t")
        assert f[2] in "+-" and f[4] in "+-" and int(f[1]) <= int(f[3])
    assert sum(len(l.split("\t")[2]) for l in s_lines) == g["n_bases"]


def test_widened_rows_golden(orc):
    """The oracle's outputs for the widened rows (remove_censored_exts, graph edges / is_compressed, fix_exts, GFA text,
    ScmapCompress, from_acgt_bytes) equal the committed vectors of tests/golden/widened.json (made by
    tests/golden/make_golden.py from this same oracle: a regression pin, the reference has no expected outputs here)."""
    import importlib.util
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = json.load(open(os.path.join(here, "golden", "widened.json")))["rows"]
    assert mg.rows() == want
    # one number here IS independently anchored: SURVEY.md §4 (second, independent restatement made during the survey) found
    # 68 nodes for synth-v1 noisy R=1000, K=31 after remove_censored_exts semantics (896 without pruning, Appendix B)
    assert want[0]["pruned_nodes"] == 68 and want[0]["n_valid"] == 3584


def test_siphash13_matches_cpython_and_the_library(orc):
    """The SipHash-1-3 behind from_acgt_bytes_hashn (Rust's DefaultHasher, zero key) is pinned twice: the oracle's plain-Python
    version and the library's C version (dbg_siphash13, host code: no GPU needed) both equal CPython's bytes hash under
    PYTHONHASHSEED=0 (CPython >= 3.11 uses SipHash-1-3 and an all-zero key when randomisation is off)."""
    import ctypes as C
    import subprocess
    import sys

    from rust_debruijn_b200 import _lib
    rng = np.random.default_rng(4)
    msgs = [b"", b"a", b"read/1", b"01234567", b"0123456789abcdef!"] + [bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)) for n in (3, 8, 15, 16, 31, 64, 200)]
    code = "import sys; print([hash(bytes.fromhex(x)) & 0xffffffffffffffff for x in sys.argv[1:]])"
    out = subprocess.run([sys.executable, "-c", code] + [m.hex() for m in msgs[1:]], capture_output=True, text=True,
                         env={"PYTHONHASHSEED": "0", "PATH": "/usr/bin:/bin"}).stdout
    ref = eval(out)
    L = _lib.lib()
    for m, r in zip(msgs[1:], ref):
        mine = orc.siphash13(m)
        if r != (-2) & 0xffffffffffffffff:   # CPython maps the value -1 to -2
            assert mine == r, m
        buf = np.frombuffer(m, np.uint8)
        assert L.dbg_siphash13(buf.ctypes.data_as(C.c_void_p), len(m)) == mine
    assert L.dbg_siphash13(None, 0) == orc.siphash13(b"")


def test_from_acgt_bytes_hashn_oracle(orc):
    """ACGT positions are untouched, non-ACGT ones are repeatable, depend on the read name and the position, and cover 0..3."""
    seqs = [b"ACGTNNNNacgtRYKM" * 8, b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", b"ACGT"]
    names = [b"read_1", b"read_2", b"x"]
    w, s, l, bad = orc.from_acgt_bytes_hashn(seqs, names)
    w2, _, _, bad2 = orc.from_acgt_bytes_hashn(seqs, names)
    assert np.array_equal(w, w2) and bad == bad2 == 8 * 8 + 40
    plain = orc.from_acgt_bytes(seqs)[0]
    b_h, b_p = orc.unpack_bases(w, 0, int(l.sum())), orc.unpack_bases(plain, 0, int(l.sum()))
    ok = np.concatenate([orc.acgt_to_bits(x)[1] for x in seqs])
    assert np.array_equal(b_h[ok], b_p[ok]) and set(b_h[~ok].tolist()) == {0, 1, 2, 3}
    w3 = orc.from_acgt_bytes_hashn(seqs, [b"read_9", b"read_2", b"x"])[0]
    assert not np.array_equal(w, w3)


def test_bincode_image_layout(orc):
    """Hand-derived bytes of a two-node BaseGraph in bincode 1.x's default encoding (serde field order of the crate)."""
    g = dict(words=np.array([0x1b00000000000000], np.uint64), n_bases=7, start=np.array([0, 4], np.uint64), length=np.array([4, 3], np.uint32),
             exts=np.array([0x12, 0x80], np.uint8), data=np.array([5, 65535], np.uint16), stranded=True)
    img = orc.graph_to_bincode(g)
    exp = (b"\x01" + b"\0" * 7 + bytes.fromhex("000000000000001b") + b"\x07" + b"\0" * 7 +
           b"\x02" + b"\0" * 7 + b"\0" * 8 + b"\x04" + b"\0" * 7 +
           b"\x02" + b"\0" * 7 + b"\x04\0\0\0\x03\0\0\0" +
           b"\x02" + b"\0" * 7 + b"\x12\x80" +
           b"\x02" + b"\0" * 7 + b"\x05\x00\xff\xff" + b"\x01")
    assert img == exp


def _reassemble_sharded_props(orc, k, contigs, g):
    """The assertions of reassemble_sharded (src/test.rs:473-503) on the stitched graph."""
    kmer_set = set(canon(x, k) for c in contigs for x in kmers_of(c, k))
    allc = set()
    mask = (1 << (2 * k)) - 1
    for i in range(g["n_nodes"]):
        ks = kmers_of(node_bases(orc, g, i), k)
        cs = set(canon(x, k) for x in ks)
        assert cs <= kmer_set
        allc |= cs
        e = int(g["exts"][i])
        for base in range(4):
            if e & (1 << base):
                assert canon((ks[0] >> 2) | (base << (2 * (k - 1))), k) in kmer_set
            if e & (1 << (4 + base)):
                assert canon(((ks[-1] << 2) & mask) | base, k) in kmer_set
    assert allc == kmer_set


@pytest.mark.parametrize("k", [31, 32])
def test_compress_graph_reassemble_sharded(orc, k):
    """src/test.rs:418-504: msp shards -> shard assemblies -> BaseGraph::combine -> compress_graph(max): the reference's own
    assertions, plus: the stitched graph is compressed (compression.rs:332) and its canonical form (sequences + Exts) equals the
    unsharded compress_kmers graph of the same contigs (every k-mer valid, so fix_exts removes nothing)."""
    rng = np.random.default_rng(300 + k)
    for it in range(5):
        contigs = simple_random_contigs(rng) if it == 0 else random_contigs(rng)
        contigs = [c for c in contigs if len(c) >= k]
        shard_graphs = msp_shard_graphs(orc, k, 6, contigs)
        combined = orc.combine_graphs(shard_graphs)
        assert combined["n_nodes"] == sum(g["n_nodes"] for g in shard_graphs)
        g = orc.compress_graph(k, combined, stranded=False, reduce_op=orc.MAX)
        assert g["error"] == 0
        _reassemble_sharded_props(orc, k, contigs, g)
        assert orc.graph_edges(k, g)[2] is None
        w, s, l = orc.seqset_from_lists(contigs + contigs)
        t = orc.filter_kmers(k, w, s, l, min_obs=2)
        whole = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], reduce_op=orc.MAX)
        assert canon_form(orc, g, with_data=False) == canon_form(orc, whole, with_data=False)


@pytest.mark.parametrize("k,stranded", [(4, False), (5, False), (6, False), (5, True), (6, True)])
def test_compress_graph_small_k(orc, k, stranded):
    """Per-k-mer nodes (an uncompressed graph) of cycle / hairpin / palindrome rich inputs: compress_graph of the trivial graph has the
    canonical form of compress_kmers for acyclic components; with censored nodes the survivors still tile their k-mers."""
    rng = np.random.default_rng(k * 11 + stranded)
    for it in range(40):
        contigs = [c for c in small_k_contigs(rng, alphabet=2 if it % 3 == 0 else 4) if len(c) >= k]
        if not contigs:
            continue
        w, s, l = orc.seqset_from_lists(contigs)
        t = orc.filter_kmers(k, w, s, l, min_obs=1, stranded=stranded)
        n = len(t["lo"])
        # one node per k-mer, Exts and counts of the table
        kb = np.array([[(int(x) >> (2 * (k - 1 - j))) & 3 for j in range(k)] for x in t["lo"]], np.uint8).reshape(-1)
        triv = dict(n_nodes=n, n_bases=n * k, words=orc.pack_bases(kb), start=np.arange(n, dtype=np.uint64) * k,
                    length=np.full(n, k, np.uint32), exts=t["exts"], data=t["counts"], stranded=stranded)
        g = orc.compress_graph(k, triv, stranded=stranded, reduce_op=orc.SAT_ADD)
        assert g["error"] == 0
        whole = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded)
        # same seed order (ascending k-mer = node order) and the same rule: identical graphs, bit for bit
        assert_graphs_equal(g, whole)
        if n > 3:
            censor = sorted(set(int(x) for x in rng.integers(0, n, size=max(1, n // 5))))
            gc = orc.compress_graph(k, triv, stranded=stranded, censor_nodes=censor)
            assert gc["error"] == 0
            keep = set(int(x) for i, x in enumerate(t["lo"]) if i not in censor)
            cf = (lambda x: x) if stranded else (lambda x: canon(x, k))
            got = [cf(x) for i in range(gc["n_nodes"]) for x in kmers_of(node_bases(orc, gc, i), k)]
            assert sorted(got) == sorted(keep)
            assert orc.graph_edges(k, gc, stranded=stranded)[2] is None
