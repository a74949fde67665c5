"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
inputs.  Bar: bit-exact (integer / byte / index work).  Mirrors the reference's own integration
tests (src/test.rs:157-231 drivers) plus the edge cases they cover."""
import json
import os

import numpy as np
import pytest

from helpers import (assert_graphs_equal, assert_tables_equal, brute_filter_colorset, enc, msp_shard_graphs, random_contigs, random_dna,
                     simple_random_contigs, small_k_contigs)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def D():
    import rust_debruijn_b200 as D
    return D


@pytest.fixture(scope="module")
def ctx(D):
    return D.Context(0)


def run_both(D, ctx, orc, k, seqset, min_obs, stranded=False, report_all=False, reduce_op=0, seq_exts=None):
    words, start, length = seqset
    seqs = (words, start, length) if seq_exts is None else (words, start, length, seq_exts)
    table, _ = D.filter_kmers(seqs, D.CountFilter(min_obs), stranded, report_all, 4, k=k, ctx=ctx)
    t = table.to_host()
    ot = orc.filter_kmers(k, words, start, length, seq_exts=seq_exts, min_obs=min_obs, stranded=stranded,
                          report_all=report_all)
    assert t["n_input"] == ot["n_input"]
    assert_tables_equal(t, ot)
    graph = D.compress_kmers_with_hash(stranded, D.SimpleCompress(reduce_op), table)
    g = graph.to_host()
    og = orc.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"], stranded=stranded, reduce_op=reduce_op)
    assert og["error"] == 0
    assert_graphs_equal(g, og)
    return t, g


with open(os.path.join(HERE, "golden", "anchors.json")) as f:
    ANCHORS = json.load(f)["rows"]


@pytest.mark.parametrize("row", ANCHORS, ids=lambda r: f"R{r[0]}-k{r[1]}-{'s' if r[2] else 'u'}-{'noisy' if r[3] else 'clean'}")
def test_anchor_configs(D, ctx, orc, row):
    """C1-style configs (BASELINE.json configs[0] at R=10^4, plus K=63 / K=32 / stranded variants):
    bit-exact vs oracle AND vs the committed Appendix-B anchors."""
    R, k, stranded, noisy, mo, N, U, V, M, Lb, mx, xv, mv, fn = row
    ss = orc.synth_reads(R, 1, orc.ERR_THR_NOISY if noisy else 0)
    t, g = run_both(D, ctx, orc, k, ss, mo, stranded=stranded, report_all=True)
    assert (t["n_input"], len(t["all_lo"]), len(t["lo"]), g["n_nodes"], g["n_bases"]) == (N, U, V, M, Lb)
    assert orc.xor_valid(t) == int(xv, 16) and orc.mix_valid(t) == int(mv, 16) and orc.fnv_nodes(g) == int(fn, 16)


def test_device_synth_matches_oracle(D, ctx, orc):
    for thr in (0, orc.ERR_THR_NOISY):
        ss = D.SeqSet.synth(ctx, 3000, 1, thr)
        w, s, l = ss.copy_out()
        ow, os_, ol = orc.synth_reads(3000, 1, thr)
        assert np.array_equal(w, ow) and np.array_equal(s, os_) and np.array_equal(l, ol)


@pytest.mark.parametrize("k", [31, 32])
def test_reassemble_contigs(D, ctx, orc, k):
    """src/test.rs:299-414 flow on seeded random contigs (each contig twice, CountFilter(2), sat-add)."""
    rng = np.random.default_rng(1000 + k)
    for it in range(8):
        contigs = simple_random_contigs(rng) if it == 0 else random_contigs(rng)
        contigs = [c for c in contigs if len(c) >= k]
        run_both(D, ctx, orc, k, orc.seqset_from_lists(contigs + contigs), 2)


@pytest.mark.parametrize("reduce_op", [0, 1, 2, 3])
def test_simplify_from_kmers_reduce_ops(D, ctx, orc, reduce_op):
    """src/test.rs:233-254 flow (CountFilter(1)) with each SimpleCompress closure the tests use."""
    rng = np.random.default_rng(77 + reduce_op)
    contigs = [c for c in random_contigs(rng) if len(c) >= 31]
    run_both(D, ctx, orc, 31, orc.seqset_from_lists(contigs), 1, reduce_op=reduce_op)


@pytest.mark.parametrize("k,stranded", [(4, False), (5, False), (6, False), (7, False), (8, False), (5, True), (6, True),
                                        (12, False), (33, False), (34, False), (47, True), (64, False), (64, True), (32, True)])
def test_small_and_odd_k(D, ctx, orc, k, stranded):
    """Cycles, hairpins, palindromes (even K), self-links: dense at small K over small alphabets.
    Also the two-word key path edges (K=33/34/64) and the all-T k-mer at K=32/64 stranded."""
    rng = np.random.default_rng(k * 31 + stranded)
    for it in range(40 if k <= 12 else 8):
        if k <= 12:
            contigs = small_k_contigs(rng, alphabet=2 if it % 3 == 0 else 4)
        else:
            contigs = small_k_contigs(rng, n_ctg=5, lo=k, hi=4 * k, alphabet=2 if it % 2 == 0 else 4)
            contigs.append(np.full(k + 7, 3, np.uint8))   # T...T
            contigs.append(np.tile(np.array([0, 1, 3], np.uint8), k)[: 2 * k + 5])  # tandem repeat -> cycle
        contigs = [c for c in contigs if len(c) >= k]
        if contigs:
            run_both(D, ctx, orc, k, orc.seqset_from_lists(contigs), 1, stranded=stranded, report_all=(it % 4 == 0))


def test_degen_and_kat_inputs(D, ctx, orc):
    ctg = enc("AAAAATAAAATAAAATAAAATAAAATAAAATAAAATAAAATAAAA")  # src/test.rs:169-180
    t, g = run_both(D, ctx, orc, 31, orc.seqset_from_lists([ctg, ctg]), 2)
    assert g["n_nodes"] == 2 and list(g["data"]) == [2, 28]
    dna = enc("TGCATTAGAAAACTCCTTGCCTGTCAGCCCGACAGGTAGAAACTCATTAATCCACACATTGA"
              "CTCTATTTCAGGTAAATATGACGTCAACTCCTGCATGTTGAAGGCAGTGAGTGGCTGAAACAGCATCAAGGCGTGAAGGC")  # dna_string.rs:1062
    t, g = run_both(D, ctx, orc, 31, orc.seqset_from_lists([dna]), 1)
    assert g["n_nodes"] == 1 and g["length"][0] == 142


def test_edge_cases(D, ctx, orc):
    e = np.zeros(0, np.uint64)
    table, _ = D.filter_kmers((e, e, np.zeros(0, np.uint32)), D.CountFilter(1), False, False, 4, k=31, ctx=ctx)
    assert len(table) == 0
    assert len(D.compress_kmers_with_hash(False, D.SimpleCompress(), table)) == 0
    # shorter than K => nothing, not an error (lib.rs:783,813); ragged lengths; K-length read with both ext nibbles
    rng = np.random.default_rng(3)
    seqs = [random_dna(rng, n) for n in (3, 30, 31, 32, 33, 64, 65, 150, 151, 700, 5000)]
    sx = rng.integers(0, 256, size=len(seqs)).astype(np.uint8)
    for stranded in (False, True):
        run_both(D, ctx, orc, 31, orc.seqset_from_lists(seqs), 1, stranded=stranded, seq_exts=sx)
    run_both(D, ctx, orc, 31, orc.seqset_from_lists(seqs), 0)          # min_obs = 0: everything valid (filter.rs:61)
    run_both(D, ctx, orc, 31, orc.seqset_from_lists(seqs), 2)          # nothing valid
    with pytest.raises(D.DbgError):
        D.filter_kmers(orc.seqset_from_lists(seqs), D.CountFilter(1), False, False, 4, k=3, ctx=ctx)
    with pytest.raises(D.DbgError):
        D.filter_kmers(orc.seqset_from_lists(seqs), D.CountFilter(1), False, False, 4, k=65, ctx=ctx)


def test_layouts(D, ctx, orc):
    """Sequence-set layouts: uniform reads with sequence-level Exts, zero-length sequences sharing a start,
    and a NON-contiguous / out-of-order layout (general warp-per-chunk kernel instead of the tile kernel)."""
    rng = np.random.default_rng(21)
    # uniform, with seq_exts
    seqs = [random_dna(rng, 40) for _ in range(300)]
    sx = rng.integers(0, 256, size=len(seqs)).astype(np.uint8)
    run_both(D, ctx, orc, 31, orc.seqset_from_lists(seqs), 1, seq_exts=sx)
    run_both(D, ctx, orc, 33, orc.seqset_from_lists(seqs), 1, seq_exts=sx, stranded=True)
    # zero-length and shorter-than-K sequences between real ones
    seqs = [random_dna(rng, n) for n in (0, 35, 0, 0, 31, 5, 0, 90, 31, 0)]
    sx = rng.integers(0, 256, size=len(seqs)).astype(np.uint8)
    run_both(D, ctx, orc, 31, orc.seqset_from_lists(seqs), 1, seq_exts=sx)
    # non-contiguous: sequences are slices (with gaps, reversed order, one overlap) of one packed buffer
    big = random_dna(rng, 6000)
    words = orc.pack_bases(big)
    start = np.array([5000, 3100, 1200, 1190, 40, 0], dtype=np.uint64)
    length = np.array([1000, 700, 650, 31, 1000, 33], dtype=np.uint32)
    sx = rng.integers(0, 256, size=len(start)).astype(np.uint8)
    run_both(D, ctx, orc, 31, (words, start, length), 1, seq_exts=sx)
    run_both(D, ctx, orc, 31, (words, start, length), 1, stranded=True)


def test_uniform_upload_entry_point(D, ctx, orc):
    w, s, l = orc.synth_reads(2500, 1, orc.ERR_THR_NOISY)
    ss = D.SeqSet.upload_uniform(ctx, w, 2500, 150)
    table, _ = D.filter_kmers(ss, D.CountFilter(2), False, False, 4, k=31)
    ot = orc.filter_kmers(31, w, s, l, min_obs=2)
    assert_tables_equal(table.to_host(), ot)
    with pytest.raises(D.DbgError):
        D.SeqSet.upload_uniform(ctx, w, 2501, 150)


def test_pipelined_upload_matches_oracle(D, ctx, orc):
    """Asynchronous chunked upload (>= 32 MB of packed reads: 8 chunks on the copy stream, the partition kernel starts on
    the chunks that have arrived): table bit-identical to the oracle, through both entry points that use it."""
    import ctypes as C
    R = 900_000                                       # 4.2 M words: the smallest size that is actually chunked
    w, s, l = orc.synth_reads(R, 3, orc.ERR_THR_NOISY)
    ot = orc.filter_kmers(31, w, s, l, min_obs=2)
    ss = D.SeqSet.upload_uniform(ctx, w, R, 150, pipelined=True)
    table, _ = D.filter_kmers(ss, D.CountFilter(2), False, False, 4, k=31)
    assert_tables_equal(table.to_host(), ot)
    og = orc.compress_kmers(31, ot["lo"], ot["hi"], ot["exts"], ot["counts"])
    gh = C.c_void_p()
    L = ctx._L
    ctx.check(L.dbg_reads_to_graph_host_uniform(ctx._h, 31, C.c_void_p(w.ctypes.data), len(w), R, 150, None, 2, 0, D.SAT_ADD,
                                                None, C.byref(gh)))
    assert_graphs_equal(D.BaseGraph(ctx, gh).to_host(), og)


def test_pipelined_upload_direct_partition_and_its_fallback(D, orc):
    """A pipelined upload takes the direct partition sized from its FIRST chunk.  Reads in sequencer order: it holds
    (direct_partition == 1).  An input whose first chunk says nothing about the rest (an eighth of identical low-complexity reads,
    then random ones): regions overflow, the call falls back to staging, the result is still the oracle's, and the context keeps
    later pipelined inputs on the staging path."""
    c2 = D.Context(0)
    R = 900_000
    w, s, l = orc.synth_reads(R, 5, orc.ERR_THR_NOISY)
    ot = orc.filter_kmers(31, w, s, l, min_obs=2, threads=os.cpu_count() or 1)
    table, _ = D.filter_kmers(D.SeqSet.upload_uniform(c2, w, R, 150, pipelined=True), D.CountFilter(2), False, False, 4, k=31)
    assert_tables_equal(table.to_host(), ot)
    assert c2.stats()["direct_partition"] == 1
    # first eighth: the same read over and over (16-read period keeps the packed words periodic); rest: the synthetic reads
    w2 = w.copy()
    n8 = (R // 8 // 16) * 16
    period = 16 * 150 // 32                                   # 16 reads = 75 words
    w2[:n8 * 150 // 32] = np.tile(w[:period], n8 // 16)
    ot2 = orc.filter_kmers(31, w2, s, l, min_obs=2, threads=os.cpu_count() or 1)
    table2, _ = D.filter_kmers(D.SeqSet.upload_uniform(c2, w2, R, 150, pipelined=True), D.CountFilter(2), False, False, 4, k=31)
    assert_tables_equal(table2.to_host(), ot2)
    assert c2.stats()["direct_partition"] == 0                 # fell back
    table3, _ = D.filter_kmers(D.SeqSet.upload_uniform(c2, w, R, 150, pipelined=True), D.CountFilter(2), False, False, 4, k=31)
    assert_tables_equal(table3.to_host(), ot)
    assert c2.stats()["direct_partition"] == 0                 # remembered: staging from the start
    table4, _ = D.filter_kmers(D.SeqSet.upload_uniform(c2, w, R, 150), D.CountFilter(2), False, False, 4, k=31)
    assert c2.stats()["direct_partition"] == 1                 # resident inputs still sample the whole input
    assert_tables_equal(table4.to_host(), ot)


def test_from_ascii_ingest(D, ctx, orc):
    """dbg_seqset_from_ascii = DnaString::from_acgt_bytes + PackedDnaStringSet::add on the device: packed words, offsets,
    lengths and the invalid-character count identical to the oracle; the packed set then feeds filter_kmers."""
    rng = np.random.default_rng(41)
    alphabet = np.frombuffer(b"ACGTacgtNn-*X", np.uint8)
    pr = np.array([.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .005, .005, .01])
    seqs = [bytes(rng.choice(alphabet, size=n, p=pr / pr.sum())) for n in (0, 1, 31, 32, 33, 64, 150, 151, 0, 1000, 95, 4097, 7)]
    ow, ost, oln, obad = orc.from_acgt_bytes(seqs)
    ss = D.SeqSet.from_ascii(ctx, seqs)
    w, st, ln = ss.copy_out()
    assert ss.n_invalid == obad and obad > 0
    assert np.array_equal(w, ow) and np.array_equal(st, ost) and np.array_equal(ln, oln)
    table, _ = D.filter_kmers(ss, D.CountFilter(1), False, False, 4, k=31)
    assert_tables_equal(table.to_host(), orc.filter_kmers(31, ow, ost, oln, min_obs=1))
    # uniform reads (the fast tile layout), all valid
    reads = [bytes(rng.choice(alphabet[:4], size=150)) for _ in range(500)]
    ss2 = D.SeqSet.from_ascii(ctx, reads)
    ow2, ost2, oln2, _ = orc.from_acgt_bytes(reads)
    assert ss2.n_invalid == 0 and np.array_equal(ss2.copy_out()[0], ow2)
    empty = D.SeqSet.from_ascii(ctx, [])
    assert len(empty.copy_out()[1]) == 0


def test_remove_censored_exts(D, ctx, orc):
    """dbg_remove_censored_exts vs the oracle's filter::remove_censored_exts / _sharded (filter.rs:238-306), then the pruned
    table through compress_kmers: Exts bytes and BaseGraph identical, both key widths, sequence-level Exts make the two
    variants differ."""
    rng = np.random.default_rng(9)
    w, st, ln = orc.synth_reads(2500, 1, orc.ERR_THR_NOISY)
    for k, stranded in ((31, False), (32, True), (63, False), (33, True)):
        sx = rng.integers(0, 256, size=len(st)).astype(np.uint8)
        ot = orc.filter_kmers(k, w, st, ln, seq_exts=sx, min_obs=2, stranded=stranded, report_all=True)
        for sharded in (False, True):
            table, _ = D.filter_kmers((w, st, ln, sx), D.CountFilter(2), stranded, True, 4, k=k, ctx=ctx)
            (D.remove_censored_exts_sharded if sharded else D.remove_censored_exts)(stranded, table)
            oe = orc.remove_censored_exts(k, ot, stranded=stranded, sharded=sharded)
            t = table.to_host()
            assert np.array_equal(t["exts"], oe) and np.array_equal(t["lo"], ot["lo"]) and np.array_equal(t["counts"], ot["counts"])
            g = D.compress_kmers_with_hash(stranded, D.SimpleCompress(D.SAT_ADD), table).to_host()
            og = orc.compress_kmers(k, ot["lo"], ot["hi"], oe, ot["counts"], stranded=stranded)
            assert og["error"] == 0
            assert_graphs_equal(g, og)
    t0, _ = D.filter_kmers((w, st, ln), D.CountFilter(2), False, False, 4, k=31, ctx=ctx)
    with pytest.raises(D.DbgError):
        D.remove_censored_exts_sharded(False, t0)     # no all_kmers in this table


def test_graph_edges(D, ctx, orc):
    """dbg_graph_edges (BaseGraph::finish + find_edges for every node and side, graph.rs:116-142, 223-291) vs the oracle:
    identical target / flag arrays; K = 31 / 32 / 63 / 64, stranded and not, cycles and hairpins at small K."""
    rng = np.random.default_rng(15)
    w, st, ln = orc.synth_reads(2500, 1, orc.ERR_THR_NOISY)
    cases = [(31, False, (w, st, ln), 2), (63, False, (w, st, ln), 2), (32, True, (w, st, ln), 2), (31, False, orc.synth_reads(1500, 1, 0), 1)]
    for k in (5, 6, 7, 64):
        contigs = small_k_contigs(rng, alphabet=2 if k < 8 else 4, lo=max(k, 8), hi=max(4 * k, 60))
        contigs = [c for c in contigs if len(c) >= k]
        cases.append((k, bool(k & 1), orc.seqset_from_lists(contigs), 1))
    for k, stranded, ss, mo in cases:
        table, _ = D.filter_kmers(ss, D.CountFilter(mo), stranded, False, 4, k=k, ctx=ctx)
        if k == 31 and mo == 2:
            D.remove_censored_exts(stranded, table)
        graph = D.compress_kmers_with_hash(stranded, D.SimpleCompress(D.SAT_ADD), table)
        g = graph.to_host()
        target, flags = graph.edges()
        ot, of, _ = orc.graph_edges(k, g, stranded=stranded)
        assert np.array_equal(target, ot) and np.array_equal(flags, of), (k, stranded)
    empty, _ = D.filter_kmers((np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint32)), D.CountFilter(1), False, False, 4, k=31, ctx=ctx)
    assert D.compress_kmers_with_hash(False, D.SimpleCompress(), empty).edges()[0].shape == (0, 2, 4)


def test_graph_fix_exts(D, ctx, orc):
    """dbg_graph_fix_exts (DebruijnGraph::fix_exts, graph.rs:337-377) vs the oracle, with and without valid_nodes, K = 31 / 63."""
    w, st, ln = orc.synth_reads(2500, 1, orc.ERR_THR_NOISY)
    for k in (31, 63):
        table, _ = D.filter_kmers((w, st, ln), D.CountFilter(2), False, False, 4, k=k, ctx=ctx)
        for use_mask in (False, True):
            graph = D.compress_kmers_with_hash(False, D.SimpleCompress(D.SAT_ADD), table)
            g = graph.to_host()
            vn = (np.arange(g["n_nodes"]) % 3 != 0) if use_mask else None
            oe = orc.graph_fix_exts(k, g, valid_nodes=vn)
            graph.fix_exts(vn)
            g2 = graph.to_host()
            assert np.array_equal(g2["exts"], oe) and (oe != g["exts"]).any()
            assert np.array_equal(g2["words"], g["words"]) and np.array_equal(g2["data"], g["data"])


def test_scmap_compress(D, ctx, orc):
    """ScmapCompress (compression.rs:66-98, join_test = data equality) through compress_kmers: BaseGraph identical to the
    oracle, fast path (noisy reads) and general path (clean reads: long unitigs broken only where the data changes)."""
    for noisy, mo in ((True, 2), (False, 1)):
        w, st, ln = orc.synth_reads(2000, 1, orc.ERR_THR_NOISY if noisy else 0)
        ot = orc.filter_kmers(31, w, st, ln, min_obs=mo)
        counts = (ot["counts"] % 3).astype(np.uint16)
        g = D.compress_kmers(False, D.ScmapCompress(), (ot["lo"], None, ot["exts"], counts), k=31, ctx=ctx).to_host()
        og = orc.compress_kmers(31, ot["lo"], ot["hi"], ot["exts"], counts, reduce_op=orc.SCMAP)
        assert og["error"] == 0
        assert_graphs_equal(g, og)
    ot63 = orc.filter_kmers(63, w, st, ln, min_obs=1)
    c63 = (ot63["counts"] % 2).astype(np.uint16)
    g = D.compress_kmers(False, D.ScmapCompress(), (ot63["lo"], ot63["hi"], ot63["exts"], c63), k=63, ctx=ctx).to_host()
    assert_graphs_equal(g, orc.compress_kmers(63, ot63["lo"], ot63["hi"], ot63["exts"], c63, reduce_op=orc.SCMAP))


def test_write_gfa(D, ctx, orc):
    """BaseGraph.write_gfa (DebruijnGraph::write_gfa, graph.rs:538-614; edges from dbg_graph_edges) = the oracle's text."""
    import io
    w, st, ln = orc.synth_reads(1200, 1, orc.ERR_THR_NOISY)
    for k, stranded in ((31, False), (32, True)):
        table, _ = D.filter_kmers((w, st, ln), D.CountFilter(2), stranded, False, 4, k=k, ctx=ctx)
        D.remove_censored_exts(stranded, table)
        graph = D.compress_kmers_with_hash(stranded, D.SimpleCompress(D.SAT_ADD), table)
        buf = io.StringIO()
        graph.write_gfa(buf)
        assert buf.getvalue() == orc.write_gfa(k, graph.to_host(), stranded=stranded)
        assert buf.getvalue().count("\nL\t") > 0


def test_count_saturation(D, ctx, orc):
    """filter.rs:57 counts saturate at 65535; compression.rs:495 single-k-mer node keeps raw data."""
    seq = enc("ACGTTGCATGCATCGATCGATCGTAGCTAGA")
    for op in (0, 2):
        t, g = run_both(D, ctx, orc, 31, orc.seqset_from_lists([seq] * 70000), 1, reduce_op=op)
        assert int(t["counts"][0]) == 65535 and int(g["data"][0]) == 65535


def test_long_sequences_chunked(D, ctx, orc):
    """Sequences far longer than one work item (chunked scan) + low-complexity stretches."""
    rng = np.random.default_rng(11)
    seqs = [random_dna(rng, 20000), np.zeros(3000, np.uint8), np.tile(np.array([0, 3], np.uint8), 2000),
            random_dna(rng, 1037)]
    run_both(D, ctx, orc, 31, orc.seqset_from_lists(seqs), 1, report_all=True)
    run_both(D, ctx, orc, 63, orc.seqset_from_lists(seqs), 1)


def test_long_unitigs_and_long_cycles(D, ctx, orc):
    """Components longer than the end-walk limit (1024 k-mers) go through the contracted-graph list ranking:
    long paths, a long cycle (circular sequence), a cycle shorter than the splitter spacing, mixed with short ones."""
    rng = np.random.default_rng(5)
    k = 31
    circ = random_dna(rng, 5000)
    circular = np.concatenate([circ, circ[:k + 5]])           # wraps past the junction: a true 5000-k-mer cycle
    small = random_dna(rng, 40)
    small_circ = np.concatenate([small, small[:k + 5]])       # 40-cycle: may draw no splitter
    seqs = [random_dna(rng, 9000), circular, small_circ, random_dna(rng, 200), random_dna(rng, 3000)]
    t, g = run_both(D, ctx, orc, k, orc.seqset_from_lists(seqs), 1)
    assert ctx.stats()["n_cycle_kmers"] >= 5040
    for stranded in (False, True):
        run_both(D, ctx, orc, 21, orc.seqset_from_lists([s[:4000] for s in seqs]), 1, stranded=stranded)
    run_both(D, ctx, orc, 63, orc.seqset_from_lists(seqs), 1)


def test_bucket_split_path(D, ctx, orc):
    """Force shared-memory table overflows (one huge bucket) so the hash-class splitting runs."""
    c2 = D.Context(0)
    c2.set_param("bucket_occ", 1 << 30)  # everything lands in a single MSP bucket
    ss = orc.synth_reads(3000, 1, orc.ERR_THR_NOISY)
    for k in (31, 63):
        t, g = run_both(D, c2, orc, k, ss, 2, report_all=True)
        # ~20 000 (27 000) distinct k-mers in one bucket against an 8192-slot table: the splitting must have run
        assert c2.stats()["n_buckets"] == 1 and c2.stats()["n_bucket_splits"] > 0, c2.stats()
    c2.close()


def test_sort_prefix_runs_and_fallback(D, ctx, orc):
    """The table sort works on the top digits plus a run fix-up; many k-mers sharing a long prefix (low-complexity
    reads that differ only near their 3' end) force the all-digit fallback.  Same bits either way."""
    rng = np.random.default_rng(8)
    base = np.zeros(140, np.uint8)
    seqs = []
    for i in range(400):   # A...A + random 10-base tail: thousands of k-mers with an identical 20+ base prefix
        seqs.append(np.concatenate([base[: 100 + i % 30], random_dna(rng, 12)]))
    for k in (31, 45):
        run_both(D, ctx, orc, k, orc.seqset_from_lists(seqs), 1, report_all=True)


def test_multi_pass_planner(D, ctx, orc):
    """memory_size semantics of filter_kmers (src/filter.rs:151-168): a small scratch budget splits the MSP buckets
    into ranges and re-scans the reads once per range; the number of passes never changes the result."""
    c2 = D.Context(0)
    c2.set_param("mem_budget_bytes", 1)
    ss = orc.synth_reads(3000, 1, orc.ERR_THR_NOISY)
    run_both(D, c2, orc, 31, ss, 2, report_all=True)
    assert c2.stats()["n_passes"] > 1
    run_both(D, c2, orc, 63, ss, 2)
    rng = np.random.default_rng(3)
    big = random_dna(rng, 6000)   # non-contiguous layout: the general partition kernel also honours the bucket range
    run_both(D, c2, orc, 31, (orc.pack_bases(big), np.array([3000, 100], np.uint64), np.array([2500, 1700], np.uint32)), 1)
    c2.close()
    assert ctx.stats()["n_passes"] in (0, 1)


def test_valid_buffer_retry(D, ctx, orc):
    """The valid-k-mer buffer is sized by an estimate; force it far too small so the exact-bound retry runs
    (second attempt re-uses the already deduplicated records)."""
    c2 = D.Context(0)
    c2.set_param("valid_est_div", 1000)
    ss = orc.synth_reads(3000, 1, 0)          # clean: every distinct k-mer is valid
    run_both(D, c2, orc, 31, ss, 1, report_all=True)
    run_both(D, c2, orc, 63, ss, 1)
    c2.close()


def test_direct_partition_and_staging_agree(D, ctx, orc):
    """The partition stage writes records straight into per-bucket regions sized by a sampling pass (large contiguous
    inputs) or stages + scatters them (everything else, and the fallback when a region overflows).  Same bits: direct
    forced on small inputs (few tiles: poor estimates, overflow -> fallback exercised too), direct off, both key widths."""
    ss = orc.synth_reads(6000, 1, orc.ERR_THR_NOISY)
    c2 = D.Context(0)
    c2.set_param("direct_min_tiles", 1)
    t, _ = run_both(D, c2, orc, 31, ss, 2, report_all=True)
    st = c2.stats()
    run_both(D, c2, orc, 63, ss, 2)
    run_both(D, c2, orc, 31, orc.synth_reads(300, 1, 0), 1)          # 11 tiles: the sample sees almost nothing
    seq = enc("ACGTTGCATGCATCGATCGATCGTAGCTAGAGGATCCA")
    run_both(D, c2, orc, 31, orc.seqset_from_lists([seq] * 3000 + [random_dna(np.random.default_rng(5), 4000)]), 1)   # one heavy bucket
    c2.set_param("direct_partition", 0)
    t2, _ = run_both(D, c2, orc, 31, ss, 2, report_all=True)
    assert c2.stats()["direct_partition"] == 0
    assert_tables_equal(t, t2)
    c2.close()
    assert st["n_records"] > 0


def test_general_compress_path_forced(D, ctx, orc):
    """compress_kmers has a fast path (all components are short paths: one walker per node) and a general path
    (per-k-mer rank + emit; long unitigs, cycles).  Same bits when the general path is forced on data the fast
    path would take, for every reduce op and both key widths."""
    c2 = D.Context(0)
    c2.set_param("fast_compress", 0)
    ss = orc.synth_reads(3000, 1, orc.ERR_THR_NOISY)
    for op in (0, 1, 2, 3):
        run_both(D, c2, orc, 31, ss, 2, reduce_op=op)
    run_both(D, c2, orc, 63, ss, 2)
    run_both(D, c2, orc, 32, ss, 2, stranded=True)
    c2.close()


def test_fast_compress_node_word_boundaries(D, ctx, orc):
    """The per-node writer of the fast path: nodes of every length 1..70 k-mers (K = 31, 33, 64) packed back to
    back, so first / last words are shared between neighbours at every bit offset."""
    rng = np.random.default_rng(77)
    for k in (31, 33, 64, 5):
        seqs = [random_dna(rng, k + n) for n in range(0, 70)]
        run_both(D, ctx, orc, k, orc.seqset_from_lists(seqs), 1)
        run_both(D, ctx, orc, k, orc.seqset_from_lists(seqs), 1, stranded=True, reduce_op=3)


def test_record_dedup_off(D, ctx, orc):
    """The per-bucket super-k-mer deduplication is an optimisation only: same bits with it switched off."""
    c2 = D.Context(0)
    c2.set_param("dedup", 0)
    ss = orc.synth_reads(3000, 1, orc.ERR_THR_NOISY)
    run_both(D, c2, orc, 31, ss, 2, report_all=True)
    seq = enc("ACGTTGCATGCATCGATCGATCGTAGCTAGA")
    run_both(D, c2, orc, 31, orc.seqset_from_lists([seq] * 70000), 1)
    c2.set_param("dedup", 2)   # also deduplicate the 32-byte records of two-word keys (off by default: ~neutral)
    run_both(D, c2, orc, 63, ss, 2, report_all=True)
    run_both(D, c2, orc, 40, orc.seqset_from_lists([np.tile(seq, 3)] * 300), 1)
    c2.close()


@pytest.mark.parametrize("k,p", [(31, 6), (31, 8), (35, 5), (63, 12), (16, 5)])
def test_msp_kmer_buckets_match_scanner(D, ctx, orc, k, p):
    """dbg_msp_kmer_buckets vs the oracle's Scanner::scan (src/msp.rs:207-276): every k-mer of every interval
    carries the interval's bucket = min_rc(minimizer), identity permutation, rc=true and rc=false."""
    rng = np.random.default_rng(k * 10 + p)
    seqs = [random_dna(rng, int(rng.integers(k, 5 * k))) for _ in range(12)] + [np.zeros(3 * k, np.uint8), random_dna(rng, k - 1)]
    ss = D.SeqSet.upload(ctx, *orc.seqset_from_lists(seqs))
    for stranded in (False, True):
        got = D.msp_kmer_buckets(ss, k, p, stranded=stranded)
        exp = []
        for sq in seqs:
            if len(sq) < k:
                continue
            iv = orc.msp_scan(k, p, sq, rc=not stranded)
            per = np.zeros(len(sq) - k + 1, np.uint32)
            for st, ln, b in zip(iv["start"], iv["len"], iv["bucket"]):
                per[int(st):int(st) + int(ln) - k + 1] = int(b)
            exp.append(per)
        assert np.array_equal(got, np.concatenate(exp))


@pytest.mark.parametrize("k,p", [(31, 6), (31, 8), (35, 5), (63, 12), (16, 5)])
def test_msp_sequence_intervals_match_scanner(D, ctx, orc, k, p):
    """dbg_msp_sequence vs the oracle's Scanner::scan + msp_sequence (src/msp.rs:207-324): same intervals (start, len), buckets
    and boundary Exts in scan order — identity permutation and a random one, rc = true / false, low-complexity sequences with
    tied minimizers (the history-dependent case), sequences shorter than k."""
    rng = np.random.default_rng(k * 13 + p)
    seqs = [random_dna(rng, int(rng.integers(k, 6 * k))) for _ in range(20)] + \
           [np.zeros(3 * k, np.uint8), np.tile(np.array([0, 1], np.uint8), 2 * k), random_dna(rng, k - 1), random_dna(rng, k),
            rng.integers(0, 2, size=5 * k).astype(np.uint8)]
    ss = D.SeqSet.upload(ctx, *orc.seqset_from_lists(seqs))
    perms = [None] + ([rng.permutation(4 ** p).astype(np.uint32)] if p <= 8 else [])
    for perm in perms:
        for rc in (True, False):
            got = D.msp_sequence(ss, k, p, permutation=perm, rc=rc)
            exp = {f: [] for f in ("seq", "start", "len", "bucket", "exts")}
            for si, sq in enumerate(seqs):
                if len(sq) < k:
                    continue
                iv = orc.msp_scan(k, p, sq, perm=None if perm is None else perm.astype(np.uint64), rc=rc)
                exp["seq"] += [si] * len(iv["start"])
                for f in ("start", "len", "bucket", "exts"):
                    exp[f] += [int(x) for x in iv[f]]
            for f in exp:
                assert np.array_equal(got[f].astype(np.int64), np.array(exp[f], np.int64)), (f, perm is None, rc)
    # uniform-length reads through the fixed-length layout
    w, s_, l = orc.synth_reads(500, 1, orc.ERR_THR_NOISY)
    got = D.msp_sequence(D.SeqSet.upload(ctx, w, s_, l), 31, 8)
    n_exp = sum(len(orc.msp_scan(31, 8, orc.unpack_bases(w, 150 * i, 150))["start"]) for i in range(500))
    assert len(got["start"]) == n_exp and int(got["len"].max()) <= 2 * 31 - 8


def test_from_ascii_hashn_on_device(D, ctx, orc):
    """dbg_seqset_from_ascii_hashn vs the oracle's restatement of DnaString::from_acgt_bytes_hashn (src/dna_string.rs:254-278)."""
    rng = np.random.default_rng(17)
    alpha = np.frombuffer(b"ACGTacgtNnRYKMSW-.", np.uint8)
    seqs = [bytes(rng.choice(alpha, size=n)) for n in (0, 1, 31, 32, 33, 150, 150, 4097, 64)]
    names = [("read:%d/%d" % (i, i * 7919)).encode() + b"x" * (i % 9) for i in range(len(seqs))]
    ss = D.SeqSet.from_ascii_hashn(ctx, seqs, names)
    ow, ost, oln, obad = orc.from_acgt_bytes_hashn(seqs, names)
    w, s_, l = ss.copy_out()
    assert np.array_equal(w, ow) and np.array_equal(s_, ost) and np.array_equal(l, oln) and ss.n_invalid == obad
    # and the graph built from it equals the oracle's on the same packed bases
    t, g = run_both(D, ctx, orc, 15, (ow, ost, oln), 1)
    assert g["n_nodes"] > 0


def test_fastq_feeding(D, ctx, orc, tmp_path):
    """FASTQ -> device sequence set (SeqSet.from_fastq: host-side record split, device-side packing with the hashn policy) equals the
    oracle's from_acgt_bytes_hashn on the same records; CRLF line ends, an empty read and a path argument included."""
    rng = np.random.default_rng(23)
    alpha = np.frombuffer(b"ACGTACGTACGTNn", np.uint8)
    seqs = [bytes(rng.choice(alpha, size=n)) for n in (150, 0, 151, 76, 33)]
    names = [b"r%d" % i for i in range(len(seqs))]
    fq = b"".join(b"@" + n + b" extra field\n" + s + (b"\r\n" if i == 2 else b"\n") + b"+\n" + b"I" * len(s) + b"\n"
                  for i, (n, s) in enumerate(zip(names, seqs)))
    path = tmp_path / "reads.fastq"
    path.write_bytes(fq)
    ow, ost, oln, obad = orc.from_acgt_bytes_hashn(seqs, names)
    for src in (fq, str(path)):
        ss = D.SeqSet.from_fastq(ctx, src)
        w, s_, l = ss.copy_out()
        assert np.array_equal(w, ow) and np.array_equal(s_, ost) and np.array_equal(l, oln) and ss.n_invalid == obad
    assert D.SeqSet.parse_fastq(fq) == (names, seqs)
    with pytest.raises(ValueError):
        D.SeqSet.parse_fastq(fq[:-20] + b"\nxx")


def test_bincode_image_round_trip(D, ctx, orc):
    """dbg_graph_serialize / dbg_graph_deserialize: the bincode image of BaseGraph<K, u16> equals the oracle's (hand-checked layout,
    tests/test_oracle.py::test_bincode_image_layout) and reads back to the same arrays."""
    for k, stranded in ((31, False), (63, False), (32, True)):
        ss = orc.synth_reads(2000, 1, orc.ERR_THR_NOISY)
        table, _ = D.filter_kmers(ss, D.CountFilter(2), stranded, False, 4, k=k, ctx=ctx)
        g = D.compress_kmers_with_hash(stranded, D.SimpleCompress(D.SAT_ADD), table)
        img = g.to_bincode()
        h = g.to_host()
        assert img == orc.graph_to_bincode(h)
        g2 = D.BaseGraph.from_bincode(img, k, ctx=ctx)
        assert_graphs_equal(g2.to_host(), h)
        assert g2.stranded == stranded
    e = np.zeros(0, np.uint64)
    table, _ = D.filter_kmers((e, e, np.zeros(0, np.uint32)), D.CountFilter(1), False, False, 4, k=31, ctx=ctx)
    g0 = D.compress_kmers_with_hash(False, D.SimpleCompress(), table)
    assert len(D.BaseGraph.from_bincode(g0.to_bincode(), 31, ctx=ctx)) == 0
    with pytest.raises(D.DbgError):
        D.BaseGraph.from_bincode(img[:-5], 31, ctx=ctx)


@pytest.mark.parametrize("k,stranded", [(31, False), (63, False), (12, True), (32, False)])
def test_count_filter_set(D, ctx, orc, k, stranded):
    """filter_kmers with CountFilterSet<u8> (src/filter.rs:68-101): valid k-mers, Exts and per-k-mer label sets against a
    brute-force restatement — shared k-mers across labels, a label with a single sequence, thresholds 1..3, sequence-level Exts."""
    rng = np.random.default_rng(500 + k + stranded)
    base = [random_dna(rng, int(rng.integers(k + 5, 6 * k))) for _ in range(6)]
    seqs, labels = [], []
    for rep in range(14):   # overlapping fragments of the same contigs under different labels
        c = base[int(rng.integers(0, len(base)))]
        a = int(rng.integers(0, max(1, len(c) - k - 2)))
        seqs.append(c[a:a + int(rng.integers(k, len(c) - a + 1))])
        labels.append(int(rng.choice([0, 1, 5, 63])))
    seqs.append(random_dna(rng, k - 1)); labels.append(7)          # shorter than k: contributes nothing
    seqs.append(base[0]); labels.append(17)                        # a label used once
    sx = rng.integers(0, 256, size=len(seqs)).astype(np.uint8)
    w, st, ln = orc.seqset_from_lists(seqs)
    for mo in (1, 2, 3):
        table, _ = D.filter_kmers((w, st, ln, sx), D.CountFilterSet(mo), stranded, False, 4, k=k, ctx=ctx, labels=labels)
        t = table.to_host()
        keys, exts, sets, nobs = brute_filter_colorset(orc, k, seqs, labels, mo, stranded=stranded, seq_exts=sx)
        got = [int(lo) | ((int(hi) << 64) if k > 32 else 0) for lo, hi in zip(t["lo"], t["hi"] if k > 32 else t["lo"])]
        assert got == keys
        assert [int(x) for x in t["exts"]] == exts and [int(x) for x in t["counts"]] == nobs
        assert table.colorsets() == sets
        # the table feeds compress_kmers like any other (ScmapCompress joins k-mers of equal data; here: the counts)
        assert len(D.compress_kmers_with_hash(stranded, D.SimpleCompress(D.MAX), table)) > 0


def _run_tool(args, timeout=900):
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    return subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout, cwd=root)


@pytest.mark.parametrize("ranks", [2, 3])
def test_multi_rank_path_on_one_gpu(ranks):
    """The multi-GPU path of the library (dbg_multi_reads_to_graph: MSP-bucket-sharded filter_kmers, bucket-sharded table,
    remote neighbour queries, cross-rank unitig walks, path records by seed-key range) with several ranks on ONE device over
    the local transport: the per-rank runs of nodes, concatenated, are the oracle's BaseGraph bit for bit — K=31 noisy / clean
    (replicated fallback), K=63, stranded + max, many buckets (p >= 13), ScmapCompress.  tools/multi_check.py."""
    r = _run_tool([os.path.join("tools", "multi_check.py"), "--local", str(ranks), "--reads", "30000"])
    assert r.returncode == 0 and "MULTI OK" in r.stdout and r.stdout.count("BIT-EXACT") == 6 and "MISMATCH" not in r.stdout, \
        r.stdout[-3000:] + r.stderr[-3000:]


def test_multi_rank_path_2m_reads_on_one_gpu():
    """Same at 2 * 10^6 reads in total (N = 2.4 * 10^8 k-mers, 2^13+ buckets, ~7 * 10^6 valid k-mers over 2 ranks)."""
    r = _run_tool([os.path.join("tools", "multi_check.py"), "--local", "2", "--reads", "2000", "--big-reads", "2000000"])
    assert r.returncode == 0 and "MULTI OK" in r.stdout and "R=2000000" in r.stdout and "MISMATCH" not in r.stdout, \
        r.stdout[-3000:] + r.stderr[-3000:]


def test_multi_gpu_nccl():
    """One process per GPU under torchrun (NCCL all-to-alls + CUDA IPC peer windows inside the library): bit-exact vs the oracle
    on the union of the ranks' reads, incl. a 2 * 10^6-read configuration.  Needs >= 2 GPUs (skipped otherwise)."""
    import sys

    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    n = 2 if n < 4 else 4
    r = _run_tool(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                   "--master-port", "29533", os.path.join("tools", "multi_check.py"), "--reads", "40000", "--big-reads", "2000000"])
    assert r.returncode == 0 and "MULTI OK" in r.stdout and r.stdout.count("BIT-EXACT") == 7 and "MISMATCH" not in r.stdout, \
        r.stdout[-3000:] + r.stderr[-3000:]


def _full_size_bit_exact(D, orc, k, also_staged):
    """BASELINE.json configs[1] / configs[2] at their FULL size (10M x 150bp synth-v1 noisy reads): the whole k-mer table
    and every BaseGraph array compared with the oracle (multi-threaded filter stage; the greedy walk is serial)."""
    import os
    R = 10_000_000
    ctx = D.Context(0)
    ss = D.SeqSet.synth(ctx, R, 1, orc.ERR_THR_NOISY)
    table, graph = D.reads_to_graph(ss, D.CountFilter(2), D.SimpleCompress(D.SAT_ADD), k=k, keep_table=True)
    st = ctx.stats()
    t, g = table.to_host(), graph.to_host()
    table.free(); graph.free()
    if also_staged:   # the staged partition (what pipelined uploads and small inputs use) must give the same table
        assert st["direct_partition"] == 1
        ctx.set_param("direct_partition", 0)
        t2 = D.filter_kmers(ss, D.CountFilter(2), False, False, 0, k=k)[0]
        assert ctx.stats()["direct_partition"] == 0
        h2 = t2.to_host()
        t2.free()
        assert_tables_equal(t, h2)
        del h2
    ss.free()
    ctx.close()
    w, s_, l = orc.synth_reads(R, 1, orc.ERR_THR_NOISY)
    ot = orc.filter_kmers(k, w, s_, l, min_obs=2, threads=os.cpu_count() or 1)
    del w, s_, l
    assert t["n_input"] == ot["n_input"] == R * (150 - k + 1)
    assert_tables_equal(t, ot)
    og = orc.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"])
    assert og["error"] == 0
    assert_graphs_equal(g, og)
    V, M = len(t["lo"]), g["n_nodes"]
    assert int(g["length"].astype(np.uint64).sum()) == V + M * (k - 1) == g["n_bases"]


def test_full_size_config1_bit_exact(D, orc):
    """configs[1]: K=31, 10^7 reads, N = 1.2 * 10^9 k-mers — table and graph byte-identical to the oracle."""
    _full_size_bit_exact(D, orc, 31, also_staged=True)


def test_full_size_config2_k63_bit_exact(D, orc):
    """configs[2]: K=63 (two-u64 keys), 10^7 reads, N = 8.8 * 10^8 k-mers — table and graph byte-identical to the oracle."""
    _full_size_bit_exact(D, orc, 63, also_staged=False)


def _canon_form(orc, g):
    """multiset of (min(seq, rc seq), exts passed through Exts::rc when the rc was chosen, data)."""
    L = orc.lib()
    out = []
    for i in range(g["n_nodes"]):
        b = orc.unpack_bases(g["words"], int(g["start"][i]), int(g["length"][i]))
        r = (3 - b[::-1]).astype(np.uint8)
        e = int(g["exts"][i])
        if bytes(r) < bytes(b):
            b, e = r, L.orc_exts_rc(e)
        out.append((bytes(b), e, int(g["data"][i])))
    return sorted(out)


@pytest.mark.parametrize("k", [31, 32, 63])
def test_canonical_form_invariant_to_seed_order(D, ctx, orc, k):
    """The reference seeds unitigs in boomphf slot order (src/compression.rs:574-580), which this repo cannot reproduce
    (third-party hash, un-pinned).  The boomphf-proof statement: the CANONICAL form of the GPU graph equals the canonical
    form of the greedy walk run under ANY seed order — checked against 5 shuffled orders on acyclic data (only cyclic
    components depend on the order: their break-point moves)."""
    rng = np.random.default_rng(900 + k)
    contigs = [c for c in random_contigs(rng) if len(c) >= k]
    ss = orc.seqset_from_lists(contigs + contigs)
    table, _ = D.filter_kmers(ss, D.CountFilter(2), False, False, 4, k=k, ctx=ctx)
    g = D.compress_kmers_with_hash(False, D.SimpleCompress(D.SAT_ADD), table).to_host()
    t = table.to_host()
    mine = _canon_form(orc, g)
    for it in range(5):
        perm = rng.permutation(len(t["lo"])).astype(np.uint32)
        og = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], seed_order=perm)
        assert og["error"] == 0 and og["n_nodes"] == g["n_nodes"]
        assert _canon_form(orc, og) == mine, f"canonical graph differs under shuffled seed order {it}"
    # noisy synthetic reads as well (censored k-mers leave dangling Exts: many short unitigs)
    ss = orc.synth_reads(3000, 1, orc.ERR_THR_NOISY)
    table, _ = D.filter_kmers(ss, D.CountFilter(2), False, False, 4, k=k, ctx=ctx)
    g = D.compress_kmers_with_hash(False, D.SimpleCompress(D.SAT_ADD), table).to_host()
    t = table.to_host()
    mine = _canon_form(orc, g)
    for it in range(5):
        perm = rng.permutation(len(t["lo"])).astype(np.uint32)
        og = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], seed_order=perm)
        assert _canon_form(orc, og) == mine


def test_is_compressed_on_device(D, ctx, orc):
    """DebruijnGraph::is_compressed (src/graph.rs:296-334) as a device kernel: None after compress_kmers on all-valid tables
    (the property src/test.rs:248-254 asserts), and the same first collapsible pair as the oracle's restatement on graphs that
    are NOT compressed (censored k-mers leave dangling Exts that end unitigs early; is_compressed ignores missing links)."""
    rng = np.random.default_rng(21)
    for k in (31, 32, 63):
        contigs = [c for c in random_contigs(rng) if len(c) >= k]
        table, _ = D.filter_kmers(orc.seqset_from_lists(contigs + contigs), D.CountFilter(2), False, False, 4, k=k, ctx=ctx)
        g = D.compress_kmers_with_hash(False, D.SimpleCompress(D.SAT_ADD), table)
        assert g.is_compressed() is None
        assert orc.graph_edges(k, g.to_host())[2] is None
    for k, R in ((31, 3000), (63, 2000), (31, 20000)):
        table, _ = D.filter_kmers(orc.synth_reads(R, 1, orc.ERR_THR_NOISY), D.CountFilter(2), False, False, 4, k=k, ctx=ctx)
        g = D.compress_kmers_with_hash(False, D.SimpleCompress(D.SAT_ADD), table)
        want = orc.graph_edges(k, g.to_host())[2]
        assert want is not None and g.is_compressed() == want
        # ScmapCompress: the data of both nodes must agree (join_test) — checked against a host evaluation of the same rule
        got = g.is_compressed(D.ScmapCompress())
        tgt, flg = g.edges()
        h = g.to_host()
        exp = None
        for i in range(h["n_nodes"]):
            for d in (0, 1):
                e1 = [(int(tgt[i, d, b]), int(flg[i, d, b]) & 1) for b in range(4) if tgt[i, d, b] != 0xffffffff]
                if len(e1) != 1:
                    continue
                nx, rd = e1[0]
                e2 = [b for b in range(4) if tgt[nx, rd, b] != 0xffffffff]
                if len(e2) == 1 and nx != i and h["data"][i] == h["data"][nx]:
                    exp = (i, nx)
                    break
            if exp:
                break
        assert got == exp


def test_compress_kmers_slice_variant(D, ctx, orc):
    """compression::compress_kmers (src/compression.rs:598-615): unordered (k-mer, (exts, data)) slice."""
    ss = orc.synth_reads(1500, 1, orc.ERR_THR_NOISY)
    ot = orc.filter_kmers(31, *ss, min_obs=2)
    perm = np.random.default_rng(0).permutation(len(ot["lo"]))
    g = D.compress_kmers(False, D.SimpleCompress(D.SAT_ADD), (ot["lo"][perm], None, ot["exts"][perm], ot["counts"][perm]),
                         k=31, ctx=ctx).to_host()
    og = orc.compress_kmers(31, ot["lo"], ot["hi"], ot["exts"], ot["counts"])
    assert_graphs_equal(g, og)


def test_inconsistent_exts_is_an_error(D, ctx, orc):
    """src/compression.rs:428-434 panic!("unreachable") -> DBG_E_INCONSISTENT_EXTS, no crash."""
    seq = enc("ACGTTGCATGCATCGATCGATCGTAGCTAGAC")  # two 31-mers
    ot = orc.filter_kmers(31, *orc.seqset_from_lists([seq]), min_obs=1, stranded=True)
    ex = ot["exts"].copy()
    ex[1] = 0  # drop the second k-mer's extension back to the first
    with pytest.raises(D.DbgError) as ei:
        D.compress_kmers(True, D.SimpleCompress(), (ot["lo"], None, ex, ot["counts"]), k=31, ctx=ctx)
    assert ei.value.status == 4


def test_fused_path_and_size_independent_properties(D, ctx, orc):
    """Larger device-generated input (2*10^5 reads, N = 2.4*10^7): checked through size-independent
    properties — ascending unique keys, checksums equal to the oracle's, node k-mers partition the
    valid set, sum of node lengths = V + M(K-1)."""
    R, k = 200000, 31
    ss = D.SeqSet.synth(ctx, R, 1, orc.ERR_THR_NOISY)
    table, graph = D.reads_to_graph(ss, D.CountFilter(2), D.SimpleCompress(D.SAT_ADD), k=k, keep_table=True)
    t, g = table.to_host(), graph.to_host()
    assert np.all(t["lo"][1:] > t["lo"][:-1])
    V, M = len(t["lo"]), g["n_nodes"]
    assert int(g["length"].astype(np.uint64).sum()) == V + M * (k - 1) == g["n_bases"]
    ot = orc.filter_kmers(k, *orc.synth_reads(R, 1, orc.ERR_THR_NOISY), min_obs=2)
    assert orc.xor_valid(t) == orc.xor_valid(ot) and orc.mix_valid(t) == orc.mix_valid(ot)
    og = orc.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"])
    assert_graphs_equal(g, og)


# ---- compress_graph / BaseGraph::combine (src/compression.rs:100-349, src/graph.rs:71-100) ------------------------
def _spec(D, reduce_op):
    return D.ScmapCompress() if reduce_op == 4 else D.SimpleCompress(reduce_op)


def _cg_both(D, ctx, orc, k, host_graphs, stranded=False, reduce_op=3, censor=None):
    dev = [D.BaseGraph.from_host(g, k, ctx=ctx) for g in host_graphs]
    dc = D.BaseGraph.combine(dev)
    oc = orc.combine_graphs(host_graphs)
    assert_graphs_equal(dc.to_host(), oc)
    assert dc.stranded == oc["stranded"]
    og = orc.compress_graph(k, oc, stranded=stranded, reduce_op=reduce_op, censor_nodes=censor)
    assert og["error"] == 0
    dg = D.compress_graph(stranded, _spec(D, reduce_op), dc.finish(), censor)
    assert_graphs_equal(dg.to_host(), og)
    assert dg.stranded == stranded
    return dg, og


def _trivial_graph(orc, k, t, stranded):
    """One node per k-mer of a table: the uncompressed graph (Exts and counts of the table)."""
    n = len(t["lo"])
    keys = [int(lo) | ((int(t["hi"][i]) << 64) if k > 32 else 0) for i, lo in enumerate(t["lo"])]
    kb = np.array([[(x >> (2 * (k - 1 - j))) & 3 for j in range(k)] for x in keys], np.uint8).reshape(-1)
    return dict(n_nodes=n, n_bases=n * k, words=orc.pack_bases(kb), start=np.arange(n, dtype=np.uint64) * np.uint64(k),
                length=np.full(n, k, np.uint32), exts=t["exts"], data=t["counts"], stranded=stranded)


@pytest.mark.parametrize("k", [31, 32, 47])
def test_compress_graph_reassemble_sharded(D, ctx, orc, k):
    """The reference's sharded flow end to end (src/test.rs:418-504): msp shards -> per-shard assemblies -> combine -> compress_graph(max);
    combine and compress_graph on the device, bit for bit against the oracle's greedy walk; the result is_compressed."""
    rng = np.random.default_rng(400 + k)
    for it in range(4):
        contigs = simple_random_contigs(rng) if it == 0 else random_contigs(rng)
        contigs = [c for c in contigs if len(c) >= k]
        shard_graphs = msp_shard_graphs(orc, k, 6, contigs)
        dg, og = _cg_both(D, ctx, orc, k, shard_graphs, reduce_op=3)
        assert dg.is_compressed(D.SimpleCompress(D.MAX)) is None
        if it == 1:   # censored nodes (compress_graph's censor_nodes), every reduce op
            m = sum(g["n_nodes"] for g in shard_graphs)
            censor = sorted(set(int(x) for x in rng.integers(0, m, size=max(1, m // 10))))
            for op in (0, 1, 2, 3, 4):
                _cg_both(D, ctx, orc, k, shard_graphs, reduce_op=op, censor=censor)


@pytest.mark.parametrize("k,stranded", [(4, False), (5, False), (6, False), (8, False), (5, True), (6, True), (12, False)])
def test_compress_graph_small_k(D, ctx, orc, k, stranded):
    """Uncompressed (one node per k-mer) graphs dense in cycles, hairpins and palindromes: compress_graph == the oracle, and without
    censoring == compress_kmers of the same table (same rule, same seed order)."""
    rng = np.random.default_rng(k * 13 + stranded)
    for it in range(25):
        contigs = [c for c in small_k_contigs(rng, alphabet=2 if it % 3 == 0 else 4) if len(c) >= k]
        if not contigs:
            continue
        w, s, l = orc.seqset_from_lists(contigs)
        t = orc.filter_kmers(k, w, s, l, min_obs=1, stranded=stranded)
        triv = _trivial_graph(orc, k, t, stranded)
        dg, og = _cg_both(D, ctx, orc, k, [triv], stranded=stranded, reduce_op=it % 4)
        whole = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded, reduce_op=it % 4)
        assert_graphs_equal(og, whole)
        n = len(t["lo"])
        if n > 3:
            censor = sorted(set(int(x) for x in rng.integers(0, n, size=max(1, n // 5))))
            _cg_both(D, ctx, orc, k, [triv], stranded=stranded, reduce_op=0, censor=censor)


def test_compress_graph_long_chains_cycles_and_long_nodes(D, ctx, orc):
    """Chains of thousands of nodes (pointer doubling), a cycle of thousands of nodes (opened at the seed), nodes of several
    thousand bases in both orientations (multi-chunk emission, reverse complement across word boundaries), K = 31 and 63."""
    rng = np.random.default_rng(77)
    for k in (31, 63):
        genome = random_dna(rng, 6000)
        w, s, l = orc.seqset_from_lists([genome])
        t = orc.filter_kmers(k, w, s, l, min_obs=1)
        whole = orc.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"])
        assert whole["n_nodes"] == 1
        # (a) one node per k-mer: a chain of ~6000 nodes
        dg, og = _cg_both(D, ctx, orc, k, [_trivial_graph(orc, k, t, False)], reduce_op=0)
        assert_graphs_equal(og, whole)
        # (b) a cycle: the genome closed on itself
        circ = np.concatenate([genome, genome[:k - 1]])
        w, s, l = orc.seqset_from_lists([circ])
        tc = orc.filter_kmers(k, w, s, l, min_obs=1)
        dg, og = _cg_both(D, ctx, orc, k, [_trivial_graph(orc, k, tc, False)], reduce_op=1)
        assert og["n_nodes"] == 1 and int(og["length"][0]) == len(tc["lo"]) + k - 1
        # (c) long partial unitigs: the k-mers split by genome position into 3 "shards", each compressed alone, then stitched
        pos = {}
        x, mask = 0, (1 << (2 * k)) - 1
        for i, b in enumerate(genome):
            x = ((x << 2) | int(b)) & mask
            if i >= k - 1:
                r = 0
                y = x
                for _ in range(k):
                    r = (r << 2) | (3 - (y & 3)); y >>= 2
                pos[min(x, r)] = i - k + 1
        keys = [int(lo) | ((int(t["hi"][i]) << 64) if k > 32 else 0) for i, lo in enumerate(t["lo"])]
        part = np.array([min(2, pos[x] // 2100) for x in keys])
        graphs = []
        for sh in range(3):
            m = part == sh
            hi = t["hi"][m] if k > 32 else t["hi"]
            g = orc.compress_kmers(k, t["lo"][m], hi, t["exts"][m], t["counts"][m])
            assert g["error"] == 0 and g["n_nodes"] == 1
            graphs.append(g)
        dg, og = _cg_both(D, ctx, orc, k, graphs, reduce_op=0)
        assert og["n_nodes"] == 1 and int(og["length"][0]) == 6000
        for order in ([2, 0, 1], [1, 2, 0]):
            _cg_both(D, ctx, orc, k, [graphs[i] for i in order], reduce_op=2)


def test_compress_graph_errors_and_empty(D, ctx, orc):
    e = np.zeros(0, np.uint64)
    table, _ = D.filter_kmers((e, e, np.zeros(0, np.uint32)), D.CountFilter(1), False, False, 4, k=31, ctx=ctx)
    g0 = D.compress_kmers_with_hash(False, D.SimpleCompress(), table)
    assert len(D.compress_graph(False, D.SimpleCompress(D.MAX), g0)) == 0
    assert len(D.BaseGraph.combine([g0, g0])) == 0
    ss = orc.synth_reads(500, 1, 0)
    ta, _ = D.filter_kmers(ss, D.CountFilter(1), False, False, 4, k=31, ctx=ctx)
    tb, _ = D.filter_kmers(ss, D.CountFilter(1), True, False, 4, k=31, ctx=ctx)
    ga = D.compress_kmers_with_hash(False, D.SimpleCompress(), ta)
    gb = D.compress_kmers_with_hash(True, D.SimpleCompress(), tb)
    with pytest.raises(D.DbgError):                       # "attempted to combine stranded and unstranded graphs", graph.rs:90-92
        D.BaseGraph.combine([ga, gb])
    with pytest.raises(D.DbgError):
        D.compress_graph(False, D.SimpleCompress(), ga, censor_nodes=[len(ga)])
    # two adjacent nodes whose Exts disagree (A points to B, B has no extension back): the reference panics "unreachable"
    a, b = enc("ACGTTGCATGCCGATAGGCT"), enc("CGTTGCATGCCGATAGGCTA")
    g = dict(n_nodes=2, n_bases=40, words=orc.pack_bases(np.concatenate([a, b])), start=np.array([0, 20], np.uint64),
             length=np.array([20, 20], np.uint32), exts=np.array([1 << 4, 1 << 6], np.uint8), data=np.array([1, 1], np.uint16), stranded=False)
    assert orc.compress_graph(20, g)["error"] == 1
    with pytest.raises(D.DbgError):
        D.compress_graph(False, D.SimpleCompress(), D.BaseGraph.from_host(g, 20, ctx=ctx))


def test_compress_graph_golden_vectors(D, ctx, orc):
    """The device's combine / compress_graph outputs hash to the committed vectors of tests/golden/widened.json (inputs regenerated
    from the same seeds as tests/golden/make_golden.py; the vectors are the oracle's — see that script's header)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = [r for r in json.load(open(os.path.join(HERE, "golden", "widened.json")))["rows"] if r.get("compress_graph")]
    rng = np.random.default_rng(2024)
    assert [r["k"] for r in want] == [31, 47]
    for row in want:
        k = row["k"]
        contigs = [c for c in random_contigs(rng) if len(c) >= k]
        shard_graphs = msp_shard_graphs(orc, k, 6, contigs)
        comb = D.BaseGraph.combine([D.BaseGraph.from_host(g, k, ctx=ctx) for g in shard_graphs])
        ch = comb.to_host()
        m = len(comb)
        censor = sorted(set(int(x) for x in rng.integers(0, m, size=max(1, m // 10))))
        assert m == row["combined_nodes"] and mg.sha(ch["words"], ch["start"], ch["length"], ch["exts"], ch["data"]) == row["combined"]
        cg = D.compress_graph(False, D.SimpleCompress(D.MAX), comb)
        gh = cg.to_host()
        assert len(cg) == row["nodes"] and mg.sha(gh["words"], gh["start"], gh["length"], gh["exts"], gh["data"]) == row["graph"]
        assert mg.sha(cg.to_bincode()) == row["bincode"]
        cgc = D.compress_graph(False, D.SimpleCompress(D.SAT_ADD), comb, censor).to_host()
        assert cgc["n_nodes"] == row["censored_nodes"]
        assert mg.sha(cgc["words"], cgc["start"], cgc["length"], cgc["exts"], cgc["data"]) == row["censored_graph"]
