"""Small all-paths driver for compute-sanitizer (memcheck / racecheck): both key widths, tile + general partition
kernels, dedup, bucket splits, cycles, long unitigs, report_all, table_from_host."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
import rust_debruijn_b200 as D  # noqa: E402

ctx = D.Context(0)
rng = np.random.default_rng(1)
ok = True


def check(k, seqset, mo, stranded=False, report_all=False, c=ctx):
    global ok
    t, _ = D.filter_kmers(seqset, D.CountFilter(mo), stranded, report_all, 4, k=k, ctx=c)
    g = D.compress_kmers_with_hash(stranded, D.SimpleCompress(D.SAT_ADD), t).to_host()
    ot = O.filter_kmers(k, *seqset[:3], min_obs=mo, stranded=stranded, report_all=report_all)
    og = O.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"], stranded=stranded)
    same = all(np.array_equal(g[f], og[f]) for f in ("words", "start", "length", "exts", "data"))
    ok &= same
    print("k", k, "stranded", stranded, "nodes", g["n_nodes"], "OK" if same else "MISMATCH", flush=True)


check(31, O.synth_reads(1500, 1, O.ERR_THR_NOISY), 2, report_all=True)
check(63, O.synth_reads(1000, 1, O.ERR_THR_NOISY), 2)
circ = rng.integers(0, 4, 2500, dtype=np.uint8)
seqs = [np.concatenate([circ, circ[:40]]), rng.integers(0, 4, 3000, dtype=np.uint8), rng.integers(0, 2, 300, dtype=np.uint8)]
check(31, O.seqset_from_lists(seqs), 1)
check(6, O.seqset_from_lists([rng.integers(0, 2, 60, dtype=np.uint8) for _ in range(8)]), 1, stranded=True)
big = rng.integers(0, 4, 3000, dtype=np.uint8)
words = O.pack_bases(big)
check(31, (words, np.array([2000, 100], np.uint64), np.array([900, 700], np.uint32)), 1)   # general (non-contiguous) kernel
c3 = D.Context(0)
c3.set_param("direct_min_tiles", 1)
check(31, O.synth_reads(4000, 1, O.ERR_THR_NOISY), 2, c=c3)                                 # direct partition (sampling + regions)
check(31, O.synth_reads(200, 1, 0), 1, c=c3)                                                # ... with a useless sample: overflow fallback
c3.set_param("fast_compress", 0)
check(31, O.synth_reads(1500, 1, O.ERR_THR_NOISY), 2, c=c3)                                 # general compression path forced
ascii_reads = [bytes(rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), size=n)) for n in (0, 31, 150, 33, 4097, 1)]
ssa = D.SeqSet.from_ascii(ctx, ascii_reads)
ow, ost, oln, obad = O.from_acgt_bytes(ascii_reads)
same = np.array_equal(ssa.copy_out()[0], ow) and ssa.n_invalid == obad
ok &= same
print("from_ascii", "OK" if same else "MISMATCH", flush=True)
c2 = D.Context(0)
c2.set_param("bucket_occ", 1 << 30)
check(31, O.synth_reads(1200, 1, O.ERR_THR_NOISY), 2, c=c2)                                 # bucket splits
print("ALL OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
