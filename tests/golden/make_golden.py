"""Regenerates tests/golden/widened.json: checksums of the oracle's outputs for the widened rows on seeded
synth-v1 reads.  The reference ships no expected outputs for these functions (and cannot be built here), so the vectors pin
the ORACLE against accidental change; the GPU is compared with the oracle directly in tests/test_gpu_parity.py.
    python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(a if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:32]


def rows():
    out = []
    for R, k, stranded in ((1000, 31, False), (1000, 63, False), (1000, 32, True)):
        w, st, ln = O.synth_reads(R, 1, O.ERR_THR_NOISY)
        t = O.filter_kmers(k, w, st, ln, min_obs=2, stranded=stranded, report_all=True)
        e1 = O.remove_censored_exts(k, t, stranded=stranded)
        g = O.compress_kmers(k, t["lo"], t["hi"], e1, t["counts"], stranded=stranded)
        target, flags, pair = O.graph_edges(k, g, stranded=stranded)
        gs = O.compress_kmers(k, t["lo"], t["hi"], t["exts"], (t["counts"] % 3).astype(np.uint16), stranded=stranded, reduce_op=O.SCMAP)
        g0 = O.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=stranded)
        out.append(dict(R=R, k=k, stranded=stranded, n_valid=int(len(t["lo"])), censored_exts=sha(e1),
                        pruned_nodes=int(g["n_nodes"]), pruned_graph=sha(g["words"], g["start"], g["length"], g["exts"], g["data"]),
                        edges=sha(target, flags), n_edges=int((target != 0xffffffff).sum()), is_compressed=pair is None,
                        fix_exts=sha(O.graph_fix_exts(k, g0, stranded=stranded)), gfa=sha(O.write_gfa(k, g, stranded=stranded).encode()),
                        scmap_nodes=int(gs["n_nodes"]), scmap_graph=sha(gs["words"], gs["start"], gs["length"], gs["exts"], gs["data"])))
    # round 2: compress_graph / combine over the reference's sharded flow (src/test.rs:418-504), hashn ingest, bincode image
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import msp_shard_graphs, random_contigs   # noqa: E402
    rng = np.random.default_rng(2024)
    for k in (31, 47):
        contigs = [c for c in random_contigs(rng) if len(c) >= k]
        shard_graphs = msp_shard_graphs(O, k, 6, contigs)
        comb = O.combine_graphs(shard_graphs)
        m = int(comb["n_nodes"])
        censor = sorted(set(int(x) for x in rng.integers(0, m, size=max(1, m // 10))))
        cg = O.compress_graph(k, comb, reduce_op=O.MAX)
        cgc = O.compress_graph(k, comb, reduce_op=O.SAT_ADD, censor_nodes=censor)
        out.append(dict(compress_graph=True, k=k, n_shards=len(shard_graphs), combined_nodes=m,
                        combined=sha(comb["words"], comb["start"], comb["length"], comb["exts"], comb["data"]),
                        nodes=int(cg["n_nodes"]), graph=sha(cg["words"], cg["start"], cg["length"], cg["exts"], cg["data"]),
                        censored_nodes=int(cgc["n_nodes"]), censored_graph=sha(cgc["words"], cgc["start"], cgc["length"], cgc["exts"], cgc["data"]),
                        bincode=sha(O.graph_to_bincode(cg))))
    hw, hst, hln, hbad = O.from_acgt_bytes_hashn([b"ACGTNNacgtXy-", b"NNNN", b""], [b"read/1", b"r2", b"empty"])
    out.append(dict(hashn=True, words=sha(hw), n_invalid=int(hbad), siphash13_abc=int(O.siphash13(b"abc"))))
    ascii_reads = [b"ACGTNNacgtXy-", b"", b"TTTTTGGGGGCCCCCAAAAATTTTTGGGGGCCCCCAAAAAT", b"g"]
    aw, ast_, aln, abad = O.from_acgt_bytes(ascii_reads)
    out.append(dict(ascii=True, words=sha(aw), start=[int(x) for x in ast_], length=[int(x) for x in aln], n_invalid=int(abad)))
    return out


if __name__ == "__main__":
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "widened.json")
    json.dump({"_doc": __doc__.split("\n")[0], "rows": rows()}, open(p, "w"), indent=1)
    print("wrote", p)
