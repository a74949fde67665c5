"""Size-independent properties of the path at BASELINE.json's full single-GPU size (configs[1]: 10M x 150bp, K=31), where the
oracle is too slow to be the checker: ascending distinct k-mers, V <= U, node bookkeeping (sum of lengths = V + M(K-1), start
= exclusive scan of length), every table k-mer appears in exactly one node (k-mers re-extracted from a sample of nodes are in
the table, node data = sat-add of their counts), and the staged and direct partitions give identical tables."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_debruijn_b200 as D  # noqa: E402

R, K = (int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000), 31
ctx = D.Context(0)
ss = D.SeqSet.synth(ctx, R, 1, 83886)
table, _ = D.filter_kmers(ss, D.CountFilter(2), False, False, 0, k=K)   # memory_size 0 = no cap: one pass
st = ctx.stats()
graph = D.compress_kmers_with_hash(False, D.SimpleCompress(D.SAT_ADD), table)
t, g = table.to_host(), graph.to_host()
V, M = len(t["lo"]), g["n_nodes"]
ok = True


def check(name, cond):
    global ok
    ok &= bool(cond)
    print(("ok   " if cond else "FAIL ") + name, flush=True)


check(f"direct partition used (records {st['n_records']})", st["direct_partition"] == 1)
check(f"k-mers strictly ascending (V = {V})", np.all(t["lo"][1:] > t["lo"][:-1]))
check(f"sum of node lengths = V + M(K-1) (M = {M})", int(g["length"].astype(np.int64).sum()) == V + M * (K - 1) == g["n_bases"])
check("start = exclusive scan of length", np.array_equal(g["start"][1:], np.cumsum(g["length"].astype(np.uint64))[:-1]) and g["start"][0] == 0)
check("counts >= min_kmer_obs", int(t["counts"].min()) >= 2)
# a sample of nodes: their k-mers (canonical) are table k-mers, each once; data = saturating sum of the counts
rng = np.random.default_rng(0)
mask = np.uint64((1 << (2 * K)) - 1)
seen = 0
nodes_ok = True
for n in rng.integers(0, M, size=3000):
    idx = np.arange(int(g["start"][n]), int(g["start"][n]) + int(g["length"][n]), dtype=np.uint64)
    b = (g["words"][(idx >> np.uint64(5)).astype(np.int64)] >> (np.uint64(62) - np.uint64(2) * (idx & np.uint64(31)))) & np.uint64(3)
    x, tot = 0, 0
    for j, base in enumerate(b):
        x = ((x << 2) | int(base)) & int(mask)
        if j >= K - 1:
            r = 0
            y = x
            for _ in range(K):
                r = (r << 2) | (3 - (y & 3))
                y >>= 2
            c = min(x, r)
            p = int(np.searchsorted(t["lo"], np.uint64(c)))
            if p >= V or int(t["lo"][p]) != c:
                nodes_ok = False
                print("FAIL node k-mer not in the table", n, j)
                break
            tot += int(t["counts"][p])
            seen += 1
    if min(tot, 65535) != int(g["data"][n]):
        nodes_ok = False
        print("FAIL node data", n, tot, int(g["data"][n]))
check(f"{seen} k-mers of 3000 sampled nodes found in the table with matching node data", nodes_ok)
# the staged partition must give the same table bit for bit
c2 = D.Context(0)
c2.set_param("direct_partition", 0)
ss2 = D.SeqSet.synth(c2, R, 1, 83886)
t2 = D.filter_kmers(ss2, D.CountFilter(2), False, False, 0, k=K)[0].to_host()
check("staged partition: identical table", all(np.array_equal(t[f], t2[f]) for f in ("lo", "exts", "counts")) and c2.stats()["direct_partition"] == 0)
print("FULLSIZE OK" if ok else "FULLSIZE FAILED", flush=True)
sys.exit(0 if ok else 1)
