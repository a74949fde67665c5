"""Distribution of super-k-mer records per MSP bucket (sizing of the per-bucket regions)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_debruijn_b200 as D
ctx = D.Context(0); L = ctx._L
R = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
ss = D.SeqSet.synth(ctx, R, 1, 83886)
n = R * (150 - k + 1)
p, bits = C.c_int(), C.c_int()
ctx.check(L.dbg_plan_filter(ctx._h, k, n, C.byref(p), C.byref(bits)))
part = C.c_void_p()
ctx.check(L.dbg_partition_reads(ctx._h, k, ss._h, 0, p.value, bits.value, C.byref(part)))
counts = np.zeros(1 << bits.value, np.uint32)
ctx.check(L.dbg_partition_bucket_counts(part, C.c_void_p(counts.ctypes.data)))
c = counts.astype(np.float64)
print("p", p.value, "bits", bits.value, "records", int(c.sum()), "mean", c.mean(), "std", c.std(), "max", c.max(), "min", c.min())
for q in (50, 90, 99, 99.9, 99.99):
    print("pct", q, np.percentile(c, q))
for f in (1.1, 1.2, 1.3, 1.5, 2.0):
    cap = int(c.mean() * f)
    ov = np.maximum(c - cap, 0)
    print(f"cap {f}x mean = {cap}: overflowing buckets {int((c > cap).sum())}, overflow records {int(ov.sum())} ({ov.sum() / c.sum() * 100:.3f}%), records in overflowing buckets {int(c[c > cap].sum())}")
