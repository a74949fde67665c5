"""Minimal driver for ncu: N steps of the configs[1] hot path (device-resident input), nothing else."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_debruijn_b200 as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--k", type=int, default=31)
ap.add_argument("--clean", action="store_true")
ap.add_argument("--p", type=int, default=0)
ap.add_argument("--bucket-occ", type=int, default=0)
ap.add_argument("--no-direct", action="store_true")
ap.add_argument("--dedup", type=int, default=-1)
a = ap.parse_args()
ctx = D.Context(0)
if a.p:
    ctx.set_param("msp_p", a.p)
if a.no_direct:
    ctx.set_param("direct_partition", 0)
if a.dedup >= 0:
    ctx.set_param("dedup", a.dedup)
if a.bucket_occ:
    ctx.set_param("bucket_occ", a.bucket_occ)
ss = D.SeqSet.synth(ctx, a.reads, 1, 0 if a.clean else 83886)
for i in range(a.steps):
    g = D.reads_to_graph(ss, D.CountFilter(1 if a.clean else 2), D.SimpleCompress(D.SAT_ADD), k=a.k)
    s = ctx.stats()
    print({k_: (round(v, 3) if isinstance(v, float) else v) for k_, v in s.items()}, flush=True)
    g.free()
