"""Stage times of the multi-rank path with N ranks sharing ONE GPU ("local" transport): for experiments on a 1-GPU box
(e.g. forcing the bucket count of an 8-GPU job).  python tools/multi_local_time.py --ranks 2 --reads 5000000 --bucket-occ 4577"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_debruijn_b200 as D  # noqa: E402
from rust_debruijn_b200 import multi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=2)
ap.add_argument("--reads", type=int, default=5_000_000, help="reads per rank")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--bucket-occ", type=int, default=0)
ap.add_argument("--p", type=int, default=0)
a = ap.parse_args()
mc = multi.MultiContext([0] * a.ranks)
for c in mc.ctxs:
    c.set_param("bucket_occ", a.bucket_occ)
    c.set_param("msp_p", a.p)
sss = [D.SeqSet.synth(mc.ctxs[r], a.reads, 1 + r, 83886) for r in range(a.ranks)]
for i in range(a.steps):
    gs = mc.reads_to_graph(sss, D.CountFilter(2), D.SimpleCompress(D.SAT_ADD), stranded=False, k=31)
    i0 = gs[0].info
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in i0.items() if k.startswith("ms_") or k in ("msp_p", "bucket_bits", "check_ok")}, flush=True)
    for g in gs:
        g.free()
mc.close()
