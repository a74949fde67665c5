"""Share of executed warp instructions / stall samples per named source-line range of one file.
usage: ncu_src_ranges.py src.csv name:lo-hi name:lo-hi ...   (lines outside every range are reported as 'other')"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ranges = []
for a in sys.argv[2:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((n, int(lo), int(hi)))
ie = samp = None
cur = None
agg = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r:
        ie, samp = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if ie is None or len(r) <= ie:
        continue
    if r[0] != "":
        try:
            cur = int(r[0])
        except ValueError:
            cur = None
        continue
    try:
        v = float(r[ie])
    except ValueError:
        continue
    s = float(r[samp]) if r[samp] not in ("", "-") else 0.0
    name = "other"
    for n, lo, hi in ranges:
        if cur is not None and lo <= cur <= hi:
            name = n
            break
    a = agg.setdefault(name, [0.0, 0.0])
    a[0] += v
    a[1] += s
tot = sum(v[0] for v in agg.values())
stot = sum(v[1] for v in agg.values())
for n, (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:14s} {v / tot * 100:6.2f}% inst  {s / max(stot, 1) * 100:6.2f}% samples")
