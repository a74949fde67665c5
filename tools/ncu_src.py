"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:<kernel>`:
share of executed warp instructions and of stall samples per CUDA source line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
cur = None
ie = samp = None
for r in rows:
    if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r:
        ie, samp = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if ie is None or len(r) <= ie:
        continue
    if r[0] != "":
        cur = (r[0], r[1].strip()[:110])
        agg.setdefault(cur, [0.0, 0.0])
        continue
    try:
        v = float(r[ie])
    except ValueError:
        continue
    s = float(r[samp]) if r[samp] not in ("", "-") else 0.0
    if cur is not None:
        agg[cur][0] += v
        agg[cur][1] += s
tot = sum(v[0] for v in agg.values())
stot = sum(v[1] for v in agg.values())
print("total warp-inst %.4g  samples %.4g" % (tot, stot))
for (l, t), (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v / tot * 100:6.2f}% inst {s / max(stot, 1) * 100:6.2f}% smp  L{l:>4s}  {t}")
