"""Per-file and (for one file) per-line-range share of executed warp instructions / stall samples of an ncu source-page CSV.
usage: ncu_src_files.py src.csv [file-substring name:lo-hi ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sel = sys.argv[2] if len(sys.argv) > 2 else None
ranges = []
for a in sys.argv[3:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((n, int(lo), int(hi)))
ie = samp = None
cur = None
fname = "?"
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
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
    name = fname
    if sel and sel in fname:
        name = fname + ":other"
        for n, lo, hi in ranges:
            if cur is not None and lo <= cur <= hi:
                name = n
                break
    a = agg.setdefault(name, [0.0, 0.0])
    a[0] += v
    a[1] += s
tot = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot:.4g}, samples {ts:.4g}")
for n, (v, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:28s} {100 * v / tot:6.2f}% inst  {100 * s / max(ts, 1):6.2f}% samples")
