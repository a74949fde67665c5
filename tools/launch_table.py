"""Per-kernel totals from `ncu --metrics gpu__time_duration.sum --csv` launch lists: python tools/launch_table.py launches.csv [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            v = float(d["Metric Value"].replace(",", ""))
            unit = d.get("Metric Unit", "ns")
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
            agg.setdefault(name, [0, 0.0])
            agg[name][0] += 1
            agg[name][1] += v
tot = sum(v[1] for k, v in agg.items() if not k.startswith("synth_reads"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:45s} {v[0] / steps:7.1f} launches/step {v[1] / steps:9.3f} ms/step {v[1] / tot * 100:6.1f}%")
print(f"total (excl. synth) {tot / steps:.3f} ms/step")
