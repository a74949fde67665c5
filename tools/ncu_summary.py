"""Summarise ncu outputs into profiles/: launch-list shares per kernel and key --set full metrics per kernel.
usage: ncu_summary.py launches.csv raw.csv out.md traffic.json title"""
import collections
import csv
import json
import sys

launch_csv, raw_csv, out_md, traffic_json, title = sys.argv[1:6]
rows = list(csv.reader(open(launch_csv)))
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
            agg.setdefault(name, [0, 0.0])
            agg[name][0] += 1
            agg[name][1] += float(d["Metric Value"].replace(",", "")) / 1e6
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
cmd = sys.argv[7] if len(sys.argv) > 7 else f"python tools/profile_step.py --steps {steps}"
synth = agg.pop("synth_reads_kernel", [0, 0.0])
tot = sum(v[1] for v in agg.values())
out = [f"# {title}\n", "## Launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n",
       f"`{cmd}` (configs[1]: 10M x 150bp noisy, K=31); synth_reads_kernel ({synth[1]:.2f} ms, input generation) excluded.\n",
       "| kernel | launches/step | ms/step | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {v[0] / steps:g} | {v[1] / steps:.3f} | {v[1] / tot * 100:.1f}% |")
out.append(f"| **total** | | **{tot / steps:.3f}** | |\n")
rr = list(csv.reader(open(raw_csv)))
h, units = rr[0], rr[1]
idx = {x: i for i, x in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
out.append("## ncu --set full --clock-control none --import-source on (one launch per kernel)\n")
traffic = {}
best = collections.OrderedDict()   # one launch per kernel name: the longest (e.g. the main pass, not the sampling pass)
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    dur = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    if name not in best or dur > best[name][0]:
        best[name] = (dur, r)
for name, (_, r) in best.items():
    out.append(f"### {name}\n\n| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in idx:
            out.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    out.append("")

    def tobytes(val, unit):
        v = float(val.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rd = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    base = name.split("<")[0]
    key = {"msp_tile_kernel": "msp_partition_kernel"}.get(base, base)
    traffic[key] = rd + wr
open(out_md, "w").write("\n".join(out) + "\n")
json.dump(traffic, open(traffic_json, "w"), indent=1)
print("\n".join(out[:40]))
