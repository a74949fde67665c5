"""One-screen summary of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        e = d.get("e2e") or {}
        print(f"{f}: n_gpus={d.get('n_gpus')} value={d['value'] / 1e9:.2f} G/s ms/step={d.get('ms_per_step', 0):.2f} "
              f"e2e={e.get('value', 0) / 1e9:.2f} G/s ({e.get('ms_per_step', 0):.2f} ms)")
        print("   stage_ms", d.get("stage_ms"))
        r = d.get("roofline") or {}
        print("   roofline", {k: r.get(k) for k in ("kernel", "kernel_ms", "frac", "other", "whole_step")})
        if d.get("multi"):
            print("   multi", d["multi"].get("stage_ms_rank0"))
        if d.get("c3"):
            c3 = d["c3"]
            print(f"   c3 value={c3['value'] / 1e9:.2f} G/s ms={c3['ms_per_step']:.2f} e2e={c3['e2e']['value'] / 1e9:.2f}", c3.get("stage_ms"))
        if d.get("cpu_baseline"):
            print("   cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"))
