#!/usr/bin/env python
"""bench.py — k-mers/sec through filter_kmers + compress_kmers (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --config c3 ...                             # configs[2] (K=63) as the main workload
    python bench.py --impl reference --gpus N --steps K ...     # reference algorithm on the host cores

One "step" = one pass of the whole hot path (reads -> valid k-mer table -> BaseGraph) over one batch of
synthetic reads.  N=1 workload = BASELINE.json configs[1]: K=31, 10M x 150 bp synth-v1 reads (noisy,
e=0.5%, 50x, CountFilter(2), SimpleCompress(sat_add), stranded=false), all MSP buckets on one GPU; the same
line carries a "c3" object with configs[2] (K=63, same reads) measured the same way with its own roofline.
`value`   : device-resident input (reads already in HBM) -> BaseGraph arrays in HBM.
`e2e`     : same metric through the reference-facing C-ABI call with HOST (pinned) buffers: H2D of the
            packed reads and D2H of the BaseGraph arrays inside the timed region.
Timing: CUDA events recorded on the library's own stream (dbg_ctx_stream) at every step boundary, barrier +
synchronize on both sides, max over ranks; ms_per_step = (last - first event) / K, the per-step event times
give median / min / max.  SM clocks and throttle reasons are read through NVML from the MAIN thread between
steps (no polling thread: a Python poller contends for the GIL with the stepping thread and showed up as
24-98 ms step walls in round 1).  Inputs (375 MB packed reads) and every intermediate are larger than the
126 MB L2, so no L2 flush is needed between iterations (stated in config.l2).
N > 1: one job over all ranks (dbg_reads_to_graph_multi: NCCL / NVLink inside the library), every step's
result is checked against cheap global invariants (all-reduced over the ranks) and the line says so.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MIN_OBS = 2
ERR_THR_NOISY = 83886
METRIC = "k-mers/sec filter_kmers+compress_kmers K=31 150bp reads"
UNIT = "k-mers/s"
CONFIGS = {"c2": dict(k=31, label="configs[1]"), "c3": dict(k=63, label="configs[2]"),
           # configs[3]: 100M reads over 8 GPUs = 12.5M reads per rank (run with --gpus 8 under torchrun)
           "c4": dict(k=31, label="configs[3]", reads=12_500_000)}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe): two light NVML reads per
    sample, taken by the main thread between steps.  nvidia-smi (one query) is only the fallback without pynvml."""

    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.nv, self.h = index, None, None
        self.sm, self.reasons, self.mx = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")   # NVML indexes physical devices
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nv = pynvml
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        try:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = int(self.get_reasons(self.h))
            for name, bit in self.BITS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def result(self):
        if self.nv is not None:
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, main thread, between steps"}
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            sm, mx = (float(x) for x in out.strip().split(","))
            return {"sm_mhz": sm, "sm_max_mhz": mx, "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def ncu_traffic():
    """dram bytes per launch of the dominant kernels from the committed ncu --set full captures (configs[1] launches:
    profiles/traffic.json, details of the round-2 captures in profiles/traffic_r02.json), if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def cpu_baseline(sample_reads, threads, k=31):
    """Reference algorithm restated in C++ (oracle/), timed on the host cores on a bounded sample."""
    import oracle as O
    words, start, length = O.synth_reads(sample_reads, 1, ERR_THR_NOISY)
    t0 = time.perf_counter()
    t = O.filter_kmers(k, words, start, length, min_obs=MIN_OBS, stranded=False, memory_gb=4, threads=threads)
    t1 = time.perf_counter()
    g = O.compress_kmers(k, t["lo"], t["hi"], t["exts"], t["counts"], stranded=False, reduce_op=O.SAT_ADD)
    t2 = time.perf_counter()
    n = t["n_input"]
    return {"value": n / (t2 - t0), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"synth-v1 noisy, {sample_reads} x 150bp reads (N={n} k-mers), K={k}, seed 1; "
                      f"filter {t1 - t0:.2f}s + compress {t2 - t1:.2f}s; host has {os.cpu_count()} cpus; "
                      "oracle built -O3 -march=native on this host",
            "n_valid": int(len(t["lo"])), "n_nodes": int(g["n_nodes"])}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (C++ restatement — the crate is Rust and cannot be
    built in this image) on the host cores, rank 0 only."""
    if rank != 0:
        return
    k = CONFIGS[args.config]["k"]
    threads = os.cpu_count() or 1
    sample = args.ref_reads
    for _ in range(args.warmup):
        cpu_baseline(max(sample // 10, 1000), threads, k)
    vals, ms = [], []
    last = None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        last = cpu_baseline(sample, threads, k)
        ms.append((time.perf_counter() - t0) * 1e3)
        vals.append(last["value"])
    v = statistics.mean(vals)
    last["value"] = v
    last["sample"] += " per step; filter stage parallel over sequence ranges and the 256 prefix buckets, compress is serial by construction"
    emit_json({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": statistics.mean(ms), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64" if k <= 32 else "u128", "data": "synthetic",
        "config": {"workload": f"K={k}, {sample} x 150bp synth-v1 noisy reads per step (bounded sample of {CONFIGS[args.config]['label']}), host cores",
                   "k": k, "min_kmer_obs": MIN_OBS, "stranded": False},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def emit_json(obj):
    """The ONE JSON line goes to the real stdout; everything else any library prints on fd 1 during the run
    (e.g. NCCL's version banner) was diverted to stderr by divert_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def divert_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def step_stats(ms):
    return {"median": round(statistics.median(ms), 3), "mean": round(statistics.mean(ms), 3), "min": round(min(ms), 3),
            "max": round(max(ms), 3)}


def main():
    divert_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="c2 = configs[1] (K=31), c3 = configs[2] (K=63), c4 = configs[3] (K=31, 100M reads over 8 GPUs)")
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (configs[1], configs[2]: 10M)")
    ap.add_argument("--cpu-sample-reads", type=int, default=1_000_000)
    ap.add_argument("--ref-reads", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2] (K=63) object of the default N=1 run")
    ap.add_argument("--no-clocks", action="store_true", help="debug: do not sample clocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if "reads" in CONFIGS[args.config]:
        args.reads = CONFIGS[args.config]["reads"]

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import rust_debruijn_b200 as D

    if world > 1:
        torch.cuda.set_device(local)
        # torch.distributed is the rendezvous only (barriers, the NCCL id broadcast, the max-over-ranks of the timings);
        # the data path's collectives run inside libdbg_b200.so on its own communicator
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = D.Context(local)
    L = ctx._L
    ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local))
    comm = None
    if world > 1:
        from rust_debruijn_b200 import multi
        comm = multi.Comm.from_torch(ctx)

    def barrier():
        torch.cuda.synchronize(local)
        ctx.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local)

    R = args.reads
    # weak scaling: every rank owns R reads of the same synth-v1 family (seed = 1 + rank => independent shards)
    ss = D.SeqSet.synth(ctx, R, 1 + rank, ERR_THR_NOISY)
    filt, spec = D.CountFilter(MIN_OBS), D.SimpleCompress(D.SAT_ADD)
    hw, hs, _ = ss.copy_out()
    pw = torch.empty(len(hw), dtype=torch.int64, pin_memory=True)
    pw.numpy()[:] = hw.view(np.int64)
    words_pinned = pw.numpy().view(np.uint64)
    n_reads = len(hs)
    del hw, hs
    peak, peak_kind = hbm_peak()
    tr = ncu_traffic()

    def disturbed(ms):
        """Host interference on the shared box shows as isolated steps far above the median (the kernels themselves
        repeat to within 1%): such a run is re-measured ONCE (B200_PROFILING.md timing hygiene); both numbers are kept."""
        flag = torch.tensor([1.0 if max(ms) > 1.25 * statistics.median(ms) else 0.0], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)   # every rank must take the same decision
        return bool(flag.item())

    def timed(fn, sampler=None):
        """K steps of fn between barriers; events at every step boundary on the library stream."""
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev[0].record(ext)
        res = None
        for i in range(args.steps):
            res = fn()
            ev[i + 1].record(ext)
            if sampler is not None and (i + 1 == args.steps // 2 or i + 1 == args.steps):
                sampler.sample()   # two samples per run: NVML reads were seen to stall a step now and then
        barrier()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        return ev[0].elapsed_time(ev[-1]), per, res

    def measure(k):
        """device-resident + e2e measurement of one configuration; returns the pieces of the JSON line."""
        out_info = {}

        def step():
            if world > 1:
                g = comm.reads_to_graph(ss, filt, spec, stranded=False, k=k)
                out_info["nodes_total"], out_info["bases_total"] = g.n_nodes_total, g.n_bases_total
                out_info["n_valid_total"] = g.n_valid_total
                out_info["check"] = g.invariants
                out_info["stage_ms_rank0"] = {k_: round(v, 3) for k_, v in g.info.items() if k_.startswith("ms_")}
                out_info["queries_sent_rank0"], out_info["exchange_bytes_sent_rank0"] = g.info["n_queries_sent"], g.info["exchange_bytes_sent"]
                out_info["replicated"], out_info["transport"] = g.replicated, comm.transport
            else:
                g = D.reads_to_graph(ss, filt, spec, stranded=False, k=k)
            n = len(g)
            g.free()
            return n

        sampler = ClockSampler(local)
        for _ in range(args.warmup):
            step()
            if not args.no_clocks:
                sampler.sample()   # NVML's first reads are slow (lazy initialisation) and showed up as a slow 2nd timed step
        sampler.sm, sampler.reasons = [], set()
        st0 = ctx.stats()
        ms_total, per, _ = timed(step, None if args.no_clocks else sampler)
        remeasured = None
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & sampler.reasons
        if (world == 1 and bad) or disturbed(per):
            remeasured = {"first_ms_per_step": ms_total / args.steps, "first_step_ms": [round(x, 3) for x in per],
                          "reason": sorted(bad) or ["isolated slow steps: host interference"]}
            sampler.sm, sampler.reasons = [], set()
            st0 = ctx.stats()
            ms_total, per, _ = timed(step, None if args.no_clocks else sampler)
        st1 = ctx.stats()
        clocks = sampler.result() if not args.no_clocks else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler off"]}
        print(f"[bench] rank {rank} K={k} per-step device ms: {[round(x, 2) for x in per]}", file=sys.stderr)

        # ---- e2e: host pinned buffers in, BaseGraph arrays out, through the C-ABI host entry point ----
        M0, nb0 = st1["n_nodes"], st1["n_bases"]  # world > 1: this rank's run of nodes
        cap_nodes, cap_words = int(M0 * 1.2) + 1024, int(nb0 * 1.2) // 32 + 1024
        ob = {"words": torch.empty(cap_words, dtype=torch.int64, pin_memory=True),
              "start": torch.empty(cap_nodes, dtype=torch.int64, pin_memory=True),
              "length": torch.empty(cap_nodes, dtype=torch.int32, pin_memory=True),
              "exts": torch.empty(cap_nodes, dtype=torch.uint8, pin_memory=True),
              "data": torch.empty(cap_nodes, dtype=torch.int16, pin_memory=True)}

        def e2e_step():
            gh = C.c_void_p()
            if world > 1:
                sse = D.SeqSet.upload_uniform(ctx, words_pinned, n_reads, 150, pipelined=True)
                ge = comm.reads_to_graph(sse, filt, spec, stranded=False, k=k)
                sse.free()
                gh, ge._h = ge._h, None
            else:
                ctx.check(L.dbg_reads_to_graph_host_uniform(ctx._h, k, C.c_void_p(words_pinned.ctypes.data), len(words_pinned),
                                                            n_reads, 150, None, MIN_OBS, 0, D.SAT_ADD, None, C.byref(gh)))
            m, nw = L.dbg_graph_len(gh), L.dbg_graph_n_words(gh)
            assert m <= cap_nodes and nw <= cap_words
            ctx.check(L.dbg_graph_copy_out(gh, C.c_void_p(ob["words"].data_ptr()), C.c_void_p(ob["start"].data_ptr()),
                                           C.c_void_p(ob["length"].data_ptr()), C.c_void_p(ob["exts"].data_ptr()),
                                           C.c_void_p(ob["data"].data_ptr())))
            L.dbg_graph_free(gh)
            return m, nw

        for _ in range(args.warmup):   # same W untimed steps as the device-resident arm (the scratch arena re-sizes itself
            e2e_step()                 # once for this entry point's allocation pattern: not part of the steady state)
        ms_e2e, e2e_per, (m_nodes, n_gw) = timed(e2e_step)
        e2e_remeasured = None
        if disturbed(e2e_per):
            e2e_remeasured = {"first_ms_per_step": ms_e2e / args.steps, "first_step_ms": [round(x, 3) for x in e2e_per]}
            ms_e2e, e2e_per, (m_nodes, n_gw) = timed(e2e_step)
        print(f"[bench] rank {rank} K={k} e2e per-step device ms: {[round(x, 2) for x in e2e_per]}", file=sys.stderr)
        h2d = int(len(words_pinned) * 8)
        d2h = int(n_gw * 8 + m_nodes * (8 + 4 + 1 + 2))
        n_kmers = R * (150 - k + 1)

        # ---- reduce over ranks: max time, sum of units ----
        times = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=f"cuda:{local}")
        units = torch.tensor([float(n_kmers), float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(units, op=dist.ReduceOp.SUM)
        ms_total, ms_e2e = float(times[0]), float(times[1])
        total_kmers = float(units[0])
        h2d, d2h = int(units[1]), int(units[2])   # whole job: every rank uploads its reads and reads back its nodes
        ms_step = ms_total / args.steps
        # roofline of the dominant kernel (SURVEY.md §8(d) algorithmic bytes: S1 = 150R/4 + (w+1)N, S2 = (w+1)N + (w+3)V)
        w = 8 if k <= 32 else 16
        N = n_kmers
        V = st1["n_valid"] if world == 1 else out_info.get("n_valid_total", 0) // world
        cand = {"msp_tile_kernel": (st1["ms_k_partition"], 150 * R / 4 + (w + 1) * N),
                "count_kernel": (st1["ms_k_count"], (w + 1) * N + (w + 3) * V)}
        dom = max(cand, key=lambda k_: cand[k_][0])
        ach = cand[dom][1] / (cand[dom][0] * 1e-3) / 1e9 if cand[dom][0] > 0 else 0.0
        M_, Lb_ = (st1["n_nodes"], st1["n_bases"]) if world == 1 else (out_info.get("nodes_total", 0) / world, out_info.get("bases_total", 0) / world)
        step_bytes = N * (37.5 / (151 - k) + 2 * (w + 1)) + V * (7 * w + 55) + Lb_ / 4 + 15 * M_   # SURVEY §8(d) total
        res = {
            "value": total_kmers / (ms_step * 1e-3), "ms_per_step": ms_step, "step_ms": step_stats(per),
            "clocks": clocks, "remeasured": remeasured,
            "e2e": {"value": total_kmers / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps, "step_ms": step_stats(e2e_per),
                    "remeasured": e2e_remeasured},
            "gpu_launches": int(st1["gpu_launches"] - st0["gpu_launches"]),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "peak_kind": peak_kind, "traffic": tr.get(f"{dom}_k{k}", tr.get(dom) if k <= 32 else None),
                         "kernel_ms": cand[dom][0],
                         "other": {n_: {"kernel_ms": round(v[0], 3), "frac": round(v[1] / (v[0] * 1e-3) / 1e9 / peak, 4) if v[0] > 0 else None}
                                   for n_, v in cand.items() if n_ != dom},
                         "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                                        "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak}},
            "stage_ms": {k_: round(st1[k_], 3) for k_ in ("ms_partition", "ms_count", "ms_sort", "ms_table", "ms_links",
                                                         "ms_rank", "ms_emit", "ms_k_partition", "ms_k_count",
                                                         "ms_filter_total", "ms_compress_total")},
            "counters": {k_: st1[k_] for k_ in ("n_records", "n_records_distinct", "n_buckets", "n_distinct", "n_bucket_splits",
                                                "rank_rounds", "n_cycle_kmers", "direct_partition")},
            "_cfg": {"k": k, "reads_per_gpu": R, "input_kmers_per_gpu": N, "valid_kmers": V,
                     "nodes": st1["n_nodes"] if world == 1 else out_info.get("nodes_total"),
                     "node_bases": st1["n_bases"] if world == 1 else out_info.get("bases_total"),
                     "msp_p": st1["msp_p"], "bucket_bits": st1["bucket_bits"]},
            "_info": out_info,
        }
        return res

    k_main = CONFIGS[args.config]["k"]
    main_res = measure(k_main)
    c3 = None
    if world == 1 and args.config == "c2" and not args.no_c3:
        c3 = measure(63)

    if rank == 0:
        cfgd = main_res.pop("_cfg")
        info = main_res.pop("_info")
        out = {
            "metric": METRIC, "value": main_res.pop("value"), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64" if k_main <= 32 else "u128", "data": "synthetic",
            "config": dict(cfgd, workload=f"{CONFIGS[args.config]['label']}: K={k_main}, {R} x 150bp synth-v1 noisy reads per GPU (e=0.5%, 50x), "
                           "CountFilter(2), SimpleCompress(sat_add), stranded=false, " +
                           ("all MSP buckets on one GPU" if world == 1 else f"one job of {world * R} reads, MSP buckets sharded over {world} GPUs"),
                           l2="inputs and every intermediate exceed the 126 MB L2; no flush needed",
                           parallelism=("MSP-bucket-sharded filter_kmers (one NCCL all-to-all of super-k-mer records inside the library); the "
                                        "k-mer table stays bucket-sharded: remote neighbour queries by all-to-all, unitig walks over NVLink "
                                        "peer memory, path records shipped to the rank owning their seed-key range; every rank keeps its "
                                        "run of nodes of the complete BaseGraph") if world > 1 else "single"),
        }
        out.update(main_res)
        if world > 1:
            out["multi"] = {k_: v for k_, v in info.items()}
        if c3 is not None:
            c3cfg = c3.pop("_cfg")
            c3.pop("_info")
            c3["config"] = dict(c3cfg, workload=f"configs[2]: K=63 (two-u64 keys), {R} x 150bp synth-v1 noisy reads, 1 GPU")
            c3["unit"], c3["dtype"] = UNIT, "u128"
            out["c3"] = c3
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args.cpu_sample_reads, 1, k_main)
        emit_json(out)
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
