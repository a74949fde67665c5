"""torchrun --nproc-per-node N tools/sharded_check.py [--reads R]: parity of the bucket-sharded multi-GPU path
against the oracle run on the union of all ranks' reads, then a timing pass at --bench-reads per rank."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker only)
import rust_debruijn_b200 as D  # noqa: E402
from rust_debruijn_b200 import sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=40000)
ap.add_argument("--bench-reads", type=int, default=0)
ap.add_argument("--k", type=int, default=31)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = D.Context(local)
ok = True
for k, noisy, mo in ((a.k, True, 2), (a.k, False, 1), (63, True, 2)):
    R = a.reads
    words, start, length = O.synth_reads(R, 1, O.ERR_THR_NOISY if noisy else 0)
    lo_r, hi_r = rank * R // world, (rank + 1) * R // world
    bases = np.concatenate([O.unpack_bases(words, i * 150, 150) for i in range(lo_r, min(hi_r, lo_r + 0))]) if False else None
    # slice this rank's reads out of the packed global set (150 bases each, contiguous)
    allb = O.unpack_bases(words, lo_r * 150, (hi_r - lo_r) * 150)
    w_loc = O.pack_bases(allb)
    ss = D.SeqSet.upload_uniform(ctx, w_loc, hi_r - lo_r, 150)
    tm = {}
    g = sharded.reads_to_graph_sharded(ss, D.CountFilter(mo), D.SimpleCompress(D.SAT_ADD), stranded=False, k=k, timings=tm)
    gh = g.to_host()
    ot = O.filter_kmers(k, words, start, length, min_obs=mo)
    og = O.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"])
    same = all(np.array_equal(gh[f], og[f]) for f in ("words", "start", "length", "exts", "data"))
    # node-sharded output: every rank keeps its run of nodes; the runs concatenated in rank order must be the same graph
    g2 = sharded.reads_to_graph_sharded(ss, D.CountFilter(mo), D.SimpleCompress(D.SAT_ADD), stranded=False, k=k, replicate=False)
    h2 = g2.to_host()
    part = dict(replicated=g2.replicated, node0=g2.node0, base0=g2.base0, n_bases=h2["n_bases"], start=h2["start"], length=h2["length"],
                exts=h2["exts"], data=h2["data"], bases=O.unpack_bases(h2["words"], 0, h2["n_bases"]))
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0)
    same2 = True
    if rank == 0:
        if parts[0]["replicated"]:
            parts = parts[:1]
        cat = {f: np.concatenate([p_[f] for p_ in parts]) for f in ("length", "exts", "data", "bases")}
        cat["start"] = np.concatenate([p_["start"] + np.uint64(p_["base0"]) for p_ in parts])
        same2 = (all(np.array_equal(cat[f], og[f]) for f in ("start", "length", "exts", "data")) and
                 np.array_equal(O.pack_bases(cat["bases"]), og["words"]) and
                 [p_["node0"] for p_ in parts] == list(np.cumsum([0] + [len(p_["length"]) for p_ in parts[:-1]])))
    ok &= same and same2
    if rank == 0:
        print(f"[sharded_check] world={world} k={k} noisy={noisy}: nodes={gh['n_nodes']} oracle={og['n_nodes']} "
              f"{'BIT-EXACT' if same else 'MISMATCH'} node-sharded {'BIT-EXACT' if same2 else 'MISMATCH'}  {tm}", flush=True)
flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
dist.all_reduce(flag)
if a.bench_reads:
    ss = D.SeqSet.synth(ctx, a.bench_reads, 1 + rank, O.ERR_THR_NOISY)
    for it in range(3):
        tm = {}
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = sharded.reads_to_graph_sharded(ss, D.CountFilter(2), D.SimpleCompress(D.SAT_ADD), k=a.k, timings=tm)
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"[sharded_check] bench it={it} ms={e0.elapsed_time(e1):.2f} nodes={len(g)} {tm}", flush=True)
        g.free()
dist.destroy_process_group()
sys.exit(int(flag.item()) != 0)
