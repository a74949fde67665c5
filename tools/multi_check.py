"""Parity of the multi-GPU path (dbg_reads_to_graph_multi / dbg_multi_reads_to_graph) against the oracle run on the union of
all ranks' reads: the per-rank runs of nodes, concatenated in rank order, must be the oracle's BaseGraph bit for bit.

    python tools/multi_check.py --local 2 [--reads R]                  # 2 ranks on ONE GPU, one process ("local" transport)
    torchrun --nproc-per-node N tools/multi_check.py [--reads R]       # one process per GPU (NCCL + CUDA IPC)
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (checker only)
import rust_debruijn_b200 as D  # noqa: E402
from rust_debruijn_b200 import multi  # noqa: E402


def split_reads(words, R, world):
    """word-aligned slices of a uniform 150-bp read set: rank r gets reads [b[r], b[r+1]) with b multiples of 16"""
    b = [((r * R // world) // 16) * 16 for r in range(world)] + [R]
    return [(words[b[r] * 150 // 32:(b[r + 1] * 150 + 31) // 32], b[r + 1] - b[r]) for r in range(world)]


def concat_runs(parts):
    """parts: per rank dict(node0, base0, replicated, host graph) -> complete graph arrays"""
    if parts[0]["replicated"]:
        parts = parts[:1]
    out = {f: np.concatenate([p["g"][f] for p in parts]) for f in ("length", "exts", "data")}
    out["start"] = np.concatenate([p["g"]["start"] + np.uint64(p["base0"]) for p in parts])
    bases = np.concatenate([O.unpack_bases(p["g"]["words"], 0, p["g"]["n_bases"]) for p in parts])
    out["words"] = O.pack_bases(bases)
    out["node0_ok"] = [p["node0"] for p in parts] == list(np.cumsum([0] + [len(p["g"]["length"]) for p in parts[:-1]]))
    return out


def same_graph(cat, og):
    return all(np.array_equal(cat[f], og[f]) for f in ("start", "length", "exts", "data", "words")) and cat["node0_ok"]


CONFIGS = [  # (k, noisy, min_obs, stranded, reduce_op, params)
    (31, True, 2, False, 0, {}),
    (31, False, 1, False, 0, {}),            # one genome-long unitig: the replicated fallback
    (63, True, 2, False, 0, {}),
    (32, True, 2, True, 3, {}),
    (31, True, 2, False, 0, {"bucket_occ": 16}),   # many buckets: minimizer length follows the bucket count (p >= 13)
    (31, True, 2, False, 4, {}),             # ScmapCompress: join_test crosses ranks
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local", type=int, default=0, help="N ranks on device 0 in this process (local transport)")
    ap.add_argument("--reads", type=int, default=40000)
    ap.add_argument("--big-reads", type=int, default=0, help="one extra K=31 noisy configuration of this many reads in total")
    a = ap.parse_args()
    ok = True
    cfgs = [(c, a.reads) for c in CONFIGS] + ([((31, True, 2, False, 0, {}), a.big_reads)] if a.big_reads else [])
    if a.local:
        world = a.local
        mc = multi.MultiContext([0] * world)
        for (k, noisy, mo, stranded, op, params), R in cfgs:
            words, start, length = O.synth_reads(R, 1, O.ERR_THR_NOISY if noisy else 0)
            for c in mc.ctxs:
                for n_, v_ in (("bucket_occ", 0), ("msp_p", 0)):
                    c.set_param(n_, params.get(n_, v_))
            sss = [D.SeqSet.upload_uniform(mc.ctxs[r], w, n, 150) for r, (w, n) in enumerate(split_reads(words, R, world))]
            spec = D.ScmapCompress() if op == 4 else D.SimpleCompress(op)
            gs = mc.reads_to_graph(sss, D.CountFilter(mo), spec, stranded=stranded, k=k)
            parts = [dict(node0=g.node0, base0=g.base0, replicated=g.replicated, g=g.to_host()) for g in gs]
            ot = O.filter_kmers(k, words, start, length, min_obs=mo, stranded=stranded, threads=os.cpu_count() or 1)
            og = O.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"], stranded=stranded, reduce_op=op)
            same = same_graph(concat_runs(parts), og) and all(g.invariants["ok"] for g in gs) and gs[0].n_valid_total == len(ot["lo"])
            ok &= bool(same)
            i0 = gs[0].info
            print(f"[multi_check] transport={mc.transport} world={world} k={k} noisy={noisy} stranded={stranded} op={op} {params} R={R}: "
                  f"nodes={gs[0].n_nodes_total} oracle={og['n_nodes']} replicated={gs[0].replicated} p={i0['msp_p']} bits={i0['bucket_bits']} "
                  f"queries={[g.info['n_queries_sent'] for g in gs]} valid={[g.info['n_valid_local'] for g in gs]} "
                  f"{'BIT-EXACT' if same else 'MISMATCH'}", flush=True)
            for g in gs:
                g.free()
            for s in sss:
                s.free()
        mc.close()
    else:
        import torch
        import torch.distributed as dist
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = D.Context(local)
        comm = multi.Comm.from_torch(ctx)
        for (k, noisy, mo, stranded, op, params), R in cfgs:
            words, start, length = O.synth_reads(R, 1, O.ERR_THR_NOISY if noisy else 0)
            for n_, v_ in (("bucket_occ", 0), ("msp_p", 0)):
                ctx.set_param(n_, params.get(n_, v_))
            w, n = split_reads(words, R, world)[rank]
            ss = D.SeqSet.upload_uniform(ctx, w, n, 150)
            spec = D.ScmapCompress() if op == 4 else D.SimpleCompress(op)
            g = comm.reads_to_graph(ss, D.CountFilter(mo), spec, stranded=stranded, k=k)
            part = dict(node0=g.node0, base0=g.base0, replicated=g.replicated, g=g.to_host(), inv=g.invariants, info=g.info)
            parts = [None] * world if rank == 0 else None
            dist.gather_object(part, parts, dst=0)
            same = True
            if rank == 0:
                ot = O.filter_kmers(k, words, start, length, min_obs=mo, stranded=stranded, threads=os.cpu_count() or 1)
                og = O.compress_kmers(k, ot["lo"], ot["hi"], ot["exts"], ot["counts"], stranded=stranded, reduce_op=op)
                same = same_graph(concat_runs(parts), og) and all(p_["inv"]["ok"] for p_ in parts) and g.n_valid_total == len(ot["lo"])
                print(f"[multi_check] transport={comm.transport} world={world} k={k} noisy={noisy} stranded={stranded} op={op} {params} R={R}: "
                      f"nodes={g.n_nodes_total} oracle={og['n_nodes']} replicated={g.replicated} p={g.info['msp_p']} bits={g.info['bucket_bits']} "
                      f"queries={[p_['info']['n_queries_sent'] for p_ in parts]} {'BIT-EXACT' if same else 'MISMATCH'} "
                      f"ms={ {k_: round(v, 2) for k_, v in g.info.items() if k_.startswith('ms_')} }", flush=True)
            ok &= bool(same)
            g.free()
            ss.free()
        flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
        dist.all_reduce(flag)
        ok = int(flag.item()) == 0
        comm.close()
        dist.destroy_process_group()
    print("MULTI OK" if ok else "MULTI FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
