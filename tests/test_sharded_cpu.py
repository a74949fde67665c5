"""world_size-2/3 gloo tests (CPU) of the host-side logic of the multi-GPU path: bucket ownership, the per-bucket count
exchange, and the quantile cuts that route path records to the rank owning their seed's key range.  The planning functions
are the library's own (dbg_plan_owner_bounds / dbg_plan_quantile_cuts through ctypes: pure host code, no GPU needed).  The
data itself moves on GPUs: tests/test_gpu_parity.py::test_multi_rank_path_on_one_gpu (several ranks on one device) and
::test_multi_gpu_nccl (torchrun, NCCL + CUDA IPC), both through tools/multi_check.py."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nb, out):
    import torch.distributed as dist

    from rust_debruijn_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    counts = rng.integers(0, 50, size=nb).astype(np.uint32)      # this rank's per-bucket record counts
    per_dst = sharded.split_by_owner(counts, world)
    recv = sharded.exchange_counts(per_dst, device="cpu")
    out[rank] = (counts, [r.copy() for r in recv])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nb", [(2, 16), (2, 2), (3, 8)])
def test_count_exchange_gloo(world, nb):
    import torch.multiprocessing as mp

    from rust_debruijn_b200 import sharded
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, nb, out), nprocs=world, join=True)
    bounds = sharded.owner_bounds(nb, world)
    assert bounds[0] == 0 and bounds[-1] == nb and all(bounds[i] <= bounds[i + 1] for i in range(world))
    for dst in range(world):
        _, recv = out[dst]
        for src in range(world):
            sent = out[src][0][bounds[dst]:bounds[dst + 1]]
            assert np.array_equal(recv[src], sent), (src, dst)
    # every bucket has exactly one owner and nothing is lost
    total_sent = sum(int(out[r][0].sum()) for r in range(world))
    total_recv = sum(int(sum(x.sum() for x in out[r][1])) for r in range(world))
    assert total_sent == total_recv


def test_owner_bounds_properties():
    """bounds agree with the device-side owner function owner(b) = (b * P) >> bits, every rank owns >= 1 bucket."""
    from rust_debruijn_b200 import sharded
    for world in (1, 2, 3, 4, 5, 8):
        assert sharded.min_bucket_bits(world) == (int(np.ceil(np.log2(world))) if world > 1 else 0)
        for bits in range(sharded.min_bucket_bits(world), 12):
            b = sharded.owner_bounds(1 << bits, world)
            sizes = np.diff(b)
            assert b[0] == 0 and b[-1] == 1 << bits and sizes.min() >= 1 and sizes.max() - sizes.min() <= 1
            owner = (np.arange(1 << bits, dtype=np.uint64) * np.uint64(world)) >> np.uint64(bits)
            for r in range(world):
                assert np.all(owner[b[r]:b[r + 1]] == r)


def test_key_range_splitters():
    from rust_debruijn_b200 import sharded
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for hist in (rng.integers(0, 100, size=65536), np.zeros(65536, np.int64), np.eye(1, 65536, 7, dtype=np.int64)[0] * 1000):
            cuts = sharded.key_range_splitters(hist, world)
            assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == 65536
            assert all(cuts[i] <= cuts[i + 1] for i in range(world))
            if hist.sum() and world > 1 and hist.max() < hist.sum() / world:
                mass = [hist[cuts[i]:cuts[i + 1]].sum() for i in range(world)]
                assert max(mass) <= hist.sum() / world + hist.max() + 1


def _seed_worker(rank, world, port, bits, n, out):
    import torch
    import torch.distributed as dist

    from rust_debruijn_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7 + rank)
    # seed k-mers (62-bit keys) crowd the low values (a seed is the minimum of its unitig), different amounts per rank
    keys = np.sort(np.minimum.reduce(rng.integers(0, 1 << 62, size=(4, n + 100 * rank), dtype=np.int64), axis=0))
    hist_local = np.bincount((keys >> (62 - bits)).astype(np.int64), minlength=1 << bits).astype(np.int64)
    hg = torch.from_numpy(hist_local.copy())
    dist.all_reduce(hg)                                             # the library: ncclAllReduce of the same histogram
    cuts = sharded.key_range_splitters(hg.numpy(), world)           # identical on every rank
    bnd = sharded.local_bounds(hist_local, cuts)
    out[rank] = (keys, bnd, cuts)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_seed_key_ranges_gloo(world):
    """Seed-key ranges of the sharded compression: contiguous slices, the same cuts on every rank, destinations balanced by node
    count although the seeds are far from uniform, ranges ordered (so that per-rank runs concatenate to the global node order)."""
    import torch.multiprocessing as mp
    bits, n = 16, 20_000
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_seed_worker, args=(world, port, bits, n, out), nprocs=world, join=True)
    per_dst = np.zeros(world, np.int64)
    hi = [-1] * world
    lo = [1 << 62] * world
    for r in range(world):
        keys, bnd, cuts = out[r]
        assert cuts == out[0][2]
        assert bnd[0] == 0 and bnd[-1] == len(keys) and all(bnd[i] <= bnd[i + 1] for i in range(world))
        for d in range(world):
            sl = keys[bnd[d]:bnd[d + 1]]
            per_dst[d] += len(sl)
            if len(sl):
                hi[d], lo[d] = max(hi[d], int(sl.max())), min(lo[d], int(sl.min()))
    for d in range(1, world):
        assert lo[d] > max(hi[:d])
    total = per_dst.sum()
    assert per_dst.max() - per_dst.min() <= 0.1 * total / world + 4096   # quantile cuts at 2^16-bin granularity


def _layout_worker(rank, world, port, nb, out):
    import torch
    import torch.distributed as dist

    from rust_debruijn_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(500 + rank)
    counts = rng.integers(0, 40, size=nb).astype(np.int32)
    counts[rng.integers(0, nb, size=max(1, nb // 4))] = 0                 # empty buckets on some ranks
    gathered = [torch.zeros(nb, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(counts))                   # the library all-gathers the same matrix (NCCL, on the device)
    allc = np.stack([g.numpy() for g in gathered]).astype(np.uint32)
    dst_off, recv_total = sharded.exchange_layout(allc, rank)
    out[rank] = (counts.astype(np.uint32), dst_off, recv_total)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nb", [(2, 64), (3, 50), (2, 2)])
def test_fused_exchange_layout_gloo(world, nb):
    """The layout every sender derives from the all-gathered per-bucket counts (dbg_plan_exchange_layout, the host restatement of what
    bucket_totals_kernel + the scan + scatter_buckets_kernel compute): all ranks agree on the receive totals, and the senders' chunks
    tile every owner's window exactly — buckets contiguous and ascending, senders in rank order inside a bucket, no gap, no overlap."""
    import torch.multiprocessing as mp

    from rust_debruijn_b200 import sharded
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_layout_worker, args=(world, port, nb, out), nprocs=world, join=True)
    bounds = sharded.owner_bounds(nb, world)
    for r in range(1, world):
        assert np.array_equal(out[r][2], out[0][2])                       # everybody computed the same receive totals
    for dst in range(world):
        total = int(out[0][2][dst])
        assert total == sum(int(out[s][0][bounds[dst]:bounds[dst + 1]].sum()) for s in range(world))
        cover = np.zeros(total, np.int32)
        pos = 0
        for b in range(bounds[dst], bounds[dst + 1]):
            for s in range(world):                                        # expected order: bucket-major, sender-minor
                n, off = int(out[s][0][b]), int(out[s][1][b])
                if n:
                    assert off == pos, (dst, b, s)
                cover[off:off + n] += 1
                pos += n
        assert pos == total and (cover == 1).all()
