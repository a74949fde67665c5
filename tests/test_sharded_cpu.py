"""world_size-2 gloo tests (CPU) of the host-side logic of the multi-GPU path: bucket ownership, the
per-bucket count exchange, and the offsets dbg_filter_from_records derives from it.  The records themselves
move on GPUs (tests/test_gpu_parity.py::test_sharded_two_gpus and tools/sharded_check.py)."""
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


@pytest.mark.parametrize("world,nb", [(2, 16), (2, 1), (3, 8)])
def test_count_exchange_gloo(world, nb):
    import torch.multiprocessing as mp

    from rust_debruijn_b200 import sharded
    if nb < world:
        nb = world
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
    from rust_debruijn_b200 import sharded
    for world in (1, 2, 3, 4, 8):
        assert sharded.min_bucket_bits(world) == int(np.ceil(np.log2(world))) if world > 1 else True
        for bits in range(sharded.min_bucket_bits(world), 12):
            b = sharded.owner_bounds(1 << bits, world)
            sizes = np.diff(b)
            assert sizes.min() >= 1 and sizes.max() - sizes.min() <= 1


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


def _seed_worker(rank, world, port, V, n, out):
    import torch
    import torch.distributed as dist

    from rust_debruijn_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7 + rank)
    # seeds crowd the low indices (minimum of a few uniform draws), different amounts per rank
    seeds = np.sort(np.minimum.reduce(rng.integers(0, V, size=(4, n + 100 * rank)), axis=0)).astype(np.int64)
    bnd = sharded.balanced_seed_bounds(torch.from_numpy(seeds), V, world)
    out[rank] = (seeds, bnd.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_balanced_seed_bounds_gloo(world):
    """Seed-range splitters of the sharded compression: contiguous slices, the same seed thresholds on every rank,
    destinations balanced by node count although the seeds are far from uniform."""
    import torch.multiprocessing as mp
    V, n = 1_000_000, 20_000
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_seed_worker, args=(world, port, V, n, out), nprocs=world, join=True)
    per_dst = np.zeros(world, np.int64)
    hi_prev = [-1] * world   # largest seed sent to each destination so far / smallest sent to the next
    for r in range(world):
        seeds, bnd = out[r]
        assert bnd[0] == 0 and bnd[-1] == len(seeds) and np.all(np.diff(bnd) >= 0)
        for d in range(world):
            sl = seeds[bnd[d]:bnd[d + 1]]
            per_dst[d] += len(sl)
            if len(sl):
                hi_prev[d] = max(hi_prev[d], int(sl.max()))
    # ranges are disjoint and ordered across ranks: everything sent to d is below everything sent to d + 1
    for r in range(world):
        seeds, bnd = out[r]
        for d in range(1, world):
            sl = seeds[bnd[d]:bnd[d + 1]]
            if len(sl):
                assert int(sl.min()) > max(hi_prev[:d])
    total = per_dst.sum()
    assert per_dst.max() - per_dst.min() <= 0.1 * total / world + 4096   # quantile cuts at 4096-bin granularity
