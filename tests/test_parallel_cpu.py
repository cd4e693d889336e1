"""CPU tests of the multi-GPU host logic: partitioning math and the world_size-2 gloo reduction of metric sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nncf_b200.parallel import (allreduce_metric_sums, held_item_shard, local_row, owner_of, partition_links_by_block,
                                ring_rotate, shard_range, shard_rows)


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (10, 3), (7, 8), (1000003, 8), (0, 4)])
def test_shard_range_partitions_without_gaps(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for a, b in zip(spans, spans[1:]):
        assert a[1] == b[0]
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_rows,world", [(10, 2), (1003, 2), (1000000, 8), (5, 8)])
def test_row_sharding_is_a_bijection(n_rows, world):
    ids = np.arange(n_rows)
    own, loc = owner_of(ids, world), local_row(ids, world)
    assert np.array_equal(loc * world + own, ids)
    for r in range(world):
        assert shard_rows(n_rows, r, world) == int(np.sum(own == r))
        if np.any(own == r):
            assert loc[own == r].max() == shard_rows(n_rows, r, world) - 1      # local rows are dense 0..rows_local-1
    assert sum(shard_rows(n_rows, r, world) for r in range(world)) == n_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank evaluates its own user range; per-user metrics are a deterministic function of the user id
    n_users = 1001
    lo, hi = shard_range(n_users, rank, world)
    u = np.arange(lo, hi, dtype=np.float64)
    kept = (u % 7 != 0)
    sums = torch.tensor([np.sum(kept * (u % 5) / 5.0), np.sum(kept * (u % 3) / 3.0), np.sum(kept * 0.1), kept.sum()],
                        dtype=torch.float64)
    out = allreduce_metric_sums(sums)
    q.put((rank, out.numpy().tolist()))
    dist.destroy_process_group()


def test_metric_sums_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    u = np.arange(1001, dtype=np.float64)
    kept = (u % 7 != 0)
    exp = [np.sum(kept * (u % 5) / 5.0), np.sum(kept * (u % 3) / 3.0), np.sum(kept * 0.1), kept.sum()]
    for r in range(world):
        np.testing.assert_allclose(res[r], exp, rtol=1e-12)


# ---------------------------------------------------------------------------------------------------------------
# stratified (DSGD) schedule
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_stratified_schedule_is_a_latin_square(world):
    """in every sub-epoch the ranks hold distinct item shards, and over `world` sub-epochs every rank holds each once"""
    for t in range(world):
        assert sorted(held_item_shard(r, t, world) for r in range(world)) == list(range(world))
    for r in range(world):
        assert sorted(held_item_shard(r, t, world) for t in range(world)) == list(range(world))


@pytest.mark.parametrize("world", [2, 3])
def test_partition_links_by_block_covers_every_link_once(world):
    rng = np.random.RandomState(0)
    train = np.stack([rng.randint(0, 101, 5000), rng.randint(0, 57, 5000), np.ones(5000, dtype=np.int64)], 1)
    seen = []
    for r in range(world):
        blocks = partition_links_by_block(train, r, world)
        assert len(blocks) == world
        for v, b in enumerate(blocks):
            assert b.dtype == np.int32
            g = np.stack([b[:, 0].astype(np.int64) * world + r, b[:, 1].astype(np.int64) * world + v, b[:, 2]], 1)   # back to global ids
            assert np.all(g[:, 0] % world == r) and np.all(g[:, 1] % world == v)
            ref = train[(train[:, 0] % world == r) & (train[:, 1] % world == v)]
            assert np.array_equal(g, ref)                       # order inside a block is the original order
            seen.append(g)
    assert sum(len(g) for g in seen) == len(train)


def _rotate_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    held = [torch.full((4, 3), float(rank)), torch.full((2,), 10.0 + rank)]       # shard `rank` and its "optimizer state"
    spare = [torch.empty_like(t) for t in held]
    trace = []
    for t in range(world):
        trace.append((int(held[0][0, 0].item()), int(held[1][0].item()) - 10))
        held, spare = ring_rotate(held, spare, rank, world)
    q.put((rank, trace, int(held[0][0, 0].item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_rotate_follows_the_schedule_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rotate_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, trace, final in res:
        assert trace == [(held_item_shard(rank, t, world),) * 2 for t in range(world)]       # tensors travel together
        assert final == rank                                                                   # home again after N rotations
