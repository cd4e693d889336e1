"""CPU tests of the multi-GPU host logic: partitioning math and the world_size-2 gloo reduction of metric sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nncf_b200.parallel import (allreduce_metric_sums, held_item_shard, local_row, n_item_strata, owner_of,
                                partition_links_by_block, partition_links_by_stratum, ring_rotate, shard_range, shard_rows,
                                slot_of, stratum_of)


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


# ---------------------------------------------------------------------------------------------------------------
# pipelined stratified schedule (2N item strata, three rotating buffers per rank)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_pipelined_schedule_invariants(world):
    """per phase the ranks train distinct strata; over an epoch of M phases every rank trains each stratum once; the
    stratum a rank trains next is the one rank + 1 trained one phase earlier (so it can travel during the phase in
    between) and nobody trains it meanwhile"""
    m = n_item_strata(world)
    assert m == (2 * world if world > 1 else 1)
    for p in range(3 * m):
        now = [stratum_of(r, p, world) for r in range(world)]
        assert len(set(now)) == world
        if world > 1:
            for r in range(world):
                nxt = stratum_of(r, p + 1, world)
                assert nxt not in now
                if p >= 1:
                    assert nxt == stratum_of((r + 1) % world, p - 1, world)
    for r in range(world):
        assert sorted(stratum_of(r, p, world) for p in range(m)) == list(range(m))
        assert sorted(stratum_of(r, p, world) for p in range(m, 2 * m)) == list(range(m))      # and again in the next epoch


def test_pipelined_buffer_roles_never_collide():
    """three buffers rotate through train / incoming / outgoing: in phase p the trained slot, the slot being filled for
    p + 1 and the slot being sent (trained in p - 1) are distinct, and the slot a sender writes at phase p + 1 is the one
    the receiver finished sending at phase p (what the credit flag guards)"""
    for p in range(1, 40):
        assert len({slot_of(p), slot_of(p + 1), slot_of(p - 1)}) == 3
        assert slot_of(p + 2) == slot_of(p - 1)


@pytest.mark.parametrize("world", [2, 3])
def test_partition_links_by_stratum_covers_every_link_once(world):
    rng = np.random.RandomState(0)
    m = n_item_strata(world)
    train = np.stack([rng.randint(0, 101, 5000), rng.randint(0, 57, 5000), np.ones(5000, dtype=np.int64)], 1)
    total = 0
    for r in range(world):
        blocks = partition_links_by_stratum(train, r, world)
        assert len(blocks) == m
        for s_, b in enumerate(blocks):
            assert b.dtype == np.int32
            g = np.stack([b[:, 0].astype(np.int64) * world + r, b[:, 1].astype(np.int64) * m + s_, b[:, 2]], 1)
            ref = train[(train[:, 0] % world == r) & (train[:, 1] % m == s_)]
            assert np.array_equal(g, ref)                       # order inside a block is the original order
            if len(b):
                assert b[:, 1].max() < shard_rows(57, s_, m)
            total += len(b)
    assert total == len(train)


def _pipeline_worker(rank, world, port, q):
    """the rotation protocol with CPU tensors over gloo: at the end of phase p - 1 a rank posts the send of the stratum it
    has just trained (slot (p - 1) % 3 -> rank - 1's slot (p + 1) % 3) and the matching receive, and only waits for them
    at the end of phase p - one full phase later, as the device version does with its flags"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = n_item_strata(world)
    slots = [torch.full((4,), -1.0) for _ in range(3)]
    slots[0].fill_(float(stratum_of(rank, 0, world)))
    slots[1].fill_(float(stratum_of(rank, 1, world)))
    trace, pending = [], []
    for p in range(2 * m):                                   # two epochs
        cur = slots[slot_of(p)]
        trace.append(int(cur[0].item()))
        cur += 100.0                                         # "training" marks the stratum: +100 per visit
        for req in pending:                                  # transfers posted one phase ago must have landed by now
            req.wait()
        nxt = p + 1                                          # advance to phase nxt: send slot (nxt - 1) % 3
        send = slots[slot_of(nxt - 1)]
        recv = slots[slot_of(nxt + 1)]
        pending = [dist.isend(send, (rank - 1) % world), dist.irecv(recv, (rank + 1) % world)]
    for req in pending:
        req.wait()
    q.put((rank, trace))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pipelined_rotation_data_flow_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = n_item_strata(world)
    for rank, trace in res:
        # what a rank finds in its training slot: the right stratum, carrying one mark per earlier visit by ANY rank
        exp = []
        for p in range(2 * m):
            s_ = stratum_of(rank, p, world)
            visits = sum(1 for pp in range(p) for r in range(world) if stratum_of(r, pp, world) == s_)
            exp.append(s_ + 100 * visits)
        assert trace == exp
