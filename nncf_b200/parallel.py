"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous and scalar reductions.

* whole@k evaluation shards by USER with no data-path collective (`shard_range`, `allreduce_metric_sums`).
* training shards the embedding tables by ROW (`owner_of`, `local_row`, `ShardedTable`): row `id` lives on rank
  `id % N` at local row `id // N`.  Shards are peer-visible allocations (CUDA IPC); the step kernels read and update
  remote rows directly over NVLink (include/nncf_b200.h, section 3b) — no all-to-all staging buffers.

* `StratifiedTrainer`: the schedule that scales.  The step moves ~2 KB of embedding rows per positive link; at the
  single-GPU rate (7e8 links/s) that is 1.4 TB/s of random row traffic per GPU, more than NVLink 5 carries (0.9 TB/s per
  direction), so ANY scheme that fetches rows from their owners per step is NVLink-bound (measured with the peer-memory
  mode: 5.3e8 links/s on 2 GPUs against 7.2e8 on one).  Stratified SGD (Gemulla et al., KDD'11, "DSGD") removes the
  per-step exchange: links are partitioned into N x N blocks by (user % N, item % N); rank r owns user shard r for good
  and, in sub-epoch t, holds item shard (r + t) % N and trains ONLY on block (r, (r + t) % N), so every row it touches
  is local and the single-GPU kernels run unchanged; between sub-epochs the item shards (and their optimizer state)
  rotate one rank down the ring (one NCCL send/recv of n_items / N rows per rank).  The blocks trained concurrently are
  disjoint in users AND items, so a sub-epoch equals a sequential pass over its N blocks in any order (checked against
  the oracle on 2 GPUs, tools/multi_gpu_check.py).  Declared difference to a single-GPU epoch: a batch's shared
  negatives come from the item shard of its block (a random 1/N of the items under id % N).

The reference is single-process / single-device (no collective anywhere); everything here is the
B200-native extension and is exercised on CPU with the gloo backend for the host-side logic (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np


# ------------------------------------------------------------------------------------------------------------------
# pure partitioning math (CPU-testable)
# ------------------------------------------------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) range of `n` units owned by `rank`: sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(ids, world: int):
    return np.asarray(ids) % world


def local_row(ids, world: int):
    return np.asarray(ids) // world


def shard_rows(n_rows: int, rank: int, world: int) -> int:
    """number of rows of a table with n_rows rows that live on `rank` under owner = id % world"""
    return (n_rows - rank + world - 1) // world


def held_item_shard(rank: int, sub_epoch: int, world: int) -> int:
    """item shard held (and trained on) by `rank` during sub-epoch `sub_epoch` of the stratified schedule"""
    return (rank + sub_epoch) % world


def partition_links_by_block(train, rank: int, world: int):
    """Rows of `train` (int [n, 3]: user, item, label) owned by `rank` (user % world == rank), split by item shard.
    Returns a list of `world` int32 arrays [n_v, 3] whose ids are LOCAL rows (id // world) of the two shards; the
    original order is kept inside every block."""
    train = np.asarray(train)
    mine = train[train[:, 0] % world == rank]
    out = []
    for v in range(world):
        b = mine[mine[:, 1] % world == v].astype(np.int32, copy=True)
        b[:, 0] //= world
        b[:, 1] //= world
        out.append(b)
    return out


def ring_rotate(tensors, spares, rank: int, world: int, group=None):
    """Every rank sends each tensor of `tensors` to rank - 1 and receives rank + 1's into the matching tensor of
    `spares` (same shapes on every rank), then the two lists are swapped: after the call `tensors` holds what rank + 1
    held.  NCCL on GPUs (ordered with the current stream), gloo on CPU tensors (tests)."""
    import torch.distributed as dist
    if world == 1:
        return tensors, spares
    dst, src = (rank - 1) % world, (rank + 1) % world
    ops_ = []
    for t, sp in zip(tensors, spares):
        ops_.append(dist.P2POp(dist.isend, t, dst, group))
        ops_.append(dist.P2POp(dist.irecv, sp, src, group))
    for req in dist.batch_isend_irecv(ops_):
        req.wait()
    return spares, tensors


def allreduce_metric_sums(sums, group=None):
    """Sum-reduce the evaluator's (sum AP, sum recall, sum precision, users kept) over ranks.  `sums` is a torch
    tensor (CUDA with nccl, CPU with gloo); returns the reduced tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def sharded_whole_eval(user_rows_fn, item_rows, truth_indptr, truth_cols, n_users, topk, precision, rank, world):
    """whole@k with users sharded across ranks.  user_rows_fn(lo, hi) -> CUDA float32 [hi-lo, d] for this rank's users;
    item_rows is the (replicated) candidate matrix; truth is the full CSR on the host (NumPy).  Returns the global means."""
    import torch
    from . import ops
    lo, hi = shard_range(n_users, rank, world)
    sums = torch.zeros(4, dtype=torch.float64, device=item_rows.device)
    if hi > lo:
        U = user_rows_fn(lo, hi)
        ids, _ = ops.eval_topk(U, item_rows, topk, precision)
        ip = truth_indptr[lo:hi + 1] - truth_indptr[lo]
        cols = truth_cols[truth_indptr[lo]:truth_indptr[hi]]
        _, sums = ops.eval_metrics(ids, torch.from_numpy(np.ascontiguousarray(ip)).to(U.device),
                                   torch.from_numpy(np.ascontiguousarray(cols)).to(U.device))
    sums = allreduce_metric_sums(sums)
    s = sums.cpu().numpy()
    n = max(s[3], 1.0)
    return {"map": s[0] / n, "recall": s[1] / n, "precision": s[2] / n, "n_users": int(s[3])}


# ------------------------------------------------------------------------------------------------------------------
# peer-visible row shards
# ------------------------------------------------------------------------------------------------------------------
class _CudaBuffer:
    """exposes a raw device allocation to torch through __cuda_array_interface__"""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerMemory:
    """One peer-visible allocation per rank + the N opened views (rank r's own pointer sits at index r)."""

    def __init__(self, nbytes: int, rank: int, world: int):
        import torch.distributed as dist
        from ._lib import check, lib
        self.rank, self.world = rank, world
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        check(lib.nncf_peer_alloc(int(nbytes), C.byref(ptr), handle))
        self.local_ptr = int(ptr.value)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        self.ptrs = []
        self._opened = []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.local_ptr)
            else:
                p = C.c_void_p()
                h = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                check(lib.nncf_peer_open(h, C.byref(p)))
                self.ptrs.append(int(p.value))
                self._opened.append(int(p.value))
        self.ptr_array = (C.c_void_p * world)(*self.ptrs)

    def close(self):
        from ._lib import lib
        for p in self._opened:
            lib.nncf_peer_close(C.c_void_p(p))
        self._opened = []
        if self.local_ptr:
            lib.nncf_peer_free(C.c_void_p(self.local_ptr))
            self.local_ptr = 0


class ShardedTable:
    """A [n_rows, dim] fp32 table sharded by row over the ranks (owner = id % world)."""

    def __init__(self, n_rows: int, dim: int, rank: int, world: int):
        import torch
        self.n_rows, self.dim, self.rank, self.world = n_rows, dim, rank, world
        self.rows_local = shard_rows(n_rows, rank, world)
        self.mem = PeerMemory(max(self.rows_local, 1) * dim * 4, rank, world)
        self.local = torch.as_tensor(_CudaBuffer(self.mem.local_ptr, (max(self.rows_local, 1), dim)), device="cuda")

    def init_uniform(self, scale: float, seed: int):
        import torch
        g = torch.Generator(device="cuda").manual_seed(seed + 1000 * self.rank)
        self.local.copy_((torch.rand(self.local.shape, device="cuda", generator=g) - 0.5) * (2 * scale))

    def gather_global(self, ids):
        """host helper for tests: rows of global ids from this rank's view of all shards (peer reads through torch)"""
        import torch
        out = []
        for i in np.asarray(ids).tolist():
            r, l = i % self.world, i // self.world
            n_r = shard_rows(self.n_rows, r, self.world)
            view = torch.as_tensor(_CudaBuffer(self.mem.ptrs[r], (max(n_r, 1), self.dim)), device="cuda")
            out.append(view[l].clone())
        return torch.stack(out)


class ShardedTrainer:
    """Row-sharded neg_shared / group_neg_shared training step over `world` GPUs (sparse SGD)."""

    def __init__(self, spec, n_users: int, n_items: int, rank: int, world: int, seed: int = 7):
        from ._lib import check, lib
        from .ops import FusedStep
        assert spec.optimizer in ("sgd", "none"), "sharded tables: sparse SGD only"
        self.spec, self.rank, self.world = spec, rank, world
        self.users = ShardedTable(n_users, spec.dim, rank, world)
        self.items = ShardedTable(n_items, spec.dim, rank, world)
        self.users.init_uniform(0.05, seed)
        self.items.init_uniform(0.05, seed + 1)
        self.flags = PeerMemory(64, rank, world)
        self.step = FusedStep(spec)
        check(lib.nncf_trainer_set_shards(self.step._h, world, rank, self.users.mem.ptr_array, self.items.mem.ptr_array,
                                          self.flags.ptr_array))

    def run(self, user_ids, item_ids, n_steps, loss_out=None):
        """ids are GLOBAL ids; every rank must call with the same n_steps (the device barriers pair up)."""
        return self.step.run(self.users.local, self.items.local, user_ids, item_ids, n_steps, loss_out=loss_out)

    def close(self):
        self.step = None
        for m in (self.users.mem, self.items.mem, self.flags):
            m.close()


class StratifiedTrainer:
    """Stratified (DSGD-style) multi-GPU training: rank r keeps user shard r, item shards rotate round the ring between
    sub-epochs, every step touches local rows only (see the module docstring).  Works with sparse SGD and lazy Adam
    (the optimizer state of the item shard travels with it)."""

    def __init__(self, spec, n_users: int, n_items: int, rank: int, world: int, seed: int = 7):
        import torch
        from .ops import FusedStep
        self.spec, self.rank, self.world = spec, rank, world
        self.n_users, self.n_items = n_users, n_items
        self.rows_u = shard_rows(n_users, rank, world)
        self.rows_i_max = shard_rows(n_items, 0, world)                 # shard 0 is the largest
        dev = torch.device("cuda")
        d = spec.dim
        # shard s of a table is initialised from (seed, s), independent of the rank that builds it
        self.users = self._init_shard(self.rows_u, d, seed + 1000 * rank, dev)
        self.sub_epoch = 0
        held = held_item_shard(rank, 0, world)
        self.items = torch.zeros((self.rows_i_max, d), dtype=torch.float32, device=dev)
        n_held = shard_rows(n_items, held, world)
        self.items[:n_held] = self._init_shard(n_held, d, seed + 1 + 1000 * held, dev)
        self._moving = [self.items]
        self.adam = None
        if spec.optimizer == "lazy_adam":
            z = torch.zeros_like
            self.adam = [z(self.users), z(self.users), z(self.items), z(self.items)]
            self._moving += [self.adam[2], self.adam[3]]
        self._spare = [torch.empty_like(t) for t in self._moving]
        self.step = FusedStep(spec)

    @staticmethod
    def _init_shard(rows, d, seed, dev):
        import torch
        g = torch.Generator(device="cuda").manual_seed(seed)
        return (torch.rand((max(rows, 1), d), device=dev, generator=g) - 0.5) * 0.1      # Keras-1 'uniform'

    @property
    def held(self) -> int:
        return held_item_shard(self.rank, self.sub_epoch, self.world)

    def run_block(self, user_ids_local, item_ids_local, n_steps, loss_out=None):
        """n_steps steps on links of block (rank, held): LOCAL ids (id // world) of the two shards"""
        return self.step.run(self.users, self.items, user_ids_local, item_ids_local, n_steps, adam_state=self.adam,
                             loss_out=loss_out)

    def run_block_host(self, user_ids_host, item_ids_host, n_steps, loss_out_host=None):
        return self.step.run_host(self.users, self.items, user_ids_host, item_ids_host, n_steps, loss_out_host)

    def rotate(self):
        """end of a sub-epoch: my item shard (and its optimizer state) goes to rank - 1, rank + 1's comes to me"""
        self._moving, self._spare = ring_rotate(self._moving, self._spare, self.rank, self.world)
        self.items = self._moving[0]
        if self.adam is not None:
            self.adam[2], self.adam[3] = self._moving[1], self._moving[2]
        self.sub_epoch += 1

    def train_epoch(self, blocks, rows_per_step=None):
        """One stratified epoch: `blocks[v]` = (user_ids_local, item_ids_local) CUDA int32 arrays of block (rank, v).
        Returns the list of per-step loss tensors.  Every rank calls it (the rotations pair up)."""
        per = (rows_per_step or self.spec.replicas * self.spec.batch_size_p)
        losses = []
        for _ in range(self.world):
            u, c = blocks[self.held]
            n_steps = u.numel() // per
            if n_steps > 0:
                losses.append(self.run_block(u, c, n_steps)["loss"])
            self.rotate()
        return losses

    def close(self):
        self.step = None
