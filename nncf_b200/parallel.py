"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous and scalar reductions.

* whole@k evaluation shards by USER with no data-path collective (`shard_range`, `allreduce_metric_sums`).
* training shards the embedding tables by ROW (`owner_of`, `local_row`, `ShardedTable`): row `id` lives on rank
  `id % N` at local row `id // N`.  Shards are peer-visible allocations (CUDA IPC); the step kernels read and update
  remote rows directly over NVLink (include/nncf_b200.h, section 3b) — no all-to-all staging buffers.

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
