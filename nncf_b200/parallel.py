"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous and scalar reductions.

* whole@k evaluation shards by USER with no data-path collective (`shard_range`, `allreduce_metric_sums`).
* training shards the embedding tables by ROW (`owner_of`, `local_row`, `ShardedTable`): row `id` lives on rank
  `id % N` at local row `id // N`.  Shards are peer-visible allocations (CUDA IPC); the step kernels read and update
  remote rows directly over NVLink (include/nncf_b200.h, section 3b) — no all-to-all staging buffers.

* `StratifiedTrainer`: the schedule that scales.  The step moves ~2 KB of embedding rows per positive link; at the
  single-GPU rate (9e8 links/s) that is 1.8 TB/s of random row traffic per GPU, more than NVLink 5 carries (0.9 TB/s per
  direction), so ANY scheme that fetches rows from their owners per step is NVLink-bound (measured with the peer-memory
  mode: 5.3e8 links/s on 2 GPUs against 7.2e8 on one).  Stratified SGD (Gemulla et al., KDD'11, "DSGD") removes the
  per-step exchange: rank r owns user shard r (user % N) for good; the items are cut into M = 2N strata (item % M) and
  links into N x M blocks.  In phase p rank r trains ONLY on block (r, (2r + p) % M), so every row it touches is local
  and the single-GPU kernels run unchanged.  The rotation is PIPELINED: while rank r trains stratum (2r + p) it also
  holds the stratum it will train next, (2r + p + 1), which rank r + 1 finished one phase earlier and pushes into r's
  memory DURING phase p (copy engines over NVLink peer mappings, `nncf_peer_copy`, on a side stream; arrival and
  buffer-free credits are flag words in peer memory, `nncf_peer_signal` / `nncf_peer_wait`, so neither the host nor
  the compute stream ever waits for a transfer that has had a whole phase to complete).  Three buffers per rank rotate
  through the roles train / incoming / outgoing.  The strata trained concurrently are disjoint in users AND items, so a
  phase equals a sequential pass over its N blocks in any order (checked against the oracle: tests/test_gpu_multi.py
  runs the 2-rank schedule on ONE device, tools/multi_gpu_check.py on two).  Declared difference to a single-GPU epoch: a
  batch's shared negatives come from the stratum of its block (a random 1/M of the items under id % M).

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


def n_item_strata(world: int) -> int:
    """item strata of the pipelined stratified schedule: two per rank (one trained, one in flight)"""
    return 2 * world if world > 1 else 1


def stratum_of(rank: int, phase: int, world: int) -> int:
    """item stratum trained by `rank` in phase `phase` (phases count on across epochs)"""
    return (2 * rank + phase) % n_item_strata(world)


def slot_of(phase: int) -> int:
    """which of a rank's three stratum buffers is trained in `phase`; (phase + 1) % 3 is being filled for the next phase,
    (phase - 1) % 3 is being sent to rank - 1 (the same on every rank, so sender and receiver agree without talking)"""
    return phase % 3


def partition_links_by_stratum(train, rank: int, world: int):
    """Rows of `train` (int [n, 3]) owned by `rank` (user % world == rank) split by item stratum (item % M, M = 2 world).
    Returns M int32 arrays [n_s, 3] with LOCAL ids (user // world, item // M); original order kept inside a block."""
    train = np.asarray(train)
    m = n_item_strata(world)
    mine = train[train[:, 0] % world == rank]
    out = []
    for s in range(m):
        b = mine[mine[:, 1] % m == s].astype(np.int32, copy=True)
        b[:, 0] //= world
        b[:, 1] //= m
        out.append(b)
    return out


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


class LocalPeerGroup:
    """In-process stand-in for the CUDA-IPC exchange: every 'rank' is an object of ONE process on ONE device and publishes
    its arena pointer here (plain device pointers are valid for every rank).  Used by the single-device schedule test;
    the kernels, copies and flag protocol are exactly the multi-process ones."""

    def __init__(self, world: int):
        self.world = world
        self.ptrs = [0] * world
        self.keep = [None] * world


class _Arena:
    """one peer-visible allocation per rank: 3 stratum slots x (items [, adam m, adam v]) + 2 flag words"""

    def __init__(self, nbytes: int, rank: int, world: int, group):
        import torch
        self.rank, self.world = rank, world
        if group is None:
            self.mem = PeerMemory(nbytes, rank, world)
            self.local_ptr, self.ptrs = self.mem.local_ptr, self.mem.ptrs
        else:
            self.mem = None
            buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
            group.keep[rank] = buf
            group.ptrs[rank] = buf.data_ptr()
            self.local_ptr, self.ptrs = buf.data_ptr(), group.ptrs          # (the list fills in as the other ranks are built)

    def close(self):
        if self.mem is not None:
            self.mem.close()


class StratifiedTrainer:
    """Stratified (DSGD-style) multi-GPU training with a pipelined stratum rotation (see the module docstring): rank r
    keeps user shard r; the items live in M = 2 world strata; in phase p rank r trains stratum (2r + p) % M out of one of
    three local buffers while the stratum of phase p + 1 arrives from rank r + 1 and the one of phase p - 1 leaves for
    rank r - 1 on the copy engines.  Sparse SGD and lazy Adam (the optimizer state of a stratum travels with it).
    `group`: None = one process per GPU (CUDA IPC through torch.distributed); a LocalPeerGroup = all ranks in this process."""

    def __init__(self, spec, n_users: int, n_items: int, rank: int, world: int, seed: int = 7, group=None):
        import torch
        from .ops import FusedStep
        self.spec, self.rank, self.world = spec, rank, world
        self.n_users, self.n_items = n_users, n_items
        self.m = n_item_strata(world)
        self.rows_u = shard_rows(n_users, rank, world)
        self.rows_s_max = shard_rows(n_items, 0, self.m)                 # stratum 0 is the largest
        dev = torch.device("cuda")
        d = spec.dim
        self.users = self._init_shard(self.rows_u, d, seed + 1000 * rank, dev)
        self.n_mov = 3 if spec.optimizer == "lazy_adam" else 1           # tensors that travel with a stratum
        self.t_bytes = max(self.rows_s_max, 1) * d * 4
        n_slots = 3 if world > 1 else 1
        self.off_flags = n_slots * self.n_mov * self.t_bytes
        self.arena = _Arena(self.off_flags + 256, rank, world, group)
        self.slots = [[torch.as_tensor(_CudaBuffer(self.arena.local_ptr + (k * self.n_mov + j) * self.t_bytes,
                                                   (max(self.rows_s_max, 1), d)), device="cuda")
                       for j in range(self.n_mov)] for k in range(n_slots)]
        for k in range(min(2, n_slots)):                                 # phase 0: stratum 2r in slot 0, 2r + 1 ready in slot 1
            st = stratum_of(rank, k, world)
            n_st = shard_rows(n_items, st, self.m)
            self.slots[k][0].zero_()
            self.slots[k][0][:n_st] = self._init_shard(n_st, d, seed + 1 + 1000 * st, dev)[:n_st]
            for j in range(1, self.n_mov):
                self.slots[k][j].zero_()
        self.user_adam = [torch.zeros_like(self.users), torch.zeros_like(self.users)] if self.n_mov == 3 else None
        self.phase = 0
        self.xfer = torch.cuda.Stream() if world > 1 else None
        self.step = FusedStep(spec)

    @staticmethod
    def _init_shard(rows, d, seed, dev):
        import torch
        g = torch.Generator(device="cuda").manual_seed(seed)
        return (torch.rand((max(rows, 1), d), device=dev, generator=g) - 0.5) * 0.1      # Keras-1 'uniform'

    # ---- what the current phase trains on
    @property
    def held(self) -> int:
        return stratum_of(self.rank, self.phase, self.world)

    @property
    def items(self):
        return self.slots[slot_of(self.phase) if self.world > 1 else 0][0]

    @property
    def adam(self):
        if self.user_adam is None:
            return None
        sl = self.slots[slot_of(self.phase) if self.world > 1 else 0]
        return [self.user_adam[0], self.user_adam[1], sl[1], sl[2]]

    def run_block(self, user_ids_local, item_ids_local, n_steps, loss_out=None):
        """n_steps steps on links of block (rank, held): LOCAL ids (user // world, item // M)"""
        return self.step.run(self.users, self.items, user_ids_local, item_ids_local, n_steps, adam_state=self.adam,
                             loss_out=loss_out)

    def run_block_host(self, user_ids_host, item_ids_host, n_steps, loss_out_host=None):
        return self.step.run_host(self.users, self.items, user_ids_host, item_ids_host, n_steps, loss_out_host)

    # ---- the pipelined rotation
    def _flag(self, rank: int, which: int) -> int:
        """address of flag word `which` (0 = arrived, 1 = credit) in `rank`'s arena"""
        return self.arena.ptrs[rank % self.world] + self.off_flags + 4 * which

    def advance(self):
        """End of a phase.  Enqueues, without any host wait: (side stream) push the stratum just trained into rank - 1's
        incoming buffer once that buffer is free, then raise rank - 1's `arrived` and rank + 1's `credit`; (compute stream)
        wait until the stratum of the new phase has arrived - it has had the whole previous phase to do so."""
        import torch
        from ._lib import check, lib
        self.phase += 1
        if self.world == 1:
            return
        p = self.phase
        cur = torch.cuda.current_stream()
        done = torch.cuda.Event()
        done.record(cur)                                     # phase p - 1 has been trained
        self.xfer.wait_event(done)
        xs = C.c_void_p(self.xfer.cuda_stream)
        if p >= 2:                                           # rank - 1 has finished SENDING the buffer I am about to overwrite
            check(lib.nncf_peer_wait(C.c_void_p(self._flag(self.rank, 1)), p - 1, xs))
        src_slot, dst_slot = slot_of(p - 1), slot_of(p + 1)
        left = self.arena.ptrs[(self.rank - 1) % self.world]
        for j in range(self.n_mov):
            check(lib.nncf_peer_copy(C.c_void_p(left + (dst_slot * self.n_mov + j) * self.t_bytes),
                                     C.c_void_p(self.arena.local_ptr + (src_slot * self.n_mov + j) * self.t_bytes),
                                     self.t_bytes, xs))
        check(lib.nncf_peer_signal(C.c_void_p(self._flag(self.rank - 1, 0)), p, xs))     # rank - 1: stratum of phase p + 1 is there
        check(lib.nncf_peer_signal(C.c_void_p(self._flag(self.rank + 1, 1)), p, xs))     # rank + 1: my slot (p - 1) % 3 is free again
        if p >= 2:                                           # phase p trains what rank + 1 pushed during phase p - 1
            check(lib.nncf_peer_wait(C.c_void_p(self._flag(self.rank, 0)), p - 1, C.c_void_p(cur.cuda_stream)))

    rotate = advance                                         # (name used by bench.py and the round-1 callers)

    def drain(self):
        """blocks the host until this rank's transfers have completed (end of training / before reading the buffers)"""
        import torch
        torch.cuda.current_stream().synchronize()
        if self.xfer is not None:
            self.xfer.synchronize()

    def train_epoch(self, blocks, rows_per_step=None):
        """One stratified epoch = M phases: `blocks[s]` = (user_ids_local, item_ids_local) CUDA int32 arrays of block
        (rank, s).  Returns the per-phase loss tensors.  Every rank calls it (the transfers pair up)."""
        per = (rows_per_step or self.spec.replicas * self.spec.batch_size_p)
        losses = []
        for _ in range(self.m):
            u, c = blocks[self.held]
            n_steps = u.numel() // per
            if n_steps > 0:
                losses.append(self.run_block(u, c, n_steps)["loss"])
            self.advance()
        return losses

    def close(self):
        self.drain()
        self.step = None
        self.slots = None
        self.arena.close()
