"""Multi-GPU training / evaluation checked against the oracle.

* `test_pipelined_stratified_schedule_on_one_device`: the 2- and 3-rank stratified schedule with its pipelined stratum
  rotation (peer copies on side streams, arrival / credit flags, three rotating buffers) runs with every rank as an object
  of THIS process on ONE device (LocalPeerGroup) - the same kernels, copies and flag protocol as across GPUs - and one
  epoch must equal the oracle's sequential pass over the same blocks.  Runs wherever `pytest -m gpu` runs.
* `test_sharded_training_and_eval_two_gpus` (needs >= 2 devices): the same under torchrun with CUDA IPC + NCCL, plus the
  peer-memory training step and user-sharded whole@k (tools/multi_gpu_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import nncf_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,opt,precision", [(2, "sgd", "fp32"), (2, "lazy_adam", "fp32"), (3, "sgd", "fp32"), (2, "sgd", "bf16")])
def test_pipelined_stratified_schedule_on_one_device(world, opt, precision):
    from nncf_b200.ops import StepSpec
    from nncf_b200.parallel import (LocalPeerGroup, StratifiedTrainer, n_item_strata, partition_links_by_stratum, shard_rows,
                                    stratum_of)
    nu, ni, B, d, lr, epochs = 1003, 777, 64, 64, 0.05, 2
    m = n_item_strata(world)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision=precision, batch_size_p=B, dim=d, optimizer=opt,
                    learn_rate=lr, replicas=1)
    group = LocalPeerGroup(world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    ranks = []
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            ranks.append(StratifiedTrainer(spec, nu, ni, r, world, seed=11, group=group))
    torch.cuda.synchronize()
    # the global tables the ranks start from (every shard / stratum is initialised from its own seed)
    EU = np.zeros((nu, d)); EV = np.zeros((ni, d))
    for s_ in range(world):
        n = shard_rows(nu, s_, world)
        EU[s_::world] = StratifiedTrainer._init_shard(n, d, 11 + 1000 * s_, "cuda")[:n].cpu().numpy()
    for s_ in range(m):
        n = shard_rows(ni, s_, m)
        EV[s_::m] = StratifiedTrainer._init_shard(n, d, 12 + 1000 * s_, "cuda")[:n].cpu().numpy()
    g = np.random.RandomState(77)
    train = np.stack([g.randint(0, nu, 9000), g.randint(0, ni, 9000), np.ones(9000, dtype=np.int64)], 1)
    blocks = [partition_links_by_stratum(train, r, world) for r in range(world)]
    dev_blocks = [[(torch.from_numpy(np.ascontiguousarray(b[:, 0])).cuda(), torch.from_numpy(np.ascontiguousarray(b[:, 1])).cuda())
                   for b in blocks[r]] for r in range(world)]
    # every rank's phases are enqueued on its own stream, phase by phase; nothing waits on the host: the hand-offs are
    # device-side flags, exactly as between processes
    for _ in range(epochs * m):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                tr = ranks[r]
                u, c = dev_blocks[r][tr.held]
                if u.numel() // B > 0:
                    tr.run_block(u, c, u.numel() // B)
                tr.advance()
    for tr in ranks:
        tr.drain()
    torch.cuda.synchronize()
    # oracle: sequential pass over the same blocks, phase-major (the blocks of a phase are disjoint in users and items)
    mU = np.zeros_like(EU); vU = np.zeros_like(EU); mV = np.zeros_like(EV); vV = np.zeros_like(EV)
    t_adam = {}
    for p in range(epochs * m):
        for r in range(world):
            s_ = stratum_of(r, p, world)
            b = blocks[r][s_]
            for k0 in range(0, (len(b) // B) * B, B):
                u = b[k0:k0 + B, 0].astype(np.int64) * world + r
                c = b[k0:k0 + B, 1].astype(np.int64) * m + s_
                ref = O.step_matmul(EU, EV, u, c, "neg_shared", "skip-gram", 128.0, 10.0)
                if opt == "sgd":
                    EU -= lr * ref["dEU"]; EV -= lr * ref["dEV"]
                else:                                           # every rank's trainer counts its own steps
                    t_adam[r] = t_adam.get(r, 0) + 1
                    EU, mU, vU = O.lazy_adam_sparse(EU, mU, vU, u, ref["dEU"], lr, t_adam[r])
                    EV, mV, vV = O.lazy_adam_sparse(EV, mV, vV, c, ref["dEV"], lr, t_adam[r])
    tol = 1e-3 if precision == "fp32" else 2e-2
    for r, tr in enumerate(ranks):
        assert tr.phase == epochs * m and tr.held == stratum_of(r, 0, world)        # strata are home again
        got_u = tr.users[:shard_rows(nu, r, world)].cpu().numpy().astype(np.float64)
        assert np.max(np.abs(got_u - EU[r::world])) / np.max(np.abs(EU)) <= tol
        # the stratum being "trained" now (2r) and the one that has just arrived for the next phase (2r + 1)
        for k in (0, 1):
            s_ = stratum_of(r, tr.phase + k, world)
            n = shard_rows(ni, s_, m)
            got_v = tr.slots[(tr.phase + k) % 3][0][:n].cpu().numpy().astype(np.float64)
            assert np.max(np.abs(got_v - EV[s_::m])) / np.max(np.abs(EV)) <= tol, (r, k, s_)
    for tr in ranks:
        tr.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_training_and_eval_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MULTI_GPU_CHECK PASSED") == 2
