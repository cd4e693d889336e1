"""Run under torchrun on >= 2 GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
       --master-port 29511 tools/multi_gpu_check.py
Checks the row-sharded training step (peer-memory gather + remote red.add + device barriers) against the CPU oracle:
after one synchronous step in which every rank contributes its own batch, the GLOBAL tables must equal
table_before - lr * sum_over_ranks(dE_rank), every gradient taken on the same snapshot.
Then the stratified (DSGD) schedule with its pipelined stratum rotation (peer copies on side streams + flag hand-offs over
CUDA IPC): two epochs must equal the oracle's SEQUENTIAL pass over the same blocks (the blocks of a phase are disjoint in
users and items, so their order does not matter)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nncf_oracle as O   # noqa: E402  (test tool: the oracle is the checker)
from nncf_b200.ops import StepSpec    # noqa: E402
from nncf_b200.parallel import (ShardedTrainer, StratifiedTrainer, n_item_strata, partition_links_by_stratum, shard_rows,  # noqa: E402
                                sharded_whole_eval, stratum_of)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    nu, ni, B, d, lr, R = 1003, 777, 256, 64, 0.1, 2
    ok = True
    for precision, tol in (("fp32", 1e-3), ("bf16", 1e-2)):
        spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision=precision, batch_size_p=B, dim=d, optimizer="sgd",
                        learn_rate=lr, replicas=R)
        tr = ShardedTrainer(spec, nu, ni, rank, world, seed=5)
        dist.barrier()
        all_u = np.arange(nu); all_i = np.arange(ni)
        EU0 = tr.users.gather_global(all_u).cpu().numpy().astype(np.float64)
        EV0 = tr.items.gather_global(all_i).cpu().numpy().astype(np.float64)
        dist.barrier()
        ids = [np.random.RandomState(100 + r) for r in range(world)]
        batches = [(g.randint(0, nu, size=R * B).astype(np.int32), g.randint(0, ni, size=R * B).astype(np.int32)) for g in ids]
        mu, mi = batches[rank]
        out = tr.run(torch.from_numpy(mu).cuda(), torch.from_numpy(mi).cuda(), 1)
        torch.cuda.synchronize()
        dist.barrier()
        EU1 = tr.users.gather_global(all_u).cpu().numpy().astype(np.float64)
        EV1 = tr.items.gather_global(all_i).cpu().numpy().astype(np.float64)
        dU = np.zeros_like(EU0); dV = np.zeros_like(EV0); my_losses = []
        for r in range(world):
            for rep in range(R):
                u, c = batches[r][0][rep * B:(rep + 1) * B], batches[r][1][rep * B:(rep + 1) * B]
                ref = O.step_matmul(EU0, EV0, u, c, "neg_shared", "skip-gram", 128.0, 10.0)
                dU += ref["dEU"]; dV += ref["dEV"]
                if r == rank:
                    my_losses.append(ref["loss"])
        eu = np.max(np.abs((EU1 - EU0) + lr * dU)) / np.max(np.abs(lr * dU))
        ev = np.max(np.abs((EV1 - EV0) + lr * dV)) / np.max(np.abs(lr * dV))
        el = np.max(np.abs(out["loss"].cpu().numpy() - np.array(my_losses)) / np.array(my_losses))
        good = eu <= tol and ev <= tol and el <= tol
        ok &= good
        print("rank %d %s: rel err dEU %.2e dEV %.2e loss %.2e -> %s" % (rank, precision, eu, ev, el, "OK" if good else "FAIL"), flush=True)
        # a few more steps must keep the ranks in lock-step (barriers pair up) and the loss finite
        out = tr.run(torch.from_numpy(np.tile(mu, 8)).cuda(), torch.from_numpy(np.tile(mi, 8)).cuda(), 8)
        torch.cuda.synchronize()
        assert np.all(np.isfinite(out["loss"].cpu().numpy()))
        dist.barrier()
        tr.close()
    # ---- stratified schedule (pipelined rotation) against the sequential oracle
    for opt in ("sgd", "lazy_adam"):
        nu, ni, B, d, lr, epochs = 1003, 777, 128, 64, 0.05, 2
        m = n_item_strata(world)
        spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision="fp32", batch_size_p=B, dim=d, optimizer=opt,
                        learn_rate=lr, replicas=1)
        st = StratifiedTrainer(spec, nu, ni, rank, world, seed=11)
        dist.barrier()
        EU = np.zeros((nu, d)); EV = np.zeros((ni, d))
        for s_ in range(world):                                # every rank can rebuild every shard's initial values
            n = shard_rows(nu, s_, world)
            EU[s_::world] = StratifiedTrainer._init_shard(n, d, 11 + 1000 * s_, "cuda")[:n].cpu().numpy()
        for s_ in range(m):
            n = shard_rows(ni, s_, m)
            EV[s_::m] = StratifiedTrainer._init_shard(n, d, 12 + 1000 * s_, "cuda")[:n].cpu().numpy()
        g = np.random.RandomState(77)
        train = np.stack([g.randint(0, nu, 12000), g.randint(0, ni, 12000), np.ones(12000, dtype=np.int64)], 1)
        blocks = partition_links_by_stratum(train, rank, world)
        dev_blocks = [(torch.from_numpy(np.ascontiguousarray(b[:, 0])).cuda(), torch.from_numpy(np.ascontiguousarray(b[:, 1])).cuda()) for b in blocks]
        for _ in range(epochs):
            st.train_epoch(dev_blocks)
        st.drain()
        dist.barrier()
        torch.cuda.synchronize()
        assert st.held == stratum_of(rank, 0, world)            # the strata are home again
        # oracle: sequential pass over the same blocks, phase-major
        mU = np.zeros_like(EU); vU = np.zeros_like(EU); mV = np.zeros_like(EV); vV = np.zeros_like(EV)
        t_adam = {}
        for p in range(epochs * m):
            for r in range(world):
                v = stratum_of(r, p, world)
                b = partition_links_by_stratum(train, r, world)[v]
                for k0 in range(0, (len(b) // B) * B, B):
                    u = b[k0:k0 + B, 0].astype(np.int64) * world + r
                    c = b[k0:k0 + B, 1].astype(np.int64) * m + v
                    ref = O.step_matmul(EU, EV, u, c, "neg_shared", "skip-gram", 128.0, 10.0)
                    if opt == "sgd":
                        EU -= lr * ref["dEU"]; EV -= lr * ref["dEV"]
                    else:                                       # every rank's trainer counts its own steps
                        t_adam[r] = t_adam.get(r, 0) + 1
                        EU, mU, vU = O.lazy_adam_sparse(EU, mU, vU, u, ref["dEU"], lr, t_adam[r])
                        EV, mV, vV = O.lazy_adam_sparse(EV, mV, vV, c, ref["dEV"], lr, t_adam[r])
        mineU = st.users[:shard_rows(nu, rank, world)].cpu().numpy().astype(np.float64)
        eu = np.max(np.abs(mineU - EU[rank::world])) / np.max(np.abs(EU))
        ev = 0.0
        for k in (0, 1):                                        # the stratum held for the current phase and the one that arrived for the next
            s_ = stratum_of(rank, st.phase + k, world)
            n = shard_rows(ni, s_, m)
            mine = st.slots[(st.phase + k) % 3][0][:n].cpu().numpy().astype(np.float64)
            ev = max(ev, np.max(np.abs(mine - EV[s_::m])) / np.max(np.abs(EV)))
        good = eu <= 1e-3 and ev <= 1e-3
        ok &= good
        print("rank %d stratified (pipelined) %s: rel err users %.2e items %.2e -> %s" % (rank, opt, eu, ev, "OK" if good else "FAIL"), flush=True)
        dist.barrier()
        st.close()
    # user-sharded whole@k: global metrics must equal the single-process value
    rng = np.random.RandomState(3)
    nU, nI, k = 500, 2000, 20
    U = rng.normal(size=(nU, 32)).astype(np.float32); V = rng.normal(size=(nI, 32)).astype(np.float32)
    truth = (rng.uniform(size=(nU, nI)) < 0.01).astype(np.int32)
    indptr = np.concatenate([[0], np.cumsum(truth.sum(1))]).astype(np.int64); cols = np.nonzero(truth)[1].astype(np.int32)
    tU, tV = torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()
    res = sharded_whole_eval(lambda lo, hi: tU[lo:hi].contiguous(), tV, indptr, cols, nU, k, "fp32", rank, world)
    ref = O.evaluate_mat(truth, U.astype(np.float64) @ V.astype(np.float64).T, k)
    good = abs(res["map"] - ref["map"]) < 1e-6 and abs(res["recall"] - ref["recall"]) < 1e-6 and res["n_users"] == ref["n_users"]
    ok &= good
    print("rank %d sharded whole@%d: map %.6f (ref %.6f) -> %s" % (rank, k, res["map"], ref["map"], "OK" if good else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    print("MULTI_GPU_CHECK %s" % ("PASSED" if ok else "FAILED"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
