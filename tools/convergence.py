"""Convergence of the throughput modes against the reference's sequential loop, on a C1-shaped problem with structure to learn.

  python tools/convergence.py [--epochs 12] [--out profiles/r02_convergence.md]

What is compared (same data, same epochs, same hyper-parameters, whole@50 on held-out links after every epoch):
  * R = 1 (the reference's strictly sequential loop, models/train_neg_shared.py:40-58) against R in {2, 4, 8, 16, 37}
    replicas per step (R batches against one table snapshot, updates summed);
  * the stratified multi-GPU schedule (nncf_b200/parallel.py: users in N shards, items in 2N strata, a batch's shared
    negatives come from one stratum) at N = 2 and 4 against the single-GPU epoch - run with every rank as an object of
    this process on one device (LocalPeerGroup): the schedule, not the transport, is what changes the statistics.

Data: the CiteULike shape (5,551 users x 16,980 items, ~205k links), but links are drawn from a planted low-rank preference
model times an item power law, so that recall@50 on held-out links measures how well the factors were learnt (pure
power-law links would only measure popularity).  10 % of the links are held out; candidates = all items; train links are
not masked (as in the reference's whole@k)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nncf_b200 import ops                                     # noqa: E402
from nncf_b200.ops import FusedStep, StepSpec                 # noqa: E402


def planted_links(n_users=5551, n_items=16980, n_links=204986, rank=16, temp=0.35, seed=2017):
    g = torch.Generator(device="cuda").manual_seed(seed)
    Us = torch.randn((n_users, rank), device="cuda", generator=g)
    Vs = torch.randn((n_items, rank), device="cuda", generator=g)
    pop = -1.0 * torch.log(torch.arange(n_items, device="cuda", dtype=torch.float32) + 10.0)
    pop = pop[torch.randperm(n_items, device="cuda", generator=g)]
    act = torch.pow(torch.arange(n_users, device="cuda", dtype=torch.float64) + 10.0, -0.8)
    users = torch.multinomial(act / act.sum(), n_links, replacement=True, generator=g)
    items = torch.empty(n_links, dtype=torch.int64, device="cuda")
    for s in range(0, n_links, 8192):
        u = users[s:s + 8192]
        logits = (Us[u] @ Vs.T) / (temp * rank ** 0.5) + pop[None, :]
        items[s:s + 8192] = torch.multinomial(torch.softmax(logits, dim=1), 1, generator=g)[:, 0]
    perm = torch.randperm(n_links, device="cuda", generator=g)
    users, items = users[perm], items[perm]
    n_test = n_links // 10
    return (users[n_test:].to(torch.int32), items[n_test:].to(torch.int32)), (users[:n_test], items[:n_test]), n_users, n_items


def planted_links_big(n_users, n_items, n_links, rank=16, n_cand=64, temp=0.5, seed=2018):
    """the same kind of planted preference data at a scale where a softmax over all items per link is too much: every link's
    user ~ (rank + 10)^-0.8, then 64 candidate items ~ (rank + 10)^-1.0 and ONE of them drawn ∝ exp(<u*, v*> / (temp sqrt(rank)))"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    Us = torch.randn((n_users, rank), device="cuda", generator=g)
    Vs = torch.randn((n_items, rank), device="cuda", generator=g)
    pu = torch.pow(torch.arange(n_users, device="cuda", dtype=torch.float64) + 10.0, -0.8); cu = torch.cumsum(pu, 0); cu = cu / cu[-1]
    pi = torch.pow(torch.arange(n_items, device="cuda", dtype=torch.float64) + 10.0, -1.0); ci = torch.cumsum(pi, 0); ci = ci / ci[-1]
    perm_u = torch.randperm(n_users, device="cuda", generator=g); perm_i = torch.randperm(n_items, device="cuda", generator=g)
    users = torch.empty(n_links, dtype=torch.int64, device="cuda"); items = torch.empty(n_links, dtype=torch.int64, device="cuda")
    for s in range(0, n_links, 1_000_000):
        n = min(1_000_000, n_links - s)
        u = perm_u[torch.searchsorted(cu, torch.rand(n, device="cuda", generator=g, dtype=torch.float64)).clamp_(max=n_users - 1)]
        cand = perm_i[torch.searchsorted(ci, torch.rand((n, n_cand), device="cuda", generator=g, dtype=torch.float64)).clamp_(max=n_items - 1)]
        logits = torch.einsum("nr,ncr->nc", Us[u], Vs[cand]) / (temp * rank ** 0.5)
        pick = torch.multinomial(torch.softmax(logits, dim=1), 1, generator=g)
        users[s:s + n] = u; items[s:s + n] = cand.gather(1, pick)[:, 0]
    n_test = n_links // 10
    return (users[n_test:].to(torch.int32), items[n_test:].to(torch.int32)), (users[:n_test], items[:n_test]), n_users, n_items


def csr_truth(users, items, n_users, n_items):
    key = torch.unique(users.to(torch.int64) * n_items + items.to(torch.int64))
    owner, cols = key // n_items, (key % n_items).to(torch.int32)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(torch.bincount(owner, minlength=n_users), 0)
    return indptr, cols


def evaluate(EU, EV, truth, k=50):
    ids, _ = ops.eval_topk(EU, EV, k, "bf16")
    _, sums = ops.eval_metrics(ids, truth[0], truth[1])
    s = sums.cpu().numpy()
    return s[1] / max(s[3], 1), s[0] / max(s[3], 1)          # recall@k, MAP@k


def init_tables(n_users, n_items, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return ((torch.rand((n_users, d), device="cuda", generator=g) - 0.5) * 0.1, (torch.rand((n_items, d), device="cuda", generator=g) - 0.5) * 0.1)


def run_single(train, truth, n_users, n_items, R, epochs, opt, lr, d, B, seed):
    EU, EV = init_tables(n_users, n_items, d, 7)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer=opt, learn_rate=lr,
                    replicas=R, neg_loss_weight=128.0, loss_gamma=10.0, u_reg=1e-6)
    step = FusedStep(spec)
    state = [torch.zeros_like(EU), torch.zeros_like(EU), torch.zeros_like(EV), torch.zeros_like(EV)] if opt == "lazy_adam" else None
    g = torch.Generator(device="cuda").manual_seed(seed)
    hist = []
    n = train[0].numel()
    for ep in range(epochs):
        perm = torch.randperm(n, device="cuda", generator=g)       # np.random.shuffle(train) per epoch (train_neg_shared.py:42)
        u, c = train[0][perm].contiguous(), train[1][perm].contiguous()
        n_steps = n // (R * B)                                      # tail dropped (:43-45)
        out = step.run(EU, EV, u, c, n_steps, adam_state=state)
        hist.append((float(out["loss"].mean()),) + evaluate(EU, EV, truth))
    return hist


def run_stratified(train, truth, n_users, n_items, world, R, epochs, opt, lr, d, B, seed):
    from nncf_b200.parallel import LocalPeerGroup, StratifiedTrainer, n_item_strata, shard_rows, stratum_of
    m = n_item_strata(world)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer=opt, learn_rate=lr,
                    replicas=R, neg_loss_weight=128.0, loss_gamma=10.0, u_reg=1e-6)
    group = LocalPeerGroup(world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    ranks = []
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            ranks.append(StratifiedTrainer(spec, n_users, n_items, r, world, seed=7, group=group))
    torch.cuda.synchronize()
    u_all, c_all = train[0].to(torch.int64), train[1].to(torch.int64)
    g = torch.Generator(device="cuda").manual_seed(seed)
    hist = []
    for ep in range(epochs):
        perm = torch.randperm(u_all.numel(), device="cuda", generator=g)
        u, c = u_all[perm], c_all[perm]
        blocks = []
        for r in range(world):
            row = []
            for s_ in range(m):
                sel = ((u % world) == r) & ((c % m) == s_)
                row.append(((u[sel] // world).to(torch.int32).contiguous(), (c[sel] // m).to(torch.int32).contiguous()))
            blocks.append(row)
        torch.cuda.synchronize()
        losses = []
        for _ in range(m):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    tr = ranks[r]
                    bu, bc = blocks[r][tr.held]
                    ns = bu.numel() // (R * B)
                    if ns > 0:
                        losses.append(tr.run_block(bu, bc, ns)["loss"])
                    tr.advance()
        for tr in ranks:
            tr.drain()
        torch.cuda.synchronize()
        # assemble the global tables for the evaluation
        EU = torch.empty((n_users, d), device="cuda"); EV = torch.empty((n_items, d), device="cuda")
        for r, tr in enumerate(ranks):
            EU[r::world] = tr.users[:shard_rows(n_users, r, world)]
            for k in (0, 1):
                s_ = stratum_of(r, tr.phase + k, world)
                EV[s_::m] = tr.slots[(tr.phase + k) % 3][0][:shard_rows(n_items, s_, m)]
        hist.append((float(torch.cat(losses).mean()) if losses else float("nan"),) + evaluate(EU, EV, truth))
    for tr in ranks:
        tr.close()
    return hist


def medium_scale(emit, args):
    """the same comparison where the throughput modes are meant to run: 200k x 200k rows, 20M links (1/5 of C3 in rows and links):
    a stratified block at N = 8 still holds 25k users x 12.5k items, a step of 37 batches touches ~10 % of the rows"""
    d, B, E = 64, 512, 8
    nu = ni = 200_000
    train, test, nu, ni = planted_links_big(nu, ni, 20_000_000)
    truth = csr_truth(test[0], test[1], nu, ni)
    emit("## medium scale: %d users x %d items, %d train / %d held-out links, d = %d, B = %d, %d epochs, recall@50 per epoch" % (
        nu, ni, train[0].numel(), test[0].numel(), d, B, E))
    emit("")
    emit("| mode | " + " | ".join("ep%d" % (e + 1) for e in range(E)) + " | MAP@50 last | loss last | recall last vs R=1 |")
    emit("|---|" + "---|" * (E + 3))
    for opt, lr1, variants in (("lazy_adam", 0.01, ((1, 0.01), (37, 0.01), (37, 0.0608), (37, 0.2))), ("sgd", 1.0, ((1, 1.0), (8, 1.0), (37, 1.0)))):
        base = None
        for R, lr in variants:
            h = run_single(train, truth, nu, ni, R, E, opt, lr, d, B, 1)
            if R == 1:
                base = h[-1][1]
            emit("| %s lr %.3g, 1 GPU, R = %d | " % (opt, lr, R) + " | ".join("%.4f" % x[1] for x in h) + " | %.4f | %.3f | %+.1f %% |" % (h[-1][2], h[-1][0], 100 * (h[-1][1] / base - 1)))
        for R, lr in ((1, lr1), (37, lr1 if opt == "sgd" else 0.0608)):
            h = run_stratified(train, truth, nu, ni, 8, R, E, opt, lr, d, B, 1)
            emit("| %s lr %.3g, stratified N = 8 (16 strata), R = %d | " % (opt, lr, R) + " | ".join("%.4f" % x[1] for x in h) + " | %.4f | %.3f | %+.1f %% |" % (h[-1][2], h[-1][0], 100 * (h[-1][1] / base - 1)))
    emit("")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=40)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sgd-lr", type=float, default=1.0)
    ap.add_argument("--sweep-sgd", action="store_true", help="only: SGD learning-rate sweep at R = 1")
    ap.add_argument("--medium", action="store_true", help="also: 200k x 200k rows, 20M links, 8 epochs (early training only; not converged)")
    args = ap.parse_args()
    d, B, E = 64, 512, args.epochs
    train, test, nu, ni = planted_links()
    truth = csr_truth(test[0], test[1], nu, ni)
    lines = []

    def emit(s=""):
        print(s, flush=True)
        lines.append(s)

    marks = sorted(set([max(E // 4, 1), max(E // 2, 1), max(3 * E // 4, 1), E]))

    def row(name, hs, base):
        rec = np.mean([[h[e][1] for e in range(E)] for h in hs], axis=0)
        mp = np.mean([h[-1][2] for h in hs]); ls = np.mean([h[-1][0] for h in hs])
        spread = np.ptp([h[-1][1] for h in hs]) / 2 if len(hs) > 1 else 0.0
        best = float(np.max(rec))
        emit("| %s | %s | %.4f | %.4f | %.4f | %.3f | %s |" % (name, " | ".join("%.4f" % rec[m - 1] for m in marks), spread, best, mp, ls,
                                                           "-" if base is None else "%+.1f %%" % (100 * (rec[-1] / base - 1))))
        return float(rec[-1])

    if args.sweep_sgd:
        for lr in (0.2, 0.5, 1.0, 2.0, 5.0, 10.0):
            h = run_single(train, truth, nu, ni, 1, E, "sgd", lr, d, B, 1)
            print("sgd lr %g (R = 1): recall@50 by epoch" % lr, " ".join("%.4f" % x[1] for x in h[::max(E // 10, 1)]), "last %.4f loss %.3f" % (h[-1][1], h[-1][0]), flush=True)
        return

    emit("# Convergence of the throughput modes vs the reference's sequential loop (tools/convergence.py)")
    emit("")
    emit("C1-shaped planted problem: %d users x %d items, %d train / %d held-out links, mf, neg_shared skip-gram, B = %d, d = %d, "
         "lambda = 128, u_reg = 1e-6, bf16 kernels; whole@50 on the held-out links (all items are candidates, train links not masked). "
         "Same data, same number of epochs for every row; R = 1 is the reference's sequential loop.  `+-` = half the range over two "
         "shuffle seeds at the last epoch; `best` = best epoch's recall." % (nu, ni, train[0].numel(), test[0].numel(), B, d))
    emit("")
    t0 = time.time()
    seeds = (1, 2) if not args.quick else (1,)
    Rs = (1, 2, 4, 8, 16, 37, 74) if not args.quick else (1, 8, 37)
    hdr = "| mode | " + " | ".join("recall@50 ep%d" % m for m in marks) + " | +- | best | MAP@50 last | loss last | recall last vs R=1 |"
    sep = "|---|" + "---|" * (len(marks) + 5)
    # ---- the reference's optimizer family (lazy Adam, lr 0.01 as in scripts/demos/run_neg_shared.sh)
    emit("## lazy Adam, lr = 0.01 (the reference's demo setting), %d epochs" % E)
    emit("")
    emit(hdr); emit(sep)
    base = None
    for R in Rs:
        for tag, lr in (("", 0.01),) + ((("lr x sqrt(R)", 0.01 * R ** 0.5), ("lr x R", min(0.01 * R, 0.2))) if R > 1 else ()):
            hs = [run_single(train, truth, nu, ni, R, E, "lazy_adam", lr, d, B, seed) for seed in seeds]
            v = row("1 GPU, R = %d%s" % (R, (", " + tag + " = %.3g" % lr) if tag else ""), hs, base)
            if R == 1:
                base = v
    for world in ((2, 4) if not args.quick else (2,)):
        for R in (1, 8):
            for tag, lr in (("", 0.01),) + ((("lr x sqrt(R)", 0.01 * R ** 0.5),) if R > 1 else ()):
                h = run_stratified(train, truth, nu, ni, world, R, E, "lazy_adam", lr, d, B, 1)
                row("stratified N = %d (%d strata), R = %d%s" % (world, 2 * world, R, (", " + tag + " = %.3g" % lr) if tag else ""), [h], base)
    emit("")
    # ---- the stratified schedule to convergence (it mixes more slowly: an item row gets its epoch's updates in one burst)
    E_long = 3 * E
    emit("## lazy Adam, lr = 0.01, stratified schedule to convergence (%d epochs, recall@50 every %d)" % (E_long, E_long // 6))
    emit("")
    emit("| mode | " + " | ".join("ep%d" % e for e in range(E_long // 6, E_long + 1, E_long // 6)) + " | best |")
    emit("|---|" + "---|" * 7)
    h = run_single(train, truth, nu, ni, 1, E_long, "lazy_adam", 0.01, d, B, 1)
    emit("| 1 GPU, R = 1 | " + " | ".join("%.4f" % h[e - 1][1] for e in range(E_long // 6, E_long + 1, E_long // 6)) + " | %.4f |" % max(x[1] for x in h))
    for world in ((2, 4) if not args.quick else (2,)):
        h = run_stratified(train, truth, nu, ni, world, 1, E_long, "lazy_adam", 0.01, d, B, 1)
        emit("| stratified N = %d (%d strata), R = 1 | " % (world, 2 * world) + " | ".join("%.4f" % h[e - 1][1] for e in range(E_long // 6, E_long + 1, E_long // 6)) + " | %.4f |" % max(x[1] for x in h))
    emit("")
    # ---- sparse SGD (this framework's throughput mode; the replicas' updates are summed, not averaged).  The reference has
    #      no SGD setting (it trains with Adam); lr 10 is the fastest at R = 1 but unstable for R >= 4, so two stable ones
    for sgd_lr in (5.0, 2.0):
        emit("## sparse SGD, lr = %g, %d epochs" % (sgd_lr, E))
        emit("")
        emit(hdr); emit(sep)
        base = None
        for R in Rs:
            hs = [run_single(train, truth, nu, ni, R, E, "sgd", sgd_lr, d, B, seed) for seed in seeds]
            v = row("1 GPU, R = %d" % R, hs, base)
            if R == 1:
                base = v
        for world in ((2, 4) if not args.quick else (2,)):
            for R in ((1, 8, 37) if world == 2 else (1, 8)):
                h = run_stratified(train, truth, nu, ni, world, R, E, "sgd", sgd_lr, d, B, 1)
                row("stratified N = %d (%d strata), R = %d" % (world, 2 * world, R), [h], base)
        emit("")
    if args.medium:
        medium_scale(emit, args)
    emit("(wall time %.0f s)" % (time.time() - t0))
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
