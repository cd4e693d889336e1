"""make_golden_sampling.py — second fixture file, tests/golden/sampling_golden.npz, again produced by EXECUTING THE
REFERENCE'S OWN CODE where it lies under /root/reference (nothing is copied):

  utils/objectives.py      get_sampled_neg_shared_loss                   loss VALUES for the 4 losses
  configs/data_utils.py    class GroupSampler (sample, sample_with_negs) run for many batches on a small train array;
                           stored: histograms (groups, (group, first member) pairs, positive / negative rows
                           per group, negative members) and batch-shape statistics, the statistical target for the device GroupSampler.

The class is Python 2 (print statements inside a try block that always fails on the undefined global `conf`,
:274-280): those print lines are blanked before compilation.  Its two `get_sampler` instances are realised with the
analytic distribution of the reference sampler (p ∝ degree^power over ids of positive degree, sampler/nodesampler.cpp
:29-49; the compiled sampler itself is time-seeded and not reproducible), `count_jit` (numba) with np.add.at.

Run:  python tests/golden/make_golden_sampling.py     (needs /root/reference; the GPU box only uses the .npz)
"""
import json
import os
import random
import re
import sys
from collections import defaultdict

import numpy as np

REF = os.environ.get("NNCF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_golden import _K, Conf, slice_function  # noqa: E402


def slice_class(path, name):
    lines = open(os.path.join(REF, path)).read().split("\n")
    start = next(i for i, l in enumerate(lines) if re.match(r"class %s\(" % re.escape(name), l))
    end = start + 1
    while end < len(lines) and (lines[end].strip() == "" or lines[end][0] in " \t#"):
        end += 1
    out = []
    skip_cont = False
    for l in lines[start:end]:
        if skip_cont:                                   # continuation of a blanked py2 print statement
            skip_cont = l.rstrip().endswith("\\")
            continue
        if re.match(r"\s+print ", l):
            out.append(re.match(r"\s+", l).group(0) + "pass")
            skip_cont = l.rstrip().endswith("\\")
            continue
        out.append(l)
    return "\n".join(out) + "\n"


def make_train(rng, n_users=60, n_items=40, n_links=1500):
    pu = (np.arange(n_users) + 3.0) ** -0.8
    pi = (np.arange(n_items) + 3.0) ** -1.0
    u = rng.choice(n_users, size=n_links, p=pu / pu.sum())
    i = rng.choice(n_items, size=n_links, p=pi / pi.sum())
    return np.stack([u, i, np.ones(n_links, dtype=np.int64)], 1)


def main():
    out, meta = {}, {"sns_cases": [], "gs_cases": []}
    rng = np.random.RandomState(4242)
    # ---- get_sampled_neg_shared_loss
    ns = {"np": np, "K": _K,
          "_get_neg_loss_weight": lambda conf: np.array(conf.neg_loss_weight, dtype=np.float64),
          "_get_gamma": lambda conf, verbose=True: np.array(conf.loss_gamma, dtype=np.float64)}
    exec(compile(slice_function("utils/objectives.py", "get_sampled_neg_shared_loss"),
                 "ref:utils/objectives.py:get_sampled_neg_shared_loss", "exec"), ns)
    j = 0
    for loss in ("skip-gram", "mse", "log-loss", "max-margin"):
        for B, k in ((5, 3), (16, 10)):
            lam, gamma = (8.0 if loss == "mse" else 128.0), (0.1 if loss == "max-margin" else 10.0)
            pred = rng.normal(size=(B, 1 + k)) * 0.5
            Lvec = ns["get_sampled_neg_shared_loss"](loss, B, k, Conf(lam, gamma))(None, pred)
            assert Lvec.shape == (B + k, 1)
            out["sns_pred_%d" % j], out["sns_L_%d" % j] = pred, float(np.mean(Lvec))
            meta["sns_cases"].append({"loss": loss, "B": B, "k": k, "lam": lam, "gamma": gamma})
            j += 1
    # ---- GroupSampler
    src = slice_class("configs/data_utils.py", "GroupSampler")

    def get_sampler(ratings, neg_dist="unigram", neg_sampling_power=0.75, column=1, rand_seed=0, batch_mode=True):
        # configs/data_utils.py:193-215 + the distribution of sampler/nodesampler.cpp:29-49
        neg_dist = neg_dist.split("_")[0]
        dist = np.bincount(ratings[:, column]).astype(float)
        if neg_dist == "uniform":
            dist[dist > 0] = 1
        w = np.where(dist > 0, dist ** neg_sampling_power, 0.0)
        p = w / w.sum()
        return (lambda n: np.random.choice(p.size, size=n, p=p).astype(np.int32)) if batch_mode else \
               (lambda: int(np.random.choice(p.size, p=p)))

    def count_jit(ids, dist):
        for i in ids:
            dist[i] += 1

    gns = {"np": np, "random": random, "defaultdict": defaultdict, "get_sampler": get_sampler, "count_jit": count_jit}
    exec(compile(src, "ref:configs/data_utils.py:GroupSampler", "exec"), gns)
    GroupSampler = gns["GroupSampler"]
    train = make_train(rng)
    out["gs_train"] = train
    n_users, n_items = int(train[:, 0].max()) + 1, int(train[:, 1].max()) + 1
    c = 0
    for neg_dist in ("unigram", "uniform"):
        for chop, B, k in ((4, 32, 3), (3, 32, 2)):
            np.random.seed(99 + c)
            random.seed(99 + c)
            gs = GroupSampler(train, group_by="item", chop=chop, neg_dist=neg_dist, neg_sign=-1)
            n_batches = 3000
            h_group = np.zeros(n_items); h_pair = np.zeros((n_items, n_users))
            for _ in range(n_batches):
                b = gs.sample(B)
                assert b.shape == (B, 3)
                np.add.at(h_group, b[::chop, 1], 1)            # first row of every group run
                np.add.at(h_pair, (b[::chop, 1], b[::chop, 0]), 1)   # first member of every run: independent draws
            h_gpos = np.zeros(n_items); h_gneg = np.zeros(n_items); h_mneg = np.zeros(n_users)
            npos, nneg_per_group = [], np.zeros(n_items)
            for _ in range(n_batches):
                b = gs.sample_with_negs(B, k)
                assert b.shape == (B * (1 + k), 3)
                pos = b[b[:, 2] == 1]; neg = b[b[:, 2] == -1]
                assert np.all(b[:len(pos), 2] == 1)            # positives first
                npos.append(len(pos))
                np.add.at(h_gpos, pos[:, 1], 1)
                np.add.at(h_gneg, neg[:, 1], 1)
                np.add.at(h_mneg, neg[:, 0], 1)
            for name, v in (("group", h_group), ("pair", h_pair), ("gpos", h_gpos), ("gneg", h_gneg), ("mneg", h_mneg),
                            ("npos", np.array(npos))):
                out["gs%d_%s" % (c, name)] = v
            meta["gs_cases"].append({"neg_dist": neg_dist, "chop": chop, "B": B, "k": k, "n_batches": n_batches})
            c += 1
    out["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(HERE, "sampling_golden.npz"), **out)
    print("wrote sampling_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
