"""CPU: the oracle restatements added for the GroupSampler, sampled_neg_shared, presample and label-driven 'original'
loss, pinned against tests/golden/sampling_golden.npz (outputs of the reference's own code, see
tests/golden/make_golden_sampling.py) and against torch-CPU autograd in fp64."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nncf_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sampling_golden.npz"), allow_pickle=False)
META = json.loads(str(G["meta"]))


@pytest.mark.parametrize("j", range(len(META["sns_cases"])))
def test_sampled_neg_shared_loss_matches_reference_code(j):
    c = META["sns_cases"][j]
    L, _ = O.sampled_neg_shared_loss_grad(G["sns_pred_%d" % j], c["loss"], c["lam"], c["gamma"])
    # the reference builds its weight matrix in float32 (objectives.py:139,149): lambda / k carries fp32 rounding
    assert abs(L - float(G["sns_L_%d" % j])) <= 1e-6 * max(1.0, abs(L))


@pytest.mark.parametrize("loss", O.LOSSES)
def test_sampled_neg_shared_gradient_matches_autograd(loss):
    rng = np.random.RandomState(3)
    B, k, lam, gamma = 9, 4, 16.0, (0.3 if loss == "max-margin" else 5.0)
    pred = rng.normal(size=(B, 1 + k)) * 0.7
    _, Gd = O.sampled_neg_shared_loss_grad(pred, loss, lam, gamma)
    p = torch.tensor(pred, dtype=torch.float64, requires_grad=True)
    if loss in ("log-loss", "max-margin"):
        D = p[:, :1] - p[:, 1:]
        L = torch.relu(gamma - D).mean() if loss == "max-margin" else (-torch.nn.functional.logsigmoid(gamma * D)).mean()
    else:
        w = torch.ones_like(p); w[:, 1:] = lam / k
        y = torch.ones_like(p); y[:, 1:] = -1.0 if loss == "skip-gram" else 0.0
        L = (-(w * torch.nn.functional.logsigmoid(y * p)).sum() / B) if loss == "skip-gram" else (w * (p - y) ** 2).sum() / B
    L.backward()
    assert np.max(np.abs(p.grad.numpy() - Gd)) <= 1e-12


@pytest.mark.parametrize("loss", ["skip-gram", "log-loss"])
@pytest.mark.parametrize("norm", [False, True])
def test_step_sampled_neg_shared_matches_autograd(loss, norm):
    rng = np.random.RandomState(5)
    nu, ni, d, B, k = 20, 15, 6, 8, 3
    EU, EV = rng.normal(size=(nu, d)), rng.normal(size=(ni, d))
    uid = np.r_[rng.randint(0, nu, B), np.zeros(k, dtype=np.int64)]
    cid = rng.randint(0, ni, B + k)
    lam, gamma, u_reg = 8.0, 2.0, 1e-2
    ref = O.step_sampled_neg_shared(EU, EV, uid, cid, B, k, loss, lam, gamma, u_reg, norm, norm)
    tU = torch.tensor(EU, requires_grad=True); tV = torch.tensor(EV, requires_grad=True)
    Ur, Vr = tU[uid], tV[cid]
    U = torch.nn.functional.normalize(Ur, dim=1, eps=1e-12) if norm else Ur
    V = torch.nn.functional.normalize(Vr, dim=1, eps=1e-12) if norm else Vr
    pred = torch.cat([(U[:B] * V[:B]).sum(1, keepdim=True), U[:B] @ V[B:].T], 1)
    if loss == "log-loss":
        L = (-torch.nn.functional.logsigmoid(gamma * (pred[:, :1] - pred[:, 1:]))).mean()
    else:
        w = torch.ones_like(pred); w[:, 1:] = lam / k
        y = torch.ones_like(pred); y[:, 1:] = -1.0
        L = -(w * torch.nn.functional.logsigmoid(y * pred)).sum() / B
    L = L + u_reg * (Ur ** 2).mean(0).sum()
    L.backward()
    assert abs(float(L) - ref["loss"]) <= 1e-12
    assert np.max(np.abs(tU.grad.numpy() - ref["dEU"])) <= 1e-12
    assert np.max(np.abs(tV.grad.numpy() - ref["dEV"])) <= 1e-12


@pytest.mark.parametrize("loss", ["skip-gram", "mse"])
def test_original_loss_with_labels(loss):
    """explicit y_true reproduces the positional result on the reference layout and follows the labels when shuffled"""
    rng = np.random.RandomState(1)
    B, k, lam = 6, 3, 12.0
    s = rng.normal(size=(1 + k) * B)
    y = np.ones((1 + k) * B); y[B:] = -1.0 if loss == "skip-gram" else 0.0
    L0, g0 = O.original_loss_grad(s, B, k, loss, lam, 1.0)
    L1, g1 = O.original_loss_grad(s, B, k, loss, lam, 1.0, y_true=y)
    assert abs(L0 - L1) <= 1e-12 and np.max(np.abs(g0 - g1)) <= 1e-12
    perm = rng.permutation(s.size)
    L2, g2 = O.original_loss_grad(s[perm], B, k, loss, lam, 1.0, y_true=y[perm])
    assert abs(L2 - L0) <= 1e-12 and np.max(np.abs(g2 - g0[perm])) <= 1e-12


def test_presample_rows_layouts():
    rng = np.random.RandomState(0)
    n, k = 7, 3
    tp = np.stack([rng.randint(0, 9, n), rng.randint(0, 9, n), np.ones(n, dtype=np.int64)], 1)
    negs = rng.randint(100, 200, n * k)
    a = O.presample_rows(tp, k, negs, 1, -1, 0)
    assert a.shape == (n * (1 + k), 3) and np.array_equal(a[::1 + k], tp)
    assert np.array_equal(a.reshape(n, 1 + k, 3)[:, 1:, 1].reshape(-1), negs)
    assert np.array_equal(a.reshape(n, 1 + k, 3)[:, 1:, 0], np.repeat(tp[:, :1], k, 1)) and np.all(a.reshape(n, 1 + k, 3)[:, 1:, 2] == -1)
    b = O.presample_rows(tp, k, negs, 0, 0, 1)
    assert np.array_equal(b[:n], tp) and np.array_equal(b[n:, 0], negs) and np.array_equal(b[n:, 1], np.repeat(tp[:, 1], k))
    assert np.all(b[n:, 2] == 0)
    sb = O.assemble_sns_batch(tp[:4], 2, np.array([5, 6]))
    assert sb.shape == (6, 3) and np.array_equal(sb[4:], [[0, 5, 0], [0, 6, 0]])


def chi2_two_sample(a, b):
    """two-sample chi-square statistic / dof for count vectors a, b (cells with enough mass only)"""
    a = np.asarray(a, dtype=np.float64).reshape(-1); b = np.asarray(b, dtype=np.float64).reshape(-1)
    keep = (a + b) >= 20
    a, b = a[keep], b[keep]
    k1, k2 = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    return float(np.sum((k1 * a - k2 * b) ** 2 / (a + b))) / max(int(keep.sum()) - 1, 1)


def group_sampler_histograms(sampler_sample, sampler_negs, c, n_items, n_users, n_batches):
    """the same statistics make_golden_sampling.py stores, from any implementation (returns dict of arrays)"""
    chop, B, k = c["chop"], c["B"], c["k"]
    h = {"group": np.zeros(n_items), "pair": np.zeros((n_items, n_users)), "gpos": np.zeros(n_items),
         "gneg": np.zeros(n_items), "mneg": np.zeros(n_users), "npos": []}
    for b in sampler_sample(B, n_batches):
        np.add.at(h["group"], b[::chop, 1], 1)
        np.add.at(h["pair"], (b[::chop, 1], b[::chop, 0]), 1)
    for b in sampler_negs(B, k, n_batches):
        pos, neg = b[b[:, 2] == 1], b[b[:, 2] == -1]
        assert len(pos) + len(neg) == B * (1 + k) and np.all(b[:len(pos), 2] == 1)
        h["npos"].append(len(pos))
        np.add.at(h["gpos"], pos[:, 1], 1); np.add.at(h["gneg"], neg[:, 1], 1); np.add.at(h["mneg"], neg[:, 0], 1)
    h["npos"] = np.array(h["npos"])
    return h


def check_group_sampler_histograms(h, ci, nb):
    """h (from group_sampler_histograms) against the golden histograms of the reference's own GroupSampler"""
    c = META["gs_cases"][ci]
    ref = {name: G["gs%d_%s" % (ci, name)] for name in ("group", "pair", "gpos", "gneg", "mneg", "npos")}
    # independent draws: groups, (group, first member) pairs, negative members -> two-sample chi-square
    for name in ("group", "pair", "mneg"):
        assert chi2_two_sample(h[name], ref[name]) < 1.5, name
    # positive rows come in complete runs of `chop` rows per group draw
    assert chi2_two_sample(h["gpos"] / c["chop"], ref["gpos"] / c["chop"]) < 1.5
    # negative rows per group: runs of ~k * chop * p_n/p_d[group] rows per draw (:357-358), cut by the batch truncation, so
    # the cells are not multinomial: compare the normalised histograms (total variation); the positives-per-batch mean below
    # pins the overall positives / negatives split
    tv = 0.5 * np.sum(np.abs(h["gneg"] / h["gneg"].sum() - ref["gneg"] / ref["gneg"].sum()))
    assert tv < 0.05, tv
    assert abs(h["npos"].mean() - ref["npos"].mean()) <= 4 * ref["npos"].std() / np.sqrt(nb) + 1e-9


@pytest.mark.parametrize("ci", range(len(META["gs_cases"])))
def test_group_sampler_oracle_matches_reference_distribution(ci):
    c = META["gs_cases"][ci]
    train = G["gs_train"]
    n_users, n_items = int(train[:, 0].max()) + 1, int(train[:, 1].max()) + 1
    o = O.GroupSamplerOracle(train, "item", c["chop"], c["neg_dist"], -1, rng=np.random.RandomState(7 + ci))
    nb = 1500
    h = group_sampler_histograms(lambda B, n: (o.sample(B) for _ in range(n)),
                                 lambda B, k, n: (o.sample_with_negs(B, k) for _ in range(n)), c, n_items, n_users, nb)
    check_group_sampler_histograms(h, ci, nb)
    # every positive row is a real train link
    links = set(map(tuple, train[:, :2]))
    b = o.sample_with_negs(c["B"], c["k"])
    assert all((u, i) in links for u, i, y in b if y == 1)
