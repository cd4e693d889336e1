"""CPU tests of the oracle itself (no GPU): analytic gradients vs torch-CPU autograd of the forward formulas in
fp64, the sampler restatement vs the REAL compiled reference sampler (oracle/_ref), metrics on hand cases, and the
committed golden fixtures."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import nncf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSSES = ["skip-gram", "mse", "log-loss", "max-margin"]


# ------------------------------------------------------------------------------------------------
# forward formulas restated in torch, literally following the reference code, for autograd
# ------------------------------------------------------------------------------------------------
def t_neg_shared(S, loss, lam, gamma):
    """utils/objectives.py:89-115 line by line (K.diag -> torch.diagonal, np.diag masks)."""
    B = S.shape[0]
    eye = torch.eye(B, dtype=S.dtype)
    if loss == "max-margin":
        diff = torch.diagonal(S) - S                         # (B,) - (B,B) broadcast, objectives.py:92
        margin_mat = -gamma * eye + gamma                    # :93
        return torch.mean(torch.relu(margin_mat - diff))
    if loss == "log-loss":
        diff = torch.diagonal(S) - S
        return torch.mean(-torch.log(torch.sigmoid(gamma * diff)))
    w = lam / (B - 1)
    W = (1 - w) * eye + w
    if loss == "skip-gram":
        Y = 2 * eye - 1
        return torch.sum(-W * torch.log(torch.sigmoid(Y * S))) / B
    return torch.sum(W * (S - eye) ** 2) / B


def t_group(P, pos, loss, lam, gamma):
    """utils/objectives.py:163-220."""
    B, nu = P.shape
    Yp = torch.zeros_like(P)
    Yp[torch.arange(B), torch.as_tensor(pos, dtype=torch.long)] = 1.0
    ppos = P[torch.arange(B), torch.as_tensor(pos, dtype=torch.long)].reshape(-1, 1)
    if loss == "max-margin":
        return torch.mean(torch.relu(gamma - (ppos - P)))
    if loss == "log-loss":
        return torch.mean(-torch.log(torch.sigmoid(gamma * (ppos - P))))
    w = lam / (nu - 1)
    W = w + (1 - w) * Yp
    if loss == "skip-gram":
        return torch.sum(-W * torch.log(torch.sigmoid((2 * Yp - 1) * P))) / B
    return torch.sum(W * (P - Yp) ** 2) / B


def t_original(s, B, k, loss, lam, gamma):
    """utils/objectives.py:46-73."""
    sp, sn = s[:B], s[B:]
    if loss in ("max-margin", "log-loss"):
        diff = sp.repeat_interleave(k) - sn
        if loss == "max-margin":
            return torch.mean(torch.relu(gamma - diff))
        return torch.mean(-torch.log(torch.sigmoid(gamma * diff)))
    w = lam / k
    y = torch.cat([torch.ones(B, dtype=s.dtype), (-1.0 if loss == "skip-gram" else 0.0) * torch.ones(k * B, dtype=s.dtype)])
    if loss == "skip-gram":
        wt = 1 + (1 - y) / 2.0 * (w - 1)
        return torch.sum(-wt * torch.log(torch.sigmoid(y * s))) / B
    wt = 1 + (1 - y) * (w - 1)
    return torch.sum(wt * (y - s) ** 2) / B


@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("B", [2, 3, 17])
def test_neg_shared_grad_matches_autograd(loss, B):
    rng = np.random.RandomState(B)
    S = rng.normal(size=(B, B))
    L, G = O.neg_shared_loss_grad(S, loss, 128.0, 0.7)
    St = torch.tensor(S, dtype=torch.float64, requires_grad=True)
    Lt = t_neg_shared(St, loss, 128.0, 0.7)
    Lt.backward()
    assert abs(L - Lt.item()) < 1e-12 * max(1, abs(L))
    np.testing.assert_allclose(G, St.grad.numpy(), atol=1e-13)


@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("B,nu", [(2, 2), (5, 3), (16, 9)])
def test_group_grad_matches_autograd(loss, B, nu):
    rng = np.random.RandomState(B * nu)
    P = rng.normal(size=(B, nu))
    pos = rng.randint(0, nu, size=B)
    L, G = O.group_neg_shared_loss_grad(P, pos, loss, 128.0, 0.7)
    Pt = torch.tensor(P, dtype=torch.float64, requires_grad=True)
    Lt = t_group(Pt, pos, loss, 128.0, 0.7)
    Lt.backward()
    assert abs(L - Lt.item()) < 1e-12 * max(1, abs(L))
    np.testing.assert_allclose(G, Pt.grad.numpy(), atol=1e-13)


@pytest.mark.parametrize("loss", LOSSES)
def test_original_grad_matches_autograd(loss):
    rng = np.random.RandomState(1)
    B, k = 6, 4
    s = rng.normal(size=(1 + k) * B)
    L, g = O.original_loss_grad(s, B, k, loss, 128.0, 0.7)
    st = torch.tensor(s, dtype=torch.float64, requires_grad=True)
    Lt = t_original(st, B, k, loss, 128.0, 0.7)
    Lt.backward()
    assert abs(L - Lt.item()) < 1e-12 * max(1, abs(L))
    np.testing.assert_allclose(g, st.grad.numpy(), atol=1e-13)


@pytest.mark.parametrize("scheme", ["neg_shared", "group_neg_shared"])
@pytest.mark.parametrize("loss", LOSSES)
def test_full_step_grads_match_autograd(scheme, loss):
    """gather -> l2-normalise -> U V^T -> loss (+ activity regulariser): table gradients vs autograd."""
    rng = np.random.RandomState(3)
    nu, ni, d, B = 20, 9, 7, 12
    EU, EV = rng.normal(size=(nu, d)), rng.normal(size=(ni, d))
    uid, cid = rng.randint(0, nu, size=B), rng.randint(0, ni, size=B)
    ref = O.step_matmul(EU, EV, uid, cid, scheme, loss, 128.0, 0.7, u_reg=1e-2, norm_u=True, norm_v=True)
    tU = torch.tensor(EU, requires_grad=True); tV = torch.tensor(EV, requires_grad=True)
    Ur = tU[torch.as_tensor(uid)]
    U = torch.nn.functional.normalize(Ur, dim=-1, eps=1e-6)
    if scheme == "neg_shared":
        V = torch.nn.functional.normalize(tV[torch.as_tensor(cid)], dim=-1, eps=1e-6)
        L = t_neg_shared(U @ V.T, loss, 128.0, 0.7)
    else:
        cu, cx = O.unique_first_occurrence(cid)
        V = torch.nn.functional.normalize(tV[torch.as_tensor(cu)], dim=-1, eps=1e-6)
        L = t_group(U @ V.T, cx, loss, 128.0, 0.7)
    L = L + 1e-2 * torch.sum(torch.mean(Ur ** 2, dim=0))
    L.backward()
    assert abs(ref["loss"] - L.item()) < 1e-12 * max(1, abs(L.item()))
    np.testing.assert_allclose(ref["dEU"], tU.grad.numpy(), atol=1e-12)
    np.testing.assert_allclose(ref["dEV"], tV.grad.numpy(), atol=1e-12)


def test_pairs_step_grads_match_autograd():
    rng = np.random.RandomState(4)
    nu, ni, d, B, k = 15, 8, 5, 6, 3
    EU, EV = rng.normal(size=(nu, d)), rng.normal(size=(ni, d))
    n = (1 + k) * B
    uid, cid = rng.randint(0, nu, size=n), rng.randint(0, ni, size=n)
    for loss in LOSSES:
        ref = O.step_mul(EU, EV, uid, cid, B, k, loss, 128.0, 0.7, u_reg=1e-2, norm_u=True, norm_v=False)
        tU = torch.tensor(EU, requires_grad=True); tV = torch.tensor(EV, requires_grad=True)
        Ur = tU[torch.as_tensor(uid)]
        s = torch.sum(torch.nn.functional.normalize(Ur, dim=-1, eps=1e-6) * tV[torch.as_tensor(cid)], dim=1)
        L = t_original(s, B, k, loss, 128.0, 0.7) + 1e-2 * torch.sum(torch.mean(Ur ** 2, dim=0))
        L.backward()
        assert abs(ref["loss"] - L.item()) < 1e-12 * max(1, abs(L.item()))
        np.testing.assert_allclose(ref["dEU"], tU.grad.numpy(), atol=1e-12)
        np.testing.assert_allclose(ref["dEV"], tV.grad.numpy(), atol=1e-12)


def test_meanpool_grad_matches_autograd():
    rng = np.random.RandomState(5)
    vocab, dw, n, L = 30, 6, 10, 9
    W = rng.normal(size=(vocab, dw)); C = rng.randint(0, vocab, size=(n, L))
    dX = rng.normal(size=(n, dw))
    tW = torch.tensor(W, requires_grad=True)
    X = tW[torch.as_tensor(C)].mean(dim=1)
    np.testing.assert_allclose(O.meanpool_fwd(W, C), X.detach().numpy(), atol=1e-14)
    (X * torch.tensor(dX)).sum().backward()
    np.testing.assert_allclose(O.meanpool_bwd(W.shape, C, dX), tW.grad.numpy(), atol=1e-13)


def test_unique_first_occurrence_semantics():
    u, x = O.unique_first_occurrence(np.array([5, 3, 5, 9, 3, 3, 1]))
    np.testing.assert_array_equal(u, [5, 3, 9, 1])          # tf.unique: order of first occurrence
    np.testing.assert_array_equal(x, [0, 1, 0, 2, 1, 1, 3])


# ------------------------------------------------------------------------------------------------
# sampler restatement pinned against the real compiled reference (oracle/_ref)
# ------------------------------------------------------------------------------------------------
def _ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libnodesampler_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libnodesampler_ref.so not built (make -C oracle; needs /root/reference)")
    ref = ctypes.CDLL(so)
    ref.ref_sampler_create.restype = ctypes.c_void_p
    ref.ref_sampler_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_ulonglong]
    ref.ref_sampler_sample_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return ref


@pytest.mark.parametrize("power", [1.0, 0.75])
def test_sampler_restatement_bit_exact_vs_compiled_reference(power):
    """The reference object is { int* neg_table; int num_vertices; uint64 seed; double power } (nodesampler.cpp:17-20).
    Read its 1e8-entry table and set its LCG seed through the handle, then compare with the NumPy restatement:
    table bit-exact, draws bit-exact."""
    ref = _ref_lib()
    rng = np.random.RandomState(0)
    deg = rng.randint(0, 40, size=1000).astype(np.float64)
    h = ref.ref_sampler_create(deg.ctypes.data_as(ctypes.c_void_p), deg.size, power, 0)
    table_ptr = ctypes.cast(h, ctypes.POINTER(ctypes.c_void_p))[0]
    table = np.ctypeslib.as_array(ctypes.cast(table_ptr, ctypes.POINTER(ctypes.c_int32)), shape=(10 ** 8,))
    mine = O.sampler_build_table(deg, power, 10 ** 8)
    assert mine.shape == table.shape
    # pow()/cumsum in NumPy vs libm may differ in the last ulp for non-integer powers: allow boundary shifts of 1 slot
    diff = np.nonzero(mine != table)[0]
    assert diff.size <= (0 if power == 1.0 else 2000), diff.size
    if diff.size:
        assert np.all(np.abs(mine[diff].astype(np.int64) - table[diff]) <= np.abs(np.diff(np.nonzero(deg)[0])).max())
    # draws: force the seed (offset 16: int* (8) + int (4) + pad (4))
    seed_ptr = ctypes.cast(h + 16, ctypes.POINTER(ctypes.c_uint64))
    seed_ptr[0] = 20171017
    out = np.zeros(5000, dtype=np.int32)
    ref.ref_sampler_sample_batch(h, out.size, out.ctypes.data_as(ctypes.c_void_p))
    idx, _ = O.sampler_lcg_indices(20171017, out.size, 10 ** 8)
    np.testing.assert_array_equal(out, table[idx])
    # and the target distribution the device sampler is validated against
    p = O.sampler_probabilities(deg, power)
    emp = np.bincount(table, minlength=deg.size) / 1e8
    assert np.max(np.abs(emp - p)) < 2e-8 + 1e-6 * p.max()
    assert np.all(emp[deg == 0] == 0)


def test_degree_histogram_and_uniform():
    train = np.array([[0, 2, 1], [1, 2, 1], [2, 0, 1], [0, 4, 1]])
    np.testing.assert_array_equal(O.degree_histogram(train, 1), [1, 0, 2, 0, 1])
    np.testing.assert_array_equal(O.degree_histogram(train, 1, "uniform_no_correction"), [1, 0, 1, 0, 1])


# ------------------------------------------------------------------------------------------------
# metrics on hand cases (utils/metrics_ranking.py)
# ------------------------------------------------------------------------------------------------
def test_eval_multiple_hand_cases():
    truth = np.array([1, 0, 1, 0, 0, 1]); pred = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.1])
    ap, rc, pr = O.eval_multiple(truth, pred, 3)
    # ranks 1 and 3 hit: AP = (1/1 + 2/3) / min(3 hits, k=3)
    assert ap == pytest.approx((1.0 + 2.0 / 3.0) / 3.0)
    assert rc == pytest.approx(2.0 / 3.0) and pr == pytest.approx(2.0 / 3.0)
    assert O.eval_multiple(np.zeros(6), pred, 3) == (0.0, 0.0, 0.0)
    ap2, rc2, pr2 = O.eval_multiple_original(truth, pred, -1)
    assert rc2 == 1.0 and pr2 == pytest.approx(0.5)
    assert ap2 == pytest.approx((1.0 + 2.0 / 3.0 + 3.0 / 6.0) / 3.0)
    # ties: lowest index first
    np.testing.assert_array_equal(O.topk_indices(np.array([1.0, 2.0, 2.0, 0.0, 2.0]), 3), [1, 2, 4])


def test_auc_matches_sklearn():
    from sklearn.metrics import roc_auc_score
    rng = np.random.RandomState(0)
    for _ in range(5):
        t = rng.randint(0, 2, size=40); t[0] = 1; t[1] = 0
        s = np.round(rng.normal(size=40), 1)        # rounded => ties
        assert O.auc_score(t, s) == pytest.approx(roc_auc_score(t, s), abs=1e-12)


def test_group_shuffle_properties():
    rng = np.random.RandomState(0)
    train = np.stack([rng.randint(0, 30, 500), rng.randint(0, 20, 500), np.ones(500, dtype=int)], 1)
    out = O.group_shuffle_train(train.copy(), "item", 4, np.arange(20), np.random.RandomState(1))
    assert sorted(map(tuple, out)) == sorted(map(tuple, train))            # a permutation of the rows
    # without chop the item column is grouped
    out0 = O.group_shuffle_train(train.copy(), "item", 0, np.arange(20), np.random.RandomState(1))
    items = out0[:, 1]
    changes = np.count_nonzero(np.diff(items))
    assert changes == len(set(items)) - 1
    # same stream, legacy shuffle: reference call order reproduces with permutation arrays
    rs = np.random.RandomState(9)
    ii, rp, bp = O.group_shuffle_perms(500, 20, 4, rs)
    exp = O.group_shuffle_train(train.copy(), "item", 4, np.arange(20), np.random.RandomState(9))
    t2 = train[rp]
    t2 = t2[np.argsort(ii[t2[:, 1]], kind="stable")]
    t2 = t2.reshape(-1, 4, 3)[bp].reshape(-1, 3)
    np.testing.assert_array_equal(t2, exp)


# ------------------------------------------------------------------------------------------------
# committed golden fixtures (tests/golden/, written by tests/golden/make_golden.py)
# ------------------------------------------------------------------------------------------------
def test_golden_fixtures():
    path = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    for i, case in enumerate(meta["loss_cases"]):
        S = g["S_%d" % i]
        if case["scheme"] == "neg_shared":
            L, G = O.neg_shared_loss_grad(S, case["loss"], case["lam"], case["gamma"])
        else:
            L, G = O.group_neg_shared_loss_grad(S, g["pos_%d" % i], case["loss"], case["lam"], case["gamma"])
        assert L == pytest.approx(float(g["L_%d" % i]), rel=1e-12)
        np.testing.assert_allclose(G, g["G_%d" % i], rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(
        O.group_shuffle_train(g["gs_train"].copy(), "item", 4, np.arange(int(g["gs_nkeys"])), np.random.RandomState(2017)),
        g["gs_out"])
    np.testing.assert_array_equal(O.sampler_build_table(g["sm_deg"], 0.75, 10 ** 5), g["sm_table"])
    r = O.evaluate_mat(g["ev_truth"], g["ev_pred"], 10)
    assert r["map"] == pytest.approx(float(g["ev_map"]), rel=1e-12)
    assert r["recall"] == pytest.approx(float(g["ev_recall"]), rel=1e-12)


@pytest.mark.parametrize("loss", ["skip-gram", "mse", "log-loss", "max-margin"])
@pytest.mark.parametrize("norm", [False, True])
def test_cpu_baseline_step_equals_checker(loss, norm):
    """bench.py's CPU arm (float32, sparse, in place, R batches on one snapshot) against the fp64 checker step_matmul"""
    rng = np.random.RandomState(5)
    nu, ni, B, d, R, lr = 90, 70, 24, 16, 3, 0.1
    EU = rng.uniform(-0.5, 0.5, size=(nu, d)).astype(np.float32); EV = rng.uniform(-0.5, 0.5, size=(ni, d)).astype(np.float32)
    uid = rng.randint(0, nu, size=R * B); cid = rng.randint(0, ni, size=R * B)
    lam, gamma = (8.0, 10.0) if loss == "mse" else (128.0, 0.1 if loss == "max-margin" else 10.0)
    dU = np.zeros((nu, d)); dV = np.zeros((ni, d)); L = 0.0
    for r in range(R):
        ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid[r * B:(r + 1) * B], cid[r * B:(r + 1) * B],
                            "neg_shared", loss, lam, gamma, u_reg=1e-2, norm_u=norm, norm_v=norm)
        dU += ref["dEU"]; dV += ref["dEV"]; L += ref["loss"] / R
    U1, V1 = EU.copy(), EV.copy()
    got = O.baseline_neg_shared_step(U1, V1, uid, cid, loss, lam, gamma, lr, u_reg=1e-2, norm=norm, replicas=R)
    assert abs(got - L) <= 1e-4 * abs(L)
    np.testing.assert_allclose(U1 - EU, -lr * dU, rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(V1 - EV, -lr * dV, rtol=2e-3, atol=2e-6)
