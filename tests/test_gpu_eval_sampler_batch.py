"""GPU parity: whole@k top-k + metrics, given@-1, sampler, batch builders, mean-pool encoder — CUDA path through the
C-ABI against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import nncf_oracle as O

pytestmark = pytest.mark.gpu


def _int_embeddings(n, d, seed):
    """integer-valued entries in {-4..4}: every dot product is exact in bf16/fp32, so ids must be bit-exact, ties included"""
    return np.random.RandomState(seed).randint(-4, 5, size=(n, d)).astype(np.float32)


@pytest.mark.parametrize("gen", ["3", "2", "1"])
@pytest.mark.parametrize("nu,ni,d,k", [(300, 2500, 50, 50), (700, 9000, 128, 10), (385, 4100, 96, 32), (260, 1500, 200, 40),
                                       (1100, 20000, 128, 64), (130, 1000, 128, 100), (600, 40000, 128, 100)])
def test_topk_generations_bit_exact(gen, nu, ni, d, k, monkeypatch):
    """the three tensor-core generations of whole@k (CTA pair with readers / selectors = default, one CTA per 128 users,
    streaming exact sets) on tie-heavy integer embeddings: odd numbers of user blocks (a padded pair), ragged item tiles,
    several item splits, d = 200 (NSUB = 4), k = 100 (falls back to the second generation): ids and scores bit-exact"""
    from nncf_b200.ops import eval_topk
    if gen == "1":
        monkeypatch.setenv("NNCF_EVAL_V1", "1")
    else:
        monkeypatch.setenv("NNCF_EVAL_GEN", gen)
    U, V = _int_embeddings(nu, d, 11), _int_embeddings(ni, d, 12)
    ids, sc = eval_topk(torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda(), k, "bf16")
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    S = U.astype(np.float64) @ V.astype(np.float64).T
    for u in range(0, nu, 3):
        exp = O.topk_indices(S[u], k)
        np.testing.assert_array_equal(ids[u], exp)
        np.testing.assert_array_equal(sc[u], S[u, exp].astype(np.float32))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("nu,ni,d,k", [(37, 500, 16, 10), (300, 2500, 50, 50), (130, 1000, 128, 100), (64, 40, 8, 50),
                                       (5, 129, 200, 128)])
def test_topk_ids_bit_exact_on_integer_embeddings(precision, nu, ni, d, k):
    from nncf_b200.ops import eval_topk
    from nncf_b200._lib import NNCFError
    U, V = _int_embeddings(nu, d, 1), _int_embeddings(ni, d, 2)
    if precision == "bf16" and d > 128 and k > 64:
        # declared limit of the tensor-core kernel: the per-row top-k sets (128 rows x k keys) share the SM's shared
        # memory with the operand tiles; k > 64 next to dim > 128 operands does not fit and fails loudly
        with pytest.raises(NNCFError):
            eval_topk(torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda(), k, precision)
        return
    ids, sc = eval_topk(torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda(), k, precision)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    S = U.astype(np.float64) @ V.astype(np.float64).T
    kk = min(k, ni)
    for u in range(nu):
        exp = O.topk_indices(S[u], kk)
        np.testing.assert_array_equal(ids[u, :kk], exp)
        np.testing.assert_array_equal(sc[u, :kk], S[u, exp].astype(np.float32))
        assert np.all(ids[u, kk:] == -1)      # fewer candidates than k: the tail is marked empty


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_whole_eval_metrics_match_oracle(precision, tol):
    from nncf_b200.ops import eval_topk, eval_metrics
    rng = np.random.RandomState(5)
    nu, ni, d, k = 400, 3000, 50, 50
    U = rng.normal(size=(nu, d)).astype(np.float32) / np.sqrt(d)
    V = rng.normal(size=(ni, d)).astype(np.float32) / np.sqrt(d)
    truth = (rng.uniform(size=(nu, ni)) < 0.004).astype(np.int32)
    truth[:7] = 0                                  # users without any relevant candidate are dropped
    # make truth correlated with the scores so the metrics are not ~0
    S = U.astype(np.float64) @ V.astype(np.float64).T
    top = np.argsort(-S, axis=1)[:, :5]
    for u in range(7, nu):
        truth[u, top[u, rng.randint(0, 5)]] = 1
    ref = O.evaluate_mat(truth, S, k)
    indptr = np.concatenate([[0], np.cumsum(truth.sum(1))]).astype(np.int64)
    cols = np.nonzero(truth)[1].astype(np.int32)
    ids, _ = eval_topk(torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda(), k, precision)
    per_user, sums = eval_metrics(ids, torch.from_numpy(indptr).cuda(), torch.from_numpy(cols).cuda())
    sums = sums.cpu().numpy()
    assert int(sums[3]) == ref["n_users"]
    assert abs(sums[0] / sums[3] - ref["map"]) <= tol * max(ref["map"], 1e-9)
    assert abs(sums[1] / sums[3] - ref["recall"]) <= tol * max(ref["recall"], 1e-9)
    assert abs(sums[2] / sums[3] - ref["precision"]) <= tol * max(ref["precision"], 1e-9)
    if precision == "fp32":
        pu = per_user.cpu().numpy()
        for u in [0, 7, 8, 100, nu - 1]:
            a, r, p = O.eval_multiple(truth[u], S[u], k) if truth[u].sum() else (0, 0, 0)
            np.testing.assert_allclose(pu[u], [a, r, p], rtol=1e-5, atol=1e-7)


def test_topk_large_property_check():
    """At a larger size the oracle is too slow; check size-independent properties instead: scores sorted, every
    returned score equals the fp32 dot of its (user, id), and no non-returned item of a sampled user beats the k-th."""
    from nncf_b200.ops import eval_topk
    g = torch.Generator(device="cuda").manual_seed(1)
    nu, ni, d, k = 1000, 200_000, 128, 100
    U = torch.randn((nu, d), device="cuda", generator=g) / d ** 0.5
    V = torch.randn((ni, d), device="cuda", generator=g) / d ** 0.5
    ids, sc = eval_topk(U, V, k, "bf16")
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    assert bool(((ids >= 0) & (ids < ni)).all())
    Ub, Vb = U.bfloat16().float(), V.bfloat16().float()
    for u in [0, 17, 999]:
        s_all = Vb @ Ub[u]
        got = s_all[ids[u].long()]
        assert torch.allclose(got, sc[u], rtol=1e-4, atol=1e-5)
        kth = sc[u, -1]
        mask = torch.ones(ni, dtype=torch.bool, device="cuda"); mask[ids[u].long()] = False
        assert float(s_all[mask].max()) <= float(kth) + 1e-4


@pytest.mark.parametrize("topk", [-1, 5, 50])
def test_given_eval_matches_oracle(topk):
    from nncf_b200.ops import score_pairs, eval_given
    rng = np.random.RandomState(3)
    nu, ni, d = 50, 80, 50
    U = rng.normal(size=(nu, d)).astype(np.float32); V = rng.normal(size=(ni, d)).astype(np.float32)
    pairs = []
    for u in range(nu):
        m = rng.randint(4, 30)
        its = rng.choice(ni, size=m, replace=False)
        t = (rng.uniform(size=m) < 0.3).astype(np.int64); t[0] = 1; t[1] = 0
        pairs += [(u, i, tt) for i, tt in zip(its, t)]
    pairs = np.array(pairs, dtype=np.int64)
    ref = O.given_eval(U.astype(np.float64), V.astype(np.float64), pairs, topk)   # k = 50 > every list: ranked whole
    tU, tV = torch.from_numpy(U).cuda(), torch.from_numpy(V).cuda()
    sc = score_pairs(tU, tV, torch.from_numpy(pairs[:, 0]).cuda(), torch.from_numpy(pairs[:, 1]).cuda())
    np.testing.assert_allclose(sc.cpu().numpy(), np.sum(U[pairs[:, 0]] * V[pairs[:, 1]], axis=1), rtol=1e-5, atol=1e-5)
    counts = np.bincount(pairs[:, 0], minlength=nu)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    out = eval_given(sc, torch.from_numpy(pairs[:, 2]).cuda(), torch.from_numpy(indptr).cuda(), topk).cpu().numpy()
    assert abs(out[:, 0].mean() - ref["map"]) < 1e-5
    assert abs(out[:, 1].mean() - ref["auc"]) < 1e-5
    assert abs(out[:, 2].mean() - ref["recall"]) < 1e-5
    assert abs(out[:, 3].mean() - ref["precision"]) < 1e-5


# ------------------------------------------------------------------------------------------------ sampler
@pytest.mark.parametrize("power", [1.0, 0.75])
def test_sampler_distribution_chi_square(power):
    from nncf_b200.ops import DeviceSampler
    rng = np.random.RandomState(0)
    n = 2000
    deg = np.floor(rng.pareto(1.2, size=n) * 3).astype(np.float64)
    deg[::7] = 0                                      # zero-degree ids must never be sampled
    p = O.sampler_probabilities(deg, power)
    s = DeviceSampler(deg, power, seed=1234)
    N = 4_000_000
    draws = s.sample_device(N).cpu().numpy()
    assert draws.dtype == np.int32 and draws.min() >= 0 and draws.max() < n
    counts = np.bincount(draws, minlength=n).astype(np.float64)
    assert np.all(counts[deg == 0] == 0)
    nz = p > 0
    chi2 = np.sum((counts[nz] - N * p[nz]) ** 2 / (N * p[nz]))
    dof = nz.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)       # ~6 sigma


def test_sampler_stream_is_counter_based_and_reproducible():
    from nncf_b200.ops import DeviceSampler
    deg = np.arange(1, 101, dtype=np.float64)
    a = DeviceSampler(deg, 0.75, seed=7); b = DeviceSampler(deg, 0.75, seed=7); c = DeviceSampler(deg, 0.75, seed=8)
    x = a.sample_host(1000)
    y = np.concatenate([b.sample_host(300), b.sample_host(700)])
    np.testing.assert_array_equal(x, y)                 # same key, consecutive counters => same stream
    assert not np.array_equal(x, c.sample_host(1000))
    b.seek(300)
    np.testing.assert_array_equal(b.sample_host(700), x[300:])
    assert a.sample_host(0).shape == (0,)


def test_sampler_alias_table_is_exact():
    from nncf_b200.ops import DeviceSampler
    deg = np.array([0, 5, 3, 0, 2, 10, 1], dtype=np.float64)
    s = DeviceSampler(deg, 1.0, seed=0)
    prob, alias = s.export_table()
    n = deg.size
    p = np.zeros(n)
    for i in range(n):
        p[i] += prob[i] / n
        p[alias[i]] += (1.0 - prob[i]) / n
    np.testing.assert_allclose(p, O.sampler_probabilities(deg, 1.0), atol=1e-6)


def test_sampler_matches_compiled_reference_distribution():
    """The real reference sampler (oracle/_ref, built from sampler/nodesampler.cpp) and the device sampler draw
    from the same distribution: two-sample chi-square on 2M draws each."""
    import ctypes, os
    so = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "libnodesampler_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    from nncf_b200.ops import DeviceSampler
    ref = ctypes.CDLL(so)
    ref.ref_sampler_create.restype = ctypes.c_void_p
    ref.ref_sampler_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_ulonglong]
    ref.ref_sampler_sample_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rng = np.random.RandomState(1)
    n = 500
    deg = rng.randint(0, 50, size=n).astype(np.float64)
    h = ref.ref_sampler_create(deg.ctypes.data_as(ctypes.c_void_p), n, 0.75, 0)
    N = 2_000_000
    r = np.zeros(N, dtype=np.int32)
    ref.ref_sampler_sample_batch(h, N, r.ctypes.data_as(ctypes.c_void_p))
    g = DeviceSampler(deg, 0.75, seed=99).sample_device(N).cpu().numpy()
    cr = np.bincount(r, minlength=n).astype(np.float64); cg = np.bincount(g, minlength=n).astype(np.float64)
    assert np.all(cr[deg == 0] == 0) and np.all(cg[deg == 0] == 0)
    nz = (cr + cg) > 0
    chi2 = np.sum((cr[nz] - cg[nz]) ** 2 / (cr[nz] + cg[nz]))
    dof = nz.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)


# ------------------------------------------------------------------------------------------------ batch builders
@pytest.mark.parametrize("n,n_items,chop,by", [(1000, 37, 4, "item"), (5000, 700, 2, "item"), (4097, 300, 3, "user"),
                                               (50, 5, 0, "item"), (70000, 70000, 4, "item"), (3, 2, 4, "item")])
def test_group_shuffle_bit_exact(n, n_items, chop, by):
    from nncf_b200.ops import group_shuffle
    rng = np.random.RandomState(n)
    train = np.stack([rng.randint(0, 91, size=n), rng.randint(0, n_items, size=n), np.ones(n, dtype=np.int64)], 1)
    col = 0 if by == "user" else 1
    n_keys = int(train[:, col].max()) + 1
    # oracle = the reference function (stable sort) driven by a seeded legacy stream
    iidx = np.arange(n_keys)
    exp = O.group_shuffle_train(train.copy(), by=by, chop=chop, iidx=iidx, rng=np.random.RandomState(42))
    # product: same stream, same draw order, permutations as index arrays
    rs = np.random.RandomState(42)
    iidx_perm, row_perm, block_perm = O.group_shuffle_perms(n, n_keys, chop, rs)
    np.testing.assert_array_equal(iidx_perm, iidx)       # the oracle shuffled iidx in place with the same draws
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    got = group_shuffle(t(train.astype(np.int32)), col, t(iidx_perm), t(row_perm), t(block_perm), chop).cpu().numpy()
    np.testing.assert_array_equal(got, exp.astype(np.int32))


def test_permute_rows_and_assemble_pairs_bit_exact():
    from nncf_b200.ops import permute_rows, assemble_pairs_batch
    rng = np.random.RandomState(0)
    n = 999
    train = np.stack([rng.randint(0, 50, size=n), rng.randint(0, 70, size=n), np.ones(n, dtype=np.int64)], 1).astype(np.int32)
    exp = train.copy(); np.random.RandomState(5).shuffle(exp)
    perm = np.arange(n); np.random.RandomState(5).shuffle(perm)
    got = permute_rows(torch.from_numpy(train).cuda(), torch.from_numpy(perm).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, exp)
    B, k = 64, 5
    negs = rng.randint(0, 70, size=B * k).astype(np.int32)
    pos = exp[:B]
    np.testing.assert_array_equal(
        assemble_pairs_batch(torch.from_numpy(pos).cuda(), k, torch.from_numpy(negs).cuda(), 1, -1).cpu().numpy(),
        O.assemble_original_batch(pos.copy(), k, negs, -1))
    np.testing.assert_array_equal(
        assemble_pairs_batch(torch.from_numpy(pos).cuda(), k, torch.from_numpy(negs).cuda(), 0, 0).cpu().numpy(),
        O.assemble_group_sample_batch(pos.copy(), k, negs, 0))


def test_unique_first_occurrence_bit_exact():
    from nncf_b200.ops import unique_first_occurrence
    rng = np.random.RandomState(1)
    # (n <= 2,048: shared-memory hash table; above: the quadratic scan.  hi = 2: a table of two long chains; 2**31 - 1: ids with the top bits set)
    for n, hi in [(1, 5), (2, 1), (512, 40), (512, 100000), (511, 2), (2048, 1500), (2048, 2 ** 31 - 1), (2049, 900), (3000, 700)]:
        ids = rng.randint(0, hi, size=n).astype(np.int32)
        uq, inv, cnt = unique_first_occurrence(torch.from_numpy(ids).cuda())
        eu, ex = O.unique_first_occurrence(ids)
        c = int(cnt.item())
        assert c == eu.size
        np.testing.assert_array_equal(uq.cpu().numpy()[:c], eu)
        np.testing.assert_array_equal(inv.cpu().numpy(), ex)


# ------------------------------------------------------------------------------------------------ mean-pool
def test_meanpool_fwd_bwd_match_oracle():
    from nncf_b200.ops import meanpool_fwd, meanpool_bwd
    rng = np.random.RandomState(2)
    vocab, dw, I, L, n = 300, 50, 120, 37, 64
    W = rng.normal(size=(vocab, dw)).astype(np.float32)
    C = rng.randint(1, vocab, size=(I, L)).astype(np.int32)
    for i in range(I):
        C[i, :rng.randint(0, L)] = 0                   # left zero padding (data/readme.txt:5)
    ids = rng.randint(0, I, size=n).astype(np.int32)
    X = O.meanpool_fwd(W.astype(np.float64), C[ids])
    tW, tC, tid = torch.from_numpy(W).cuda(), torch.from_numpy(C).cuda(), torch.from_numpy(ids).cuda()
    got = meanpool_fwd(tW, tC, tid, n).cpu().numpy()
    np.testing.assert_allclose(got, X, rtol=1e-5, atol=1e-6)
    dX = rng.normal(size=(n, dw)).astype(np.float32)
    dW = torch.zeros_like(tW)
    meanpool_bwd(dW, tC, tid, n, torch.from_numpy(dX).cuda())
    ref = O.meanpool_bwd(W.shape, C[ids], dX.astype(np.float64))
    assert np.max(np.abs(dW.cpu().numpy() - ref)) <= 1e-4 * np.max(np.abs(ref))


@pytest.mark.parametrize("k", [10, 50, 100])
def test_topk_c4_item_count_property_check(k):
    """BASELINE config 4's candidate count (2M items, d = 128) with a shard of 512 users: exactness properties of the
    append + compaction top-k at full width — sorted output, exact scores, nothing outside the list beats the k-th, and
    the same ids as torch.topk on the bf16-rounded operands for sampled users (continuous scores: no ties)."""
    from nncf_b200.ops import eval_topk
    g = torch.Generator(device="cuda").manual_seed(2)
    nu, ni, d = 512, 2_000_000, 128
    U = torch.randn((nu, d), device="cuda", generator=g) / d ** 0.5
    V = torch.randn((ni, d), device="cuda", generator=g) / d ** 0.5
    ids, sc = eval_topk(U, V, k, "bf16")
    assert bool((sc[:, :-1] >= sc[:, 1:]).all()) and bool(((ids >= 0) & (ids < ni)).all())
    Ub, Vb = U.bfloat16().float(), V.bfloat16().float()
    for u in [0, 101, 511]:
        s_all = Vb @ Ub[u]
        ref = torch.topk(s_all, k)
        # fp32 accumulation order differs between the tensor core and torch: allow swaps between scores closer than 1e-5
        assert torch.allclose(sc[u], ref.values, rtol=1e-4, atol=1e-5)
        same = (ids[u].long() == ref.indices)
        assert bool(same.all()) or float((sc[u][~same] - ref.values[~same]).abs().max()) < 1e-5


# ------------------------------------------------------------------------------------------------ content-tower tail
@pytest.mark.parametrize("use_bn", [True, False])
@pytest.mark.parametrize("act", ["relu", "tanh", "linear"])
@pytest.mark.parametrize("rows,n,d", [(512, 377, 50), (128, 128, 32), (64, 3, 70)])
def test_tower_bn_act_kernels_match_torch_batchnorm(rows, n, d, act, use_bn):
    """BatchNorm over the first n rows (count in device memory) + activation, forward and backward, against torch's
    BatchNorm1d (training mode, Keras-1 defaults eps 1e-3 / momentum 0.99 -> torch momentum 0.01) under autograd."""
    from nncf_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    h = torch.randn((rows, d), device="cuda", generator=g) * 0.7 + 0.3
    dy = torch.randn((rows, d), device="cuda", generator=g)
    nv = torch.tensor([n], dtype=torch.int32, device="cuda")
    bn = torch.nn.BatchNorm1d(d, eps=1e-3, momentum=0.01).cuda() if use_bn else None
    ref_bn = torch.nn.BatchNorm1d(d, eps=1e-3, momentum=0.01).cuda() if use_bn else None
    if use_bn:
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
            ref_bn.weight.copy_(bn.weight); ref_bn.bias.copy_(bn.bias)
    hr = h[:n].clone().requires_grad_(True)
    z = ref_bn(hr) if use_bn else hr
    yr = torch.relu(z) if act == "relu" else (torch.tanh(z) if act == "tanh" else z)
    yr.backward(dy[:n])
    with torch.no_grad():
        y, xhat, rstd = ops.tower_bn_act_fwd(h, nv, bn, act)
        dh, dg, db = ops.tower_bn_act_bwd(dy, y, xhat, rstd, nv, bn, act)
    torch.cuda.synchronize()
    assert torch.allclose(y[:n], yr, rtol=1e-4, atol=1e-5)
    assert float(y[n:].abs().max()) == 0.0 if n < rows else True
    assert torch.allclose(dh[:n], hr.grad, rtol=1e-3, atol=2e-5), float((dh[:n] - hr.grad).abs().max())
    assert float(dh[n:].abs().max()) == 0.0 if n < rows else True
    if use_bn:
        assert torch.allclose(dg, ref_bn.weight.grad, rtol=1e-3, atol=2e-5) and torch.allclose(db, ref_bn.bias.grad, rtol=1e-3, atol=2e-5)
        assert torch.allclose(bn.running_mean, ref_bn.running_mean, rtol=1e-5, atol=1e-7)
        assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-5, atol=1e-7)


def test_meanpool_device_count_variant_matches_plain():
    from nncf_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    V, dw, L, n_slots, n = 900, 50, 37, 96, 61
    W = torch.randn((V, dw), device="cuda", generator=g)
    content = torch.randint(0, V, (300, L), device="cuda", generator=g, dtype=torch.int32)
    content[:, :5] = 0                                                   # 0-left-padded, as the reference's content matrix
    ids = torch.randint(0, 300, (n_slots,), device="cuda", generator=g, dtype=torch.int32)
    ids[n:] = 2_000_000_000                                              # slots beyond the count must never be dereferenced
    nv = torch.tensor([n], dtype=torch.int32, device="cuda")
    out = ops.meanpool_fwd_n(W, content, ids, nv)
    ref = ops.meanpool_fwd(W, content, ids[:n].contiguous(), n)
    assert torch.equal(out[:n], ref) and float(out[n:].abs().max()) == 0.0
    gy = torch.randn((n_slots, dw), device="cuda", generator=g)
    dW1, dW2 = torch.zeros_like(W), torch.zeros_like(W)
    ops.meanpool_bwd_n(dW1, content, ids, nv, gy)
    ops.meanpool_bwd(dW2, content, ids[:n].contiguous(), n, gy[:n].contiguous())
    torch.cuda.synchronize()
    assert torch.allclose(dW1, dW2, rtol=1e-5, atol=1e-6)
