"""
nncf_oracle.py — CPU restatement (NumPy) of the reference's sampling-and-scoring hot path.

*** TEST INFRASTRUCTURE ONLY. ***  Nothing under nncf_b200/ may import this module.  The only callers
are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, where it is
the checker (or the timed CPU baseline), never the product.

Parity status: **parity unpinned** for everything except the sampler.  The reference (/root/reference)
ships no tests, golden vectors or fixtures for this path, and its Python-2 /
Keras-1.2.2 / TensorFlow-1.0 stack cannot be installed or run offline.  Third-party arithmetic that is
not under /root/reference — Keras 1.2.2 (`Embedding`, `Dense`, `BatchNormalization`, loss averaging,
`Adam`) and TensorFlow 1.0 (`unique`, `gather`, `matmul`, `l2_normalize` eps=1e-12, `sigmoid`, `log`,
`relu`, `scatter_nd_add`), pinned only in prose at README.md:63-65 — is restated here from its published
semantics and anchored on the reference's call sites cited per function.  What IS pinned: the sampler
restatement is checked against the real compiled reference sampler (oracle/_ref, built from
sampler/nodesampler.cpp by oracle/Makefile), and every analytic gradient below is checked against
torch-CPU autograd of the forward formulas in fp64 (tests/test_oracle.py).

All formulas follow the reference file:line given in each docstring (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

LOSSES = ("skip-gram", "mse", "log-loss", "max-margin")


# ---------------------------------------------------------------------------------------------
# elementary pieces
# ---------------------------------------------------------------------------------------------
def sigmoid(x):
    x = np.asarray(x)
    out = np.empty_like(x, dtype=np.result_type(x, np.float32))
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def log_sigmoid(x):
    """log(sigmoid(x)); the reference uses the naive K.log(K.sigmoid(.)) (utils/objectives.py:58,63,98,104);
    this is its mathematically identical, overflow-safe form."""
    x = np.asarray(x)
    return -(np.maximum(-x, 0) + np.log1p(np.exp(-np.abs(x))))


def l2_normalize(x, eps=1e-12):
    """tf.nn.l2_normalize(x, dim=-1) as called at models/model_framework.py:63,110-111:
    x * rsqrt(max(sum(x^2), eps))."""
    ss = np.sum(x * x, axis=-1, keepdims=True)
    inv = 1.0 / np.sqrt(np.maximum(ss, eps))
    return x * inv, inv


def l2_normalize_bwd(xhat, inv, dxhat):
    """Backward of l2_normalize: dx = (dxhat - xhat*(xhat.dxhat)) * inv."""
    dot = np.sum(xhat * dxhat, axis=-1, keepdims=True)
    return (dxhat - xhat * dot) * inv


def unique_first_occurrence(ids):
    """tf.unique(x) -> (y, idx) with y in order of first occurrence (models/model_framework.py:45-48)."""
    ids = np.asarray(ids)
    _, first_idx, inv = np.unique(ids, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")          # sorted-unique slot -> first-occurrence rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    cid_u = ids[np.sort(first_idx)]
    cid_x = rank[inv.reshape(-1)].astype(np.int32)
    return cid_u.astype(ids.dtype), cid_x


# ---------------------------------------------------------------------------------------------
# losses + dL/dScore  (utils/objectives.py)
# ---------------------------------------------------------------------------------------------
def neg_shared_loss_grad(S, loss, neg_loss_weight, gamma):
    """get_neg_shared_loss, utils/objectives.py:78-117 (+ tensorflow_diag utils/utilities.py:74-83).
    S: [B,B], diagonal = positives.  Returns (scalar loss, dL/dS)."""
    S = np.asarray(S)
    B = S.shape[0]
    assert S.shape == (B, B)
    I = np.eye(B, dtype=S.dtype)
    if loss in ("skip-gram", "mse"):
        w = neg_loss_weight / (B - 1.0)                              # :100, :107
        W = I * (1.0 - w) + w                                        # :101-102
        if loss == "skip-gram":
            Y = 2.0 * I - 1.0                                        # :103
            L = np.sum(-W * log_sigmoid(Y * S)) / B                  # :104-105
            G = W * (sigmoid(S) - I) / B
        else:
            L = np.sum(W * (S - I) ** 2) / B                         # :108-112
            G = 2.0 * W * (S - I) / B
        return L, G
    diag = np.diagonal(S)                                            # K.diag -> shape (B,)
    D = diag[None, :] - S                                            # :92,:97  (B,) - (B,B): D[i,j]=S[j,j]-S[i,j]
    if loss == "log-loss":
        L = np.mean(-log_sigmoid(gamma * D))                         # :98
        A = -gamma * sigmoid(-gamma * D) / (B * B)
    elif loss == "max-margin":
        M = gamma * (1.0 - I)                                        # :93
        L = np.mean(np.maximum(M - D, 0.0))                          # :94
        A = -((M - D) > 0).astype(S.dtype) / (B * B)
    else:
        raise AssertionError("[ERROR!] loss %s not specified." % loss)
    G = -A + I * np.sum(A, axis=0)[None, :]
    return L, G


def group_neg_shared_loss_grad(P, pos_col, loss, neg_loss_weight, gamma):
    """get_group_neg_shared_loss, utils/objectives.py:163-220.  P: [B,n_u]; positives at (i, pos_col[i])
    (pos_idxs, models/model_framework.py:133-134).  Returns (loss, dL/dP)."""
    P = np.asarray(P)
    B, nu = P.shape
    Yp = np.zeros_like(P)
    Yp[np.arange(B), pos_col] = 1.0
    if loss in ("skip-gram", "mse"):
        w = neg_loss_weight / (nu - 1.0)                             # :177, :203, :211
        W = w + (1.0 - w) * Yp                                       # create_mask(pos=1, neg=w) :179-192
        if loss == "skip-gram":
            Y = 2.0 * Yp - 1.0                                       # create_mask(1, -1)
            L = np.sum(-W * log_sigmoid(Y * P)) / B                  # :208-209
            G = W * (sigmoid(P) - Yp) / B
        else:
            L = np.sum(W * (P - Yp) ** 2) / B                        # :216-217
            G = 2.0 * W * (P - Yp) / B
        return L, G
    ppos = P[np.arange(B), pos_col][:, None]                         # :175
    D = ppos - P                                                     # :196, :200 (row-wise)
    if loss == "log-loss":
        L = np.mean(-log_sigmoid(gamma * D))                         # :201
        A = -gamma * sigmoid(-gamma * D) / (B * nu)
    elif loss == "max-margin":
        L = np.mean(np.maximum(gamma - D, 0.0))                      # :197 (no zero margin at the positive)
        A = -((gamma - D) > 0).astype(P.dtype) / (B * nu)
    else:
        raise AssertionError("[ERROR!] loss %s not specified." % loss)
    G = -A + Yp * np.sum(A, axis=1, keepdims=True)
    return L, G


def original_loss_grad(s, B, k, loss, neg_loss_weight, gamma):
    """get_original_loss, utils/objectives.py:35-75.  s: [(1+k)B]; first B positives, then k consecutive
    negatives per positive (models/train_original.py:50-53).  y_true is implied by the scheme:
    +1 for positives; -1 (skip-gram) or 0 (mse) for negatives (train_original.py:13-14).  Returns (loss, dL/ds)."""
    s = np.asarray(s).reshape(-1)
    assert s.size == (1 + k) * B
    sp, sn = s[:B], s[B:]
    g = np.zeros_like(s)
    if loss == "skip-gram":
        w = neg_loss_weight / k                                      # :60
        L = (np.sum(-log_sigmoid(sp)) + w * np.sum(-log_sigmoid(-sn))) / B     # :61-64
        g[:B] = (sigmoid(sp) - 1.0) / B
        g[B:] = w * sigmoid(sn) / B
    elif loss == "mse":
        w = neg_loss_weight / k                                      # :66
        L = (np.sum((1.0 - sp) ** 2) + w * np.sum(sn ** 2)) / B      # :67-70
        g[:B] = -2.0 * (1.0 - sp) / B
        g[B:] = 2.0 * w * sn / B
    else:
        d = np.repeat(sp, k) - sn                                    # :51-52, :56-57
        if loss == "log-loss":
            L = np.mean(-log_sigmoid(gamma * d))                     # :58
            a = -gamma * sigmoid(-gamma * d) / (k * B)
        elif loss == "max-margin":
            L = np.mean(np.maximum(gamma - d, 0.0))                  # :53
            a = -((gamma - d) > 0).astype(s.dtype) / (k * B)
        else:
            raise AssertionError("[ERROR!] loss %s not specified." % loss)
        g[B:] = -a
        g[:B] = a.reshape(B, k).sum(axis=1)
    return L, g


# ---------------------------------------------------------------------------------------------
# full training-step forward/backward on embedding tables ('mf' item tower)
# ---------------------------------------------------------------------------------------------
def _scatter_rows(n_rows, ids, rows):
    out = np.zeros((n_rows, rows.shape[1]), dtype=rows.dtype)
    np.add.at(out, ids, rows)
    return out


def step_matmul(EU, EV, uid, cid, scheme, loss, neg_loss_weight, gamma, u_reg=0.0,
                norm_u=False, norm_v=False):
    """One neg_shared / group_neg_shared batch on plain embedding tables.
    Graph: models/model_framework.py:40-65 (user side), :85-88 / :98-111 (item side), :126-136 (scores),
    modules/interaction/interaction_dot.py:100-107 ('matmul'), loss per scheme, activity regulariser
    utils/utilities.py:122-135 on the un-normalised user rows.
    Returns dict(loss, task_loss, dEU [U,d], dEV [I,d], S)."""
    uid = np.asarray(uid); cid = np.asarray(cid)
    B = uid.shape[0]
    U_raw = EU[uid]
    U, inv_u = l2_normalize(U_raw) if norm_u else (U_raw, None)
    if scheme == "neg_shared":
        V_raw = EV[cid]
        V, inv_v = l2_normalize(V_raw) if norm_v else (V_raw, None)
        S = U @ V.T
        L, G = neg_shared_loss_grad(S, loss, neg_loss_weight, gamma)
        col_ids = cid
    elif scheme == "group_neg_shared":
        cid_u, cid_x = unique_first_occurrence(cid)
        V_raw = EV[cid_u]
        V, inv_v = l2_normalize(V_raw) if norm_v else (V_raw, None)
        S = U @ V.T
        L, G = group_neg_shared_loss_grad(S, cid_x, loss, neg_loss_weight, gamma)
        col_ids = cid_u
    else:
        raise AssertionError(scheme)
    dU = G @ V
    dV = G.T @ U
    if norm_u:
        dU = l2_normalize_bwd(U, inv_u, dU)
    if norm_v:
        dV = l2_normalize_bwd(V, inv_v, dV)
    reg = u_reg * np.sum(np.mean(U_raw ** 2, axis=0))
    dU = dU + 2.0 * u_reg * U_raw / B
    return dict(loss=L + reg, task_loss=L, S=S,
                dEU=_scatter_rows(EU.shape[0], uid, dU), dEV=_scatter_rows(EV.shape[0], col_ids, dV))


def step_mul(EU, EV, uid, cid, B, k, loss, neg_loss_weight, gamma, u_reg=0.0, norm_u=False, norm_v=False):
    """One 'original' / 'group_sample' batch: (1+k)B listed pairs, row-wise dot
    (modules/interaction/interaction_dot.py:92-99), get_original_loss.  The regulariser averages over all
    (1+k)B gathered user rows (utils/utilities.py:129-135: K.mean over axis 0 of the layer output)."""
    uid = np.asarray(uid); cid = np.asarray(cid)
    n = uid.shape[0]
    assert n == (1 + k) * B
    U_raw, V_raw = EU[uid], EV[cid]
    U, inv_u = l2_normalize(U_raw) if norm_u else (U_raw, None)
    V, inv_v = l2_normalize(V_raw) if norm_v else (V_raw, None)
    s = np.sum(U * V, axis=1)
    L, g = original_loss_grad(s, B, k, loss, neg_loss_weight, gamma)
    dU = g[:, None] * V
    dV = g[:, None] * U
    if norm_u:
        dU = l2_normalize_bwd(U, inv_u, dU)
    if norm_v:
        dV = l2_normalize_bwd(V, inv_v, dV)
    reg = u_reg * np.sum(np.mean(U_raw ** 2, axis=0))
    dU = dU + 2.0 * u_reg * U_raw / n
    return dict(loss=L + reg, task_loss=L, s=s,
                dEU=_scatter_rows(EU.shape[0], uid, dU), dEV=_scatter_rows(EV.shape[0], cid, dV))


# ---------------------------------------------------------------------------------------------
# mean-of-word-vectors item encoder (modules/content/mean_pool.py)
# ---------------------------------------------------------------------------------------------
def meanpool_fwd(W, content):
    """AverageEmbeddings.call, modules/content/mean_pool.py:27-33: the mask is `content != -1`, so with the
    0-padded C matrix (data/readme.txt:5) every one of the L positions counts, pad id 0 included."""
    c = (content != -1).astype(W.dtype)
    cnt = np.sum(c, axis=1, keepdims=True)
    return np.sum(W[content], axis=1) / cnt


def meanpool_bwd(W_shape, content, dX, dtype=np.float64):
    """Gradient of meanpool_fwd w.r.t. the word table: each position scatters dX[n]/L into row content[n,l]."""
    n, L = content.shape
    c = (content != -1)
    cnt = np.sum(c, axis=1, keepdims=True).astype(dtype)
    dW = np.zeros(W_shape, dtype=dtype)
    np.add.at(dW, content.reshape(-1), np.repeat(dX / cnt, L, axis=0))
    return dW


def dense_bn_relu_fwd(X, Wd, bd, bn_gamma, bn_beta, eps=1e-3):
    """Dense -> BatchNormalization (training statistics over the batch of unique items, Keras-1 defaults
    epsilon=1e-3, axis=-1) -> relu; modules/content/mean_pool.py:81-97."""
    H = X @ Wd + bd
    mu = H.mean(axis=0)
    var = H.var(axis=0)
    Hn = (H - mu) / np.sqrt(var + eps)
    Y = Hn * bn_gamma + bn_beta
    return np.maximum(Y, 0.0), dict(H=H, mu=mu, var=var, Hn=Hn, Y=Y)


# ---------------------------------------------------------------------------------------------
# optimizers
# ---------------------------------------------------------------------------------------------
def sgd_sparse(table, grad_dense, lr):
    """Sparse SGD on the touched rows; duplicates are summed (the dense gradient is already the sum)."""
    return table - lr * grad_dense


def lazy_adam_sparse(table, m, v, ids, grad_dense, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """utils/optimizer.py:108-147 (_apply_sparse + _finish) with duplicate ids pre-summed (declared deviation:
    the reference's scatter_update is last-writer-wins on duplicates).  t is the 1-based step count."""
    ids = np.unique(np.asarray(ids))
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)       # :109-111
    g = grad_dense[ids]
    m = m.copy(); v = v.copy(); table = table.copy()
    m[ids] = beta1 * m[ids] + (1.0 - beta1) * g                       # :113-119
    v[ids] = beta2 * v[ids] + (1.0 - beta2) * g * g                   # :122-128
    table[ids] -= lr_t * m[ids] / (np.sqrt(v[ids]) + eps)             # :130-133
    return table, m, v


# ---------------------------------------------------------------------------------------------
# batch-index builders (configs/data_utils.py, models/train_*.py)
# ---------------------------------------------------------------------------------------------
def group_shuffle_train(train, by="item", chop=0, iidx=None, rng=None):
    """group_shuffle_train, configs/data_utils.py:218-241, restated with a *stable* argsort (declared tie
    rule; NumPy's default introsort is not stable) and an explicit legacy RandomState instead of the global
    np.random stream.  Consumes the stream in the reference's order: shuffle(iidx), shuffle(train rows),
    shuffle(chop blocks).  iidx is shuffled IN PLACE and persists across epochs (train_group_neg_shared.py:43-44).
    Returns the new train array (the input is also row-shuffled in place, as in the reference)."""
    rng = np.random if rng is None else rng
    col = 0 if by == "user" else 1
    if iidx is None:
        iidx = np.arange(np.max(train[:, col]) + 1)
    rng.shuffle(iidx)
    rng.shuffle(train)
    key = iidx[train[:, col]]
    train = train[np.argsort(key, kind="stable")]
    if chop > 0:
        bulk_len = (train.shape[0] // chop) * chop
        bulk = train[:bulk_len].reshape((-1, chop, train.shape[-1])).copy()
        rng.shuffle(bulk)
        train = np.vstack([bulk.reshape((-1, train.shape[-1])), train[bulk_len:]])
    return train


def group_shuffle_perms(n_rows, n_keys, chop, rng):
    """The three permutations group_shuffle_train draws, as index arrays, from the same stream in the same
    order (RandomState.shuffle of an n-element array performs the same Fisher-Yates swaps whatever the
    payload): returns (iidx_perm, row_perm, block_perm) such that shuffle(x) == x[perm]."""
    iidx_perm = np.arange(n_keys); rng.shuffle(iidx_perm)
    row_perm = np.arange(n_rows); rng.shuffle(row_perm)
    block_perm = None
    if chop > 0:
        block_perm = np.arange(n_rows // chop); rng.shuffle(block_perm)
    return iidx_perm, row_perm, block_perm


def assemble_original_batch(train_batch_p, k, neg_items, neg_sign):
    """models/train_original.py:49-56: positives first, then each positive repeated k times with column 1
    replaced by the sampled negatives and column 2 by neg_sign."""
    tb_n = train_batch_p.repeat(k, axis=0)
    tb_n[:, 1] = neg_items
    tb_n[:, 2] = neg_sign
    return np.vstack((train_batch_p, tb_n))


def assemble_group_sample_batch(train_batch_p, k, neg_users, neg_sign):
    """models/train_group_sample.py:75-85 (by == 'item' branch): negatives are sampled *users*."""
    tb_n = train_batch_p.repeat(k, axis=0)
    tb_n[:, 0] = neg_users
    tb_n[:, 2] = neg_sign
    return np.vstack((train_batch_p, tb_n))


# ---------------------------------------------------------------------------------------------
# negative sampler (sampler/nodesampler.cpp)
# ---------------------------------------------------------------------------------------------
def sampler_probabilities(dist, power):
    """Target distribution of NodeSampler::set_table, sampler/nodesampler.cpp:29-49: p_i ∝ deg_i^power for
    deg_i > 0, zero-degree ids never sampled (:34,:39)."""
    dist = np.asarray(dist, dtype=np.float64)
    w = np.where(dist > 0, np.power(np.where(dist > 0, dist, 1.0), power), 0.0)
    return w / w.sum()


def sampler_build_table(dist, power, table_size):
    """set_table restated for an arbitrary table size (the reference hard-codes 1e8, nodesampler.cpp:14):
    table[k] = i while k/table_size < cumsum_i/sum."""
    dist = np.asarray(dist, dtype=np.float64)
    nz = np.nonzero(dist)[0]
    por = np.cumsum(np.power(dist[nz], power))
    por = por / por[-1]
    # number of k with k/table_size < por_i  ==  ceil(por_i * table_size) (k integer, strict <)
    upto = np.minimum(np.ceil(por * table_size).astype(np.int64), table_size)
    upto[-1] = table_size
    counts = np.diff(np.concatenate([[0], upto]))
    return np.repeat(nz, counts).astype(np.int32)


def sampler_lcg_indices(seed, n, table_size):
    """NodeSampler::Rand, nodesampler.cpp:23-26: seed = seed*25214903917 + 11 (mod 2^64); (seed>>16) % size."""
    out = np.empty(n, dtype=np.int64)
    s = int(seed) & ((1 << 64) - 1)
    for i in range(n):
        s = (s * 25214903917 + 11) & ((1 << 64) - 1)
        out[i] = (s >> 16) % table_size
    return out, s


def degree_histogram(train, column, neg_dist="unigram"):
    """get_sampler, configs/data_utils.py:193-209: degree histogram of `column`; 'uniform' => dist[dist>0]=1;
    any suffix after '_' in neg_dist is ignored (:201)."""
    neg_dist = neg_dist.split("_")[0]
    assert neg_dist in ("uniform", "unigram"), [neg_dist]
    dist = np.bincount(train[:, column], minlength=int(np.max(train[:, column])) + 1).astype(np.float64)
    if neg_dist == "uniform":
        dist[dist > 0] = 1
    return dist


# ---------------------------------------------------------------------------------------------
# evaluation (utils/metrics_ranking.py, utils/objectives.py)
# ---------------------------------------------------------------------------------------------
def topk_indices(pred_scores, k):
    """Deterministic realisation of the reference's top-k selection + ordering
    (utils/metrics_ranking.py:7-13: argpartition then sort by (score, uniform noise) descending, i.e. ties are
    broken at random): score descending, then LOWEST column index first (declared tie rule)."""
    pred_scores = np.asarray(pred_scores)
    order = np.lexsort((np.arange(pred_scores.size), -pred_scores.astype(np.float64)))
    return order[:k]


def eval_multiple(true_scores, pred_scores, topk):
    """eval_multiple, utils/metrics_ranking.py:6-35, with the deterministic tie rule of topk_indices.
    nhits is counted over ALL candidates (:24)."""
    true_scores = np.asarray(true_scores); pred_scores = np.asarray(pred_scores)
    idx = topk_indices(pred_scores, topk)
    nhits_run = 0.0; nhits_topk = 0.0; sumap = 0.0
    k = topk if topk >= 0 else len(idx)
    for i, j in enumerate(idx):
        if true_scores[j] != 0:
            nhits_run += 1.0
            if i < k:
                nhits_topk += 1
                sumap += nhits_run / (i + 1.0)
    nhits = float(np.sum(true_scores))
    if nhits != 0:
        return sumap / min(nhits, k), nhits_topk / nhits, nhits_topk / k
    return 0.0, 0.0, 0.0


def eval_multiple_original(true_scores, pred_scores, topk):
    """eval_multiple_original, utils/metrics_ranking.py:38-61 (full sort; used when topk == -1)."""
    true_scores = np.asarray(true_scores); pred_scores = np.asarray(pred_scores)
    idx = topk_indices(pred_scores, len(pred_scores))
    k = topk if topk >= 0 else len(idx)
    nhits = 0.0; nhits_topk = 0.0; sumap = 0.0
    for i, j in enumerate(idx):
        if true_scores[j] != 0:
            nhits += 1.0
            if i < k:
                nhits_topk += 1
                sumap += nhits / (i + 1.0)
    if nhits != 0:
        return sumap / min(nhits, k), nhits_topk / nhits, nhits_topk / k
    return 0.0, 0.0, 0.0


def evaluate_mat(truth_mat, pred_mat, topk):
    """test_eval_mat + evaluate_mat, utils/objectives.py:296-321,333-370: users with no relevant candidate
    are dropped (:316), metrics are unweighted means over the kept users."""
    keep = np.sum(truth_mat, axis=1) > 0
    ap, rc, pr = [], [], []
    for u in np.nonzero(keep)[0]:
        a, r, p = (eval_multiple_original if topk == -1 else eval_multiple)(truth_mat[u], pred_mat[u], topk)
        ap.append(a); rc.append(r); pr.append(p)
    n = len(ap)
    if n == 0:
        return {"map": float("nan"), "recall": float("nan"), "precision": float("nan"), "n_users": 0}
    return {"map": float(np.mean(ap)), "recall": float(np.mean(rc)), "precision": float(np.mean(pr)), "n_users": n}


def prepare_whole_eval(train, test, user_count):
    """Evaluator._prepare_for_whole_eval, utils/objectives.py:395-421: candidate lists are the sorted unique
    item ids of train / test; dense int32 truth matrices; test truth only from label-1 rows (:414)."""
    train_items = np.array(sorted(set(train[:, 1].tolist())), dtype=np.int64)
    test_items = np.array(sorted(set(test[:, 1].tolist())), dtype=np.int64)
    tr_pos = {int(it): i for i, it in enumerate(train_items)}
    te_pos = {int(it): i for i, it in enumerate(test_items)}
    train_true = np.zeros((user_count, len(train_items)), dtype=np.int32)
    test_true = np.zeros((user_count, len(test_items)), dtype=np.int32)
    for u, it, _ in train:
        train_true[u, tr_pos[int(it)]] = 1
    for u, it, r in test:
        if r == 1:
            test_true[u, te_pos[int(it)]] = 1
    return train_items, test_items, train_true, test_true


def whole_eval(user_emb, item_emb, train, test, topk):
    """Evaluator.run(eval_scheme='whole'), utils/objectives.py:463-492, on final embeddings:
    pred = U @ V[candidates].T (model_neg_shared 'matmul' view, interaction_dot.py:100-107)."""
    user_count = user_emb.shape[0]
    tr_items, te_items, tr_true, te_true = prepare_whole_eval(train, test, user_count)
    r_train = evaluate_mat(tr_true, user_emb @ item_emb[tr_items].T, topk)
    r_test = evaluate_mat(te_true, user_emb @ item_emb[te_items].T, topk)
    return r_train, r_test


def auc_score(true_scores, pred_scores):
    """sklearn.metrics.roc_auc_score (utils/objectives.py:277) restated: Mann-Whitney U with average ranks."""
    t = np.asarray(true_scores) != 0
    s = np.asarray(pred_scores, dtype=np.float64)
    n_pos = int(t.sum()); n_neg = t.size - n_pos
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    order = np.argsort(s, kind="stable")
    ranks = np.empty(s.size, dtype=np.float64)
    ss = s[order]
    i = 0
    while i < ss.size:
        j = i
        while j + 1 < ss.size and ss[j + 1] == ss[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return (ranks[t].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)


def given_eval(user_emb, item_emb, pairs, topk=-1):
    """test_eval + evaluate, utils/objectives.py:231-294 ('given' scheme): row-wise dot on the listed
    (user, item, truth) pairs, grouped by user; AP via eval_multiple_original when topk == -1, AUC per user;
    unweighted means rounded to 16 dp (:286-294)."""
    pairs = np.asarray(pairs)
    pred = np.sum(user_emb[pairs[:, 0]] * item_emb[pairs[:, 1]], axis=1)
    ap, auc = [], []
    for u in np.unique(pairs[:, 0]):
        sel = pairs[:, 0] == u
        f = eval_multiple_original if topk == -1 else eval_multiple
        a, _, _ = f(pairs[sel, 2], pred[sel], topk)
        ap.append(a); auc.append(auc_score(pairs[sel, 2], pred[sel]))
    return {"map": round(float(np.mean(ap)), 16), "auc": round(float(np.mean(auc)), 16)}


# ---------------------------------------------------------------------------------------------
# CPU baseline (bench.py's cpu_baseline / --impl reference legs): the same neg_shared step as step_matmul,
# written the way the reference's CPU path executes it — gather, one sgemm for S, element-wise loss and
# gradient, two sgemms for dU/dV, sparse update in place — without the dense [table] gradient buffers that
# the checker above returns.  float32 throughout (Keras floatx).  BASELINE.md §2.
# ---------------------------------------------------------------------------------------------
def baseline_neg_shared_sgd_step(EU, EV, uid, cid, neg_loss_weight, lr):
    """One neg_shared skip-gram step with sparse SGD, in place on EU / EV (float32).  Returns the loss.
    ref: models/train_neg_shared.py:46-50 -> models/model_framework.py:59-61,86-87,126 ->
    modules/interaction/interaction_dot.py:100-103 -> utils/objectives.py:99-105."""
    B = uid.shape[0]
    U = EU[uid]
    V = EV[cid]
    S = U @ V.T
    w = np.float32(neg_loss_weight / (B - 1.0))
    sg = 1.0 / (1.0 + np.exp(-S))
    diag = np.arange(B)
    # loss = -sum(W * log(sigmoid(Y * S))) / B
    ls = np.log1p(np.exp(-np.abs(S))) + np.maximum(S, 0)          # softplus(S) = -log sigmoid(-S)
    loss = w * (ls.sum() - ls[diag, diag].sum()) + (ls[diag, diag] - S[diag, diag]).sum()
    G = sg * w
    G[diag, diag] = sg[diag, diag] - 1.0
    G /= np.float32(B)
    dU = G @ V
    dV = G.T @ U
    np.add.at(EU, uid, -lr * dU)
    np.add.at(EV, cid, -lr * dV)
    return float(loss) / B


def baseline_whole_eval_block(U, V, k):
    """Reference-style whole@k on a block of users: blocked U @ V.T (1024 items per block, utils/objectives.py:308-313),
    np.argpartition top-k per user (metrics_ranking.py:7) and the k-element sort.  Returns top-k column ids."""
    pred = np.hstack([U @ V[b:b + 1024].T for b in range(0, V.shape[0], 1024)])
    idx = np.argpartition(-pred, k, axis=1)[:, :k]
    part = np.take_along_axis(pred, idx, axis=1)
    order = np.argsort(-part, axis=1, kind="stable")
    return np.take_along_axis(idx, order, axis=1)
