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


def original_loss_grad(s, B, k, loss, neg_loss_weight, gamma, y_true=None):
    """get_original_loss, utils/objectives.py:35-75.  s: [(1+k)B]; first B positives, then k consecutive
    negatives per positive (models/train_original.py:50-53).  y_true is implied by the scheme:
    +1 for positives; -1 (skip-gram) or 0 (mse) for negatives (train_original.py:13-14).  Returns (loss, dL/ds).
    With an explicit y_true the pointwise losses weight each row by its own label exactly as the reference does
    (:59-70), wherever the positives sit in the batch (models/train_presample.py feeds shuffled batches)."""
    s = np.asarray(s).reshape(-1)
    assert s.size == (1 + k) * B
    if y_true is not None and loss in ("skip-gram", "mse"):
        y = np.asarray(y_true, dtype=s.dtype).reshape(-1)
        wv = neg_loss_weight / k
        if loss == "skip-gram":
            w = 1.0 + (1.0 - y) / 2.0 * (wv - 1.0)                   # :60-62
            L = np.sum(-w * log_sigmoid(y * s)) / B                  # :63-64
            g = -w * y * sigmoid(-y * s) / B
        else:
            w = 1.0 + (1.0 - y) * (wv - 1.0)                         # :66-68
            L = np.sum(w * (y - s) ** 2) / B                         # :69-70
            g = -2.0 * w * (y - s) / B
        return L, g
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
                norm_u=False, norm_v=False, ubias=None, cbias=None):
    """One neg_shared / group_neg_shared batch on plain embedding tables.
    Graph: models/model_framework.py:40-65 (user side), :85-88 / :98-111 (item side), :126-136 (scores),
    modules/interaction/interaction_dot.py:100-107 ('matmul'), loss per scheme, activity regulariser
    utils/utilities.py:122-135 on the un-normalised user rows.
    ubias [U] / cbias [I] (optional): InteractionDot's bias terms, S += ubias[uid][:, None] + cbias[col_ids][None, :]
    (modules/interaction/interaction_dot.py:104-107); their gradients come back as dubias / dcbias.
    Returns dict(loss, task_loss, dEU [U,d], dEV [I,d], S)."""
    uid = np.asarray(uid); cid = np.asarray(cid)
    B = uid.shape[0]
    U_raw = EU[uid]
    U, inv_u = l2_normalize(U_raw) if norm_u else (U_raw, None)
    if scheme == "neg_shared":
        V_raw = EV[cid]
        V, inv_v = l2_normalize(V_raw) if norm_v else (V_raw, None)
        S = U @ V.T
        if ubias is not None:
            S = S + np.asarray(ubias)[uid][:, None]
        if cbias is not None:
            S = S + np.asarray(cbias)[cid][None, :]
        L, G = neg_shared_loss_grad(S, loss, neg_loss_weight, gamma)
        col_ids = cid
    elif scheme == "group_neg_shared":
        cid_u, cid_x = unique_first_occurrence(cid)
        V_raw = EV[cid_u]
        V, inv_v = l2_normalize(V_raw) if norm_v else (V_raw, None)
        S = U @ V.T
        if ubias is not None:
            S = S + np.asarray(ubias)[uid][:, None]
        if cbias is not None:
            S = S + np.asarray(cbias)[cid_u][None, :]
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
    out = dict(loss=L + reg, task_loss=L, S=S,
               dEU=_scatter_rows(EU.shape[0], uid, dU), dEV=_scatter_rows(EV.shape[0], col_ids, dV))
    if ubias is not None:
        out["dubias"] = _scatter_rows(EU.shape[0], uid, G.sum(axis=1, keepdims=True))[:, 0]
    if cbias is not None:
        out["dcbias"] = _scatter_rows(EV.shape[0], col_ids, G.sum(axis=0)[:, None])[:, 0]
    return out


def step_mul(EU, EV, uid, cid, B, k, loss, neg_loss_weight, gamma, u_reg=0.0, norm_u=False, norm_v=False, y_true=None,
             ubias=None, cbias=None):
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
    if ubias is not None:                                             # interaction_dot.py:96-99
        s = s + np.asarray(ubias)[uid]
    if cbias is not None:
        s = s + np.asarray(cbias)[cid]
    L, g = original_loss_grad(s, B, k, loss, neg_loss_weight, gamma, y_true)
    dU = g[:, None] * V
    dV = g[:, None] * U
    if norm_u:
        dU = l2_normalize_bwd(U, inv_u, dU)
    if norm_v:
        dV = l2_normalize_bwd(V, inv_v, dV)
    reg = u_reg * np.sum(np.mean(U_raw ** 2, axis=0))
    dU = dU + 2.0 * u_reg * U_raw / n
    out = dict(loss=L + reg, task_loss=L, s=s,
               dEU=_scatter_rows(EU.shape[0], uid, dU), dEV=_scatter_rows(EV.shape[0], cid, dV))
    if ubias is not None:
        out["dubias"] = _scatter_rows(EU.shape[0], uid, g[:, None])[:, 0]
    if cbias is not None:
        out["dcbias"] = _scatter_rows(EV.shape[0], cid, g[:, None])[:, 0]
    return out


def sampled_neg_shared_loss_grad(pred, loss, neg_loss_weight, gamma):
    """get_sampled_neg_shared_loss, utils/objectives.py:120-161.  pred [B, 1+k]: column 0 = the positive's score,
    columns 1.. = scores against the k shared sampled negatives.  Returns (loss, dL/dpred)."""
    pred = np.asarray(pred)
    B, k = pred.shape[0], pred.shape[1] - 1
    G = np.zeros_like(pred)
    if loss in ("max-margin", "log-loss"):
        D = pred[:, :1] - pred[:, 1:]                                # :131, :135
        if loss == "max-margin":
            L = np.mean(np.maximum(gamma - D, 0.0))                  # :132
            A = -((gamma - D) > 0).astype(pred.dtype) / (B * k)
        else:
            L = np.mean(-log_sigmoid(gamma * D))                     # :136
            A = -gamma * sigmoid(-gamma * D) / (B * k)
        G[:, 1:] = -A
        G[:, 0] = A.sum(axis=1)
    elif loss == "skip-gram":
        w = np.ones_like(pred); w[:, 1:] = neg_loss_weight / k       # :138-141
        y = np.ones_like(pred); y[:, 1:] = -1.0                      # :142-144
        L = np.sum(-w * log_sigmoid(y * pred)) / B                   # :145-146
        G = -w * y * sigmoid(-y * pred) / B
    elif loss == "mse":
        w = np.ones_like(pred); w[:, 1:] = neg_loss_weight / k       # :148-151
        y = np.ones_like(pred); y[:, 1:] = 0.0                       # :152-154
        L = np.sum(w * (pred - y) ** 2) / B                          # :155-156
        G = 2.0 * w * (pred - y) / B
    else:
        raise AssertionError("[ERROR!] loss %s not specified." % loss)
    return L, G


def step_sampled_neg_shared(EU, EV, uid, cid, B, k, loss, neg_loss_weight, gamma, u_reg=0.0, norm_u=False, norm_v=False):
    """One 'sampled_neg_shared' batch of B + k rows: rows [0, B) are positive (user, item) links, rows [B, B+k) carry
    the k sampled negative items (and a dummy user id, 0 in models/train_sampled_neg_shared.py:28).
    Graph: models/model_framework.py:113-118 (front / back split), :138-143 (pred = [mul(front), matmul(U_front,
    C_back^T)]); the activity regulariser sees the Embedding output of ALL B + k user ids (utils/utilities.py:129-135)."""
    uid = np.asarray(uid); cid = np.asarray(cid)
    n = uid.shape[0]
    assert n == B + k
    U_raw, V_raw = EU[uid], EV[cid]
    U, inv_u = l2_normalize(U_raw) if norm_u else (U_raw, None)
    V, inv_v = l2_normalize(V_raw) if norm_v else (V_raw, None)
    Uf, Vf, Vb = U[:B], V[:B], V[B:]
    pred = np.concatenate([np.sum(Uf * Vf, axis=1, keepdims=True), Uf @ Vb.T], axis=1)
    L, G = sampled_neg_shared_loss_grad(pred, loss, neg_loss_weight, gamma)
    dU = np.zeros_like(U); dV = np.zeros_like(V)
    dU[:B] = G[:, :1] * Vf + G[:, 1:] @ Vb
    dV[:B] = G[:, :1] * Uf
    dV[B:] = G[:, 1:].T @ Uf
    if norm_u:
        dU = l2_normalize_bwd(U, inv_u, dU)
    if norm_v:
        dV = l2_normalize_bwd(V, inv_v, dV)
    reg = u_reg * np.sum(np.mean(U_raw ** 2, axis=0))
    dU = dU + 2.0 * u_reg * U_raw / n
    return dict(loss=L + reg, task_loss=L, pred=pred, dU_rows=dU, dV_rows=dV,
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
def presample_rows(train_p, k, negs, neg_col, neg_sign, layout):
    """The (1+k)N rows models/train_presample.py builds each epoch from the (already shuffled) positives and N k
    sampled ids `negs` (negs[p k + j] = j-th negative of positive p).
    layout 0 (shuffle_st 'original' / 'reverse', :46-60): train_p.repeat(1+k); column neg_col <- samples; column 2 <-
    neg_sign; every (1+k)-th row restored to the positive.  layout 1 (:61-65): vstack((train_p, train_n))."""
    train_p = np.asarray(train_p)
    n = train_p.shape[0]
    negs = np.asarray(negs).reshape(n, k)
    if layout == 0:
        out = train_p.repeat(1 + k, axis=0)
        blk = out.reshape(n, 1 + k, 3)
        blk[:, 1:, neg_col] = negs
        blk[:, 1:, 2] = neg_sign
        return blk.reshape(-1, 3)
    train_n = train_p.repeat(k, axis=0)
    train_n[:, neg_col] = negs.reshape(-1)
    train_n[:, 2] = neg_sign
    return np.vstack((train_p, train_n))


def assemble_sns_batch(train_batch_p, k, neg_items):
    """models/train_sampled_neg_shared.py:28,46-49: B positives + k rows (0, sampled item, 0)."""
    train_batch_n = np.zeros((k, 3), dtype=train_batch_p.dtype)
    train_batch_n[:, 1] = neg_items
    return np.vstack((train_batch_p, train_batch_n))


class GroupSamplerOracle(object):
    """Restatement of class GroupSampler, configs/data_utils.py:244-408, with every random draw taken from one explicit
    np.random.RandomState (the reference mixes np.random, `random` and its time-seeded C++ sampler).  The two inner
    samplers are the analytic distributions of get_sampler (:193-215): p ∝ degree^power over ids with degree > 0."""

    def __init__(self, train, group_by="item", chop=1, neg_dist="unigram", neg_sign=0, neg_sampling_power=0.75, rng=None):
        self.rng = np.random.RandomState(0) if rng is None else rng
        no_correction = neg_dist == "uniform_no_correction"           # :253-257
        if no_correction:
            neg_dist = "uniform"
        self.chop, self.neg_sign, self.group_by = chop, neg_sign, group_by
        gcol, mcol = (1, 0) if group_by == "item" else (0, 1)        # :267-272
        train = np.asarray(train)
        gdeg = np.bincount(train[:, gcol]).astype(np.float64)
        mdeg = np.bincount(train[:, mcol]).astype(np.float64)
        self.p_group = gdeg / gdeg.sum()                              # :281-282 unigram, power 1
        mw = np.where(mdeg > 0, 1.0 if neg_dist == "uniform" else mdeg ** neg_sampling_power, 0.0)
        self.p_member = mw / mw.sum()                                 # :283-285
        gset = np.nonzero(gdeg)[0]
        self.p_n_div_p_d = np.ones(gdeg.size)                         # :287-299
        if neg_dist == "uniform" and not no_correction:
            self.p_n_div_p_d[gset] = 1.0 / gset.size / self.p_group[gset]
        order = np.argsort(train[:, gcol], kind="stable")            # :301-304 (members in train order)
        self.members = train[order, mcol]
        self.indptr = np.concatenate([[0], np.cumsum(np.bincount(train[:, gcol]))])

    def _members(self, group, n):
        lo, hi = self.indptr[group], self.indptr[group + 1]
        return self.members[lo + self.rng.randint(0, hi - lo, size=n)]   # np.random.choice(gms, chop), :322

    def sample(self, batch_size_p, strict_return_shape=True):
        n_groups = int(np.ceil(float(batch_size_p) / self.chop)) if strict_return_shape else batch_size_p // self.chop
        groups = self.rng.choice(self.p_group.size, size=n_groups, p=self.p_group)          # :318
        mem = np.concatenate([self._members(g, self.chop) for g in groups])
        out = np.stack([mem, np.repeat(groups, self.chop), np.ones(mem.size, dtype=np.int64)], 1)   # :324-329
        if self.group_by == "user":
            out = out[:, [1, 0, 2]]
        return out[:batch_size_p] if strict_return_shape else out

    def sample_with_negs(self, batch_size_p, k):
        whole = batch_size_p * (1 + k)
        gpos, mpos, gneg, mneg = [], [], [], []

        def add(groups):                                              # :352-364
            for g in groups:
                mpos.append(self._members(g, self.chop)); gpos.extend([g] * self.chop)
                x = k * self.chop * self.p_n_div_p_d[g]
                nn = int(x) + int(self.rng.random_sample() < x - int(x))   # random_rounding
                if nn > 0:
                    gneg.extend([g] * nn)
                    mneg.append(self.rng.choice(self.p_member.size, size=nn, p=self.p_member))

        add(self.rng.choice(self.p_group.size, size=int(np.ceil(float(batch_size_p) / self.chop)), p=self.p_group))
        for _ in range(10):                                           # :366-381
            diff = whole - (len(gpos) + len(gneg))
            if diff <= 0:
                break
            add(self.rng.choice(self.p_group.size, size=diff if diff <= 100 else diff // (self.chop * (1 + k)) + 100,
                                p=self.p_group))
        pos = np.stack([np.concatenate(mpos), np.array(gpos), np.ones(len(gpos), dtype=np.int64)], 1)
        neg = np.stack([np.concatenate(mneg), np.array(gneg), np.full(len(gneg), self.neg_sign, dtype=np.int64)], 1)
        out = np.vstack((pos, neg))                                   # :383-399
        if self.group_by == "user":
            out = out[:, [1, 0, 2]]
        return out[:whole]


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
    ap, rc, pr, auc = [], [], [], []
    for u in np.unique(pairs[:, 0]):
        sel = pairs[:, 0] == u
        f = eval_multiple_original if topk == -1 else eval_multiple
        a, r, p = f(pairs[sel, 2], pred[sel], topk)
        ap.append(a); rc.append(r); pr.append(p); auc.append(auc_score(pairs[sel, 2], pred[sel]))
    return {"map": round(float(np.mean(ap)), 16), "recall": round(float(np.mean(rc)), 16),
            "precision": round(float(np.mean(pr)), 16), "auc": round(float(np.mean(auc)), 16)}


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


def baseline_neg_shared_step(EU, EV, uid, cid, loss, neg_loss_weight, gamma, lr, u_reg=0.0, norm=False, replicas=1):
    """The reference's CPU step for any of the four neg_shared losses (float32, sparse SGD in place), `replicas` batches
    of B = len(uid) / replicas links computed against ONE snapshot of the tables with their updates summed - what bench.py's
    GPU arm calls a step (replicas = 1 is the reference's sequential loop).  Per batch: gather, one sgemm for S
    (interaction_dot.py:100-103), loss + dL/dS element-wise (utils/objectives.py:78-117), two sgemms for dU / dV,
    l2-normalise forward / backward when `norm` (model_framework.py:62-63,109-111), activity regulariser on the raw user
    rows (utils/utilities.py:129-135).  Returns the mean loss of the batches."""
    f32 = np.float32
    n = uid.shape[0]
    B = n // replicas
    upd = []
    total = 0.0
    for r in range(replicas):
        u, c = uid[r * B:(r + 1) * B], cid[r * B:(r + 1) * B]
        U_raw = EU[u]; V_raw = EV[c]
        if norm:
            iu = (1.0 / np.sqrt(np.maximum(np.sum(U_raw * U_raw, axis=1, keepdims=True), 1e-12))).astype(f32)
            iv = (1.0 / np.sqrt(np.maximum(np.sum(V_raw * V_raw, axis=1, keepdims=True), 1e-12))).astype(f32)
            U, V = U_raw * iu, V_raw * iv
        else:
            U, V = U_raw, V_raw
        S = U @ V.T
        diag = np.arange(B)
        if loss == "skip-gram":
            w = f32(neg_loss_weight / (B - 1.0))
            sg = 1.0 / (1.0 + np.exp(-S))
            ls = np.log1p(np.exp(-np.abs(S))) + np.maximum(S, 0)      # softplus(S) = -log sigmoid(-S)
            L = (w * (ls.sum() - ls[diag, diag].sum()) + (ls[diag, diag] - S[diag, diag]).sum()) / B
            G = sg * w
            G[diag, diag] = sg[diag, diag] - 1.0
            G /= f32(B)
        elif loss == "mse":
            w = f32(neg_loss_weight / (B - 1.0))
            T = S.copy(); T[diag, diag] -= 1.0
            W = np.full_like(S, w); W[diag, diag] = 1.0
            L = float(np.sum(W * T * T)) / B
            G = (2.0 * W * T / f32(B)).astype(f32)
        else:
            D = S[diag, diag][None, :] - S                            # (B,) - (B,B): D[i,j] = S[j,j] - S[i,j]  (:92,:97)
            if loss == "log-loss":
                x = -f32(gamma) * D
                L = float(np.mean(np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0)))
                A = (-f32(gamma) / (1.0 + np.exp(-x)) / f32(B * B)).astype(f32)
            else:                                                     # max-margin (:91-94)
                M = np.full_like(S, f32(gamma)); M[diag, diag] = 0.0
                T = M - D
                L = float(np.mean(np.maximum(T, 0.0)))
                A = (-(T > 0).astype(f32) / f32(B * B))
            G = -A
            G[diag, diag] += A.sum(axis=0)
        dU = G @ V
        dV = G.T @ U
        if norm:
            dU = (dU - U * np.sum(dU * U, axis=1, keepdims=True)) * iu
            dV = (dV - V * np.sum(dV * V, axis=1, keepdims=True)) * iv
        if u_reg:
            L += float(u_reg * np.sum(np.mean(U_raw * U_raw, axis=0)))
            dU = dU + f32(2.0 * u_reg / B) * U_raw
        total += float(L)
        upd.append((u, c, dU, dV))
    for u, c, dU, dV in upd:                                          # every batch saw the same snapshot
        np.add.at(EU, u, (-lr * dU).astype(f32))
        np.add.at(EV, c, (-lr * dV).astype(f32))
    return total / replicas


def baseline_whole_eval_block(U, V, k):
    """Reference-style whole@k on a block of users: blocked U @ V.T (1024 items per block, utils/objectives.py:308-313),
    np.argpartition top-k per user (metrics_ranking.py:7) and the k-element sort.  Returns top-k column ids."""
    pred = np.hstack([U @ V[b:b + 1024].T for b in range(0, V.shape[0], 1024)])
    idx = np.argpartition(-pred, k, axis=1)[:, :k]
    part = np.take_along_axis(pred, idx, axis=1)
    order = np.argsort(-part, axis=1, kind="stable")
    return np.take_along_axis(idx, order, axis=1)
