"""Evaluator: whole@k and given@k evaluation behind the reference's interface.

ref: utils/objectives.py:373-492 (Evaluator), :395-421 (_prepare_for_whole_eval), :296-321 (test_eval_mat),
     :333-370 (evaluate_mat), :231-294 (test_eval / evaluate).
What changed: the dense int32 [users, items] truth matrices become CSR (they would be 80 TB at 10M x 2M), the
[U, I] prediction matrix is never materialised (fused GEMM + top-k kernel), and the per-user python metric loop is
a kernel.  Printed lines keep the reference's format.  Tie rule: score descending, then lowest column index.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from . import ops
from .utilities import get_cur_time


def _csr_truth(links, user_count, item_pos):
    """CSR over users of candidate-column indices (sorted, unique) — the sparse form of the reference's
    `true_mat[user, items_dict[item]] = 1` loops (utils/objectives.py:412-416)."""
    u = links[:, 0].astype(np.int64)
    c = item_pos[links[:, 1]].astype(np.int64)
    key = np.unique(u * (item_pos.max() + 2) + c)
    uu = key // (item_pos.max() + 2)
    cc = key % (item_pos.max() + 2)
    indptr = np.zeros(user_count + 1, dtype=np.int64)
    np.add.at(indptr, uu + 1, 1)
    return np.cumsum(indptr), cc.astype(np.int32)


class Evaluator(object):
    def __init__(self, data_helper, data_spec, conf, eval_scheme=None):
        self.data_helper = data_helper
        self.data_spec = data_spec
        self.conf = conf
        self.eval_scheme = eval_scheme
        if eval_scheme is not None:
            self._check_eval_sheme(eval_scheme)
        self.train = data_helper.data['train']
        self.test_seen = data_helper.data['test_seen']
        self.test = data_helper.data['test']
        self.C = data_helper.data['C']
        self.user_count = data_spec.user_count
        self.item_count = data_spec.item_count
        self.eval_topk = conf.eval_topk
        self.precision = getattr(conf, 'precision', 'bf16')
        self._prepare_for_whole_eval()

    def _check_eval_sheme(self, eval_scheme):
        assert eval_scheme == 'given' or eval_scheme == 'whole', \
            '[Error] Unknown eval_scheme {}.'.format(eval_scheme)

    def _prepare_for_whole_eval(self):
        train, test = self.train, self.test
        # candidate lists: column j <-> j-th smallest item id (sorted(set(...)), objectives.py:400-401)
        self.train_items = np.unique(train[:, 1])
        self.test_items = np.unique(test[:, 1])
        n_items = int(max(self.item_count, train[:, 1].max() + 1, test[:, 1].max() + 1))
        pos_tr = np.full(n_items, -1, dtype=np.int64); pos_tr[self.train_items] = np.arange(self.train_items.size)
        pos_te = np.full(n_items, -1, dtype=np.int64); pos_te[self.test_items] = np.arange(self.test_items.size)
        self.train_truth = _csr_truth(train, self.user_count, pos_tr)
        self.test_truth = _csr_truth(test[test[:, 2] == 1], self.user_count, pos_te)    # label-1 rows only (:414)
        self._dev_truth = None

    def _truth_on_device(self, dev):
        if self._dev_truth is None:
            f = lambda t: (torch.from_numpy(t[0]).to(dev), torch.from_numpy(t[1]).to(dev))
            self._dev_truth = (f(self.train_truth), f(self.test_truth))
        return self._dev_truth

    # ---- whole@k ------------------------------------------------------------------------------------------------
    def _whole_one(self, state, items, truth_dev, topk):
        dev = state.device
        users = torch.arange(self.user_count, device=dev, dtype=torch.int32)
        U = state.user_emb(users)
        V = state.item_emb(torch.from_numpy(items.astype(np.int32)).to(dev))
        ids, scores = ops.eval_topk(U, V, topk, self.precision)
        per_user, sums = ops.eval_metrics(ids, truth_dev[0], truth_dev[1])
        return ids, scores, per_user, sums

    @staticmethod
    def _summ(sums, topk):
        s = sums.cpu().numpy() if isinstance(sums, torch.Tensor) else sums
        n = max(s[3], 1.0)
        return {'map@%d' % topk: s[0] / n, 'recall@%d' % topk: s[1] / n, 'precision@%d' % topk: s[2] / n, 'auc': -1.0}

    def run(self, model, predict_only=False, verbose=True, eval_scheme=None, batch_size=1024, use_async_eval=False):
        """Same arguments as the reference.  `model` is model_dict['model_neg_shared'] (whole) or the model_dict (given).
        use_async_eval: the device work is enqueued, the D2H read of the four metric sums and the print happen in a
        thread (the reference runs its python metric loop in a thread, objectives.py:478-485)."""
        if eval_scheme is None:
            eval_scheme = self.eval_scheme
        self._check_eval_sheme(eval_scheme)
        eval_topk = self.eval_topk
        if eval_scheme == 'given':
            state = model['_state']
            res = [self._given_one(state, t, eval_topk, predict_only) for t in (self.test_seen, self.test)]
            if predict_only:
                return res[0], res[1]
            if verbose:
                print(get_cur_time(), 'train map/auc', res[0]['map@%s' % eval_topk], res[0]['auc'],
                      'test map/auc', res[1]['map@%s' % eval_topk], res[1]['auc'])
            return res[0], res[1]
        assert eval_topk > 0, '[ERROR] eval_top {} must > 0'.format(eval_topk) + 'when eval_scheme=whole'
        state = model.state
        tr_truth, te_truth = self._truth_on_device(state.device)
        tr = self._whole_one(state, self.train_items, tr_truth, eval_topk)
        te = self._whole_one(state, self.test_items, te_truth, eval_topk)
        if predict_only:
            # scalable prediction dump: top-k ids (candidate columns) and scores per user instead of the dense
            # (truth_mat, pred_mat) pair of the reference
            f = lambda r, items: {'items': items, 'topk_cols': r[0].cpu().numpy(), 'topk_scores': r[1].cpu().numpy()}
            return f(tr, self.train_items), f(te, self.test_items)

        def finish():
            a, b = self._summ(tr[3], eval_topk), self._summ(te[3], eval_topk)
            print('train recall/map', a['recall@%s' % eval_topk], a['map@%s' % eval_topk],
                  'test recall/map', b['recall@%s' % eval_topk], b['map@%s' % eval_topk])
            return a, b
        if verbose and use_async_eval:
            ev = torch.cuda.Event()
            ev.record()

            def worker():
                ev.synchronize()
                finish()
            t = threading.Thread(target=worker)
            print(get_cur_time(), end=' ')
            t.start()
            return t, t
        if verbose:
            print(get_cur_time(), end=' ')
            return finish()
        return self._summ(tr[3], eval_topk), self._summ(te[3], eval_topk)

    # ---- given@k ------------------------------------------------------------------------------------------------
    def _given_one(self, state, test, topk, predict_only):
        dev = state.device
        order = np.argsort(test[:, 0], kind='stable')                      # pandas groupby('uid') order
        t = test[order]
        uid = torch.from_numpy(t[:, 0].astype(np.int32)).to(dev)
        cid = torch.from_numpy(t[:, 1].astype(np.int32)).to(dev)
        U = state.user_emb(torch.arange(self.user_count, device=dev, dtype=torch.int32))
        V = state.item_emb(torch.arange(self.item_count, device=dev, dtype=torch.int32))
        scores = ops.score_pairs(U, V, uid, cid)
        if predict_only:
            return list(zip(t[:, 0].tolist(), t[:, 2].tolist(), scores.cpu().numpy().tolist()))
        assert topk == -1 or topk >= 1, '[ERROR] eval_topk {} must be -1 or > 0 when eval_scheme=given'.format(topk)
        users, counts = np.unique(t[:, 0], return_counts=True)
        indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        out = ops.eval_given(scores, torch.from_numpy(t[:, 2].astype(np.int32)).to(dev),
                             torch.from_numpy(indptr).to(dev), topk).cpu().numpy().astype(np.float64)
        # ref utils/objectives.py:286-294: unweighted means over users, rounded to 16 dp
        return {'map@%d' % topk: round(float(np.mean(out[:, 0])), 16),
                'recall@%d' % topk: round(float(np.mean(out[:, 2])), 16),
                'precision@%d' % topk: round(float(np.mean(out[:, 3])), 16),
                'auc': round(float(np.nanmean(out[:, 1])), 16)}
