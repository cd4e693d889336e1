"""Trainers for the scoped train_scheme list: original | neg_shared | group_sample | group_neg_shared.

ref: models/train_base.py:6-41 (TrainerBase protocol), models/train_neg_shared.py:22-78,
     models/train_group_neg_shared.py:29-76, models/train_original.py:21-88, models/train_group_sample.py:37-108.
Kept: epoch 0 only evaluates (`while epoch > 0`), the tail smaller than batch_size_p is dropped, cost = mean of the
per-batch losses, evaluation after every epoch with the async metric thread joined before the next one, the printed
line formats, NaN abort.  Changed: an epoch's ids live on the device and embedding-table models run the whole epoch
as one C-ABI call (no per-step host round trip); with conf.replicas = R > 1, R consecutive batches form one
synchronous data-parallel super-step (declared in DESIGN.md; R = 1 is the reference's sequential loop).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ops
from .data_utils import group_shuffle_train
from .objectives import Evaluator
from .utilities import get_cur_time, nan_detection, pickle_dump


class TrainerBase(object):
    def __init__(self, model_dict, conf, data_helper):
        self.model_dict = model_dict
        self.conf = conf
        self.data_helper = data_helper
        self.data_spec = data_helper.data_spec
        self.model_predict = model_dict['model_neg_shared']
        self.evaluater = Evaluator(data_helper, self.data_spec, conf)
        self.train_time = []
        print('[INFO] Timestamps below are recorded at the end of training/evaluation respectively')

    def test(self, eval_scheme, predict_only=False, use_async_eval=False):
        if eval_scheme == 'given':
            return self.evaluater.run(self.model_dict, eval_scheme=eval_scheme, predict_only=predict_only,
                                      use_async_eval=use_async_eval)
        elif eval_scheme == 'whole':
            return self.evaluater.run(self.model_predict, eval_scheme=eval_scheme, predict_only=predict_only,
                                      use_async_eval=use_async_eval)

    def predict(self, eval_scheme, pred_saveto=None):
        result = self.test(eval_scheme, predict_only=True)
        if pred_saveto is not None:
            pickle_dump(pred_saveto, result)
        return result

    # ---- shared epoch machinery ---------------------------------------------------------------------------------
    def _device_train(self):
        if getattr(self, '_train_dev', None) is None:
            self._train_dev = torch.from_numpy(np.ascontiguousarray(self.data_helper.data['train'], dtype=np.int32)).cuda()
        return self._train_dev

    def _run_epoch(self, model, uid, cid, rows_per_batch):
        """uid/cid: device int32 arrays holding the epoch's batches back to back.  Returns (cost_sum, iterations)."""
        state = model.state
        R = state.conf.replicas if state.item_table is not None else 1
        n_batches = uid.numel() // rows_per_batch
        if n_batches == 0:
            return 0.0, 0
        if state.item_table is not None:
            n_steps = n_batches // R                          # a super-step needs R whole batches
            if n_steps == 0:
                return 0.0, 0
            losses = model.train_on_batches(uid, cid, n_steps)
            return float(losses.double().sum().item()), n_steps * R
        cost = 0.0
        for b in range(n_batches):
            s = slice(b * rows_per_batch, (b + 1) * rows_per_batch)
            cost += model.train_on_batch([uid[s], cid[s]], None)
        return cost, n_batches

    def _epoch_report(self, epoch, it, cost, eval_scheme, use_async_eval, ps):
        print(get_cur_time(), 'epoch %d (%d it)' % (epoch, it), 'cost %.5f' % (cost / it if it > 0 else -1), end=' ')
        nan_detection('cost', cost)
        if eval_scheme is None:
            print('')
            return ps
        async_eval = True if use_async_eval and epoch != self.conf.max_epoch else False
        try:
            ps[-1].join()
        except Exception:
            pass
        return self.test(eval_scheme, use_async_eval=async_eval)

    def _finish(self):
        torch.cuda.synchronize()
        print('Training time (sec) per epoch:', np.mean(self.train_time) if self.train_time else float('nan'))


class NegSharedTrainer(TrainerBase):
    """models/train_neg_shared.py — shuffle train, slice B positives, one neg_shared step per slice."""

    def __init__(self, model_dict, conf, data_helper):
        super().__init__(model_dict, conf, data_helper)
        self.model_train = model_dict['model_neg_shared']
        if conf.neg_dist != 'unigram':
            print('[WARNING] Only unigram neg_dist is currently supported for group_neg_shared training. '
                  'Set neg_dist = unigram.')

    def train(self, eval_scheme=None, use_async_eval=True):
        conf = self.conf
        train = self._device_train()
        n, B = train.shape[0], conf.batch_size_p
        ps = None
        for epoch in range(conf.max_epoch + 1):
            perm = np.arange(n)
            np.random.shuffle(perm)                                   # np.random.shuffle(train), same stream/draws
            train = ops.permute_rows(train, torch.from_numpy(perm).cuda())
            self._train_dev = train
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = n // B                                            # uneven tail dropped
                cost, it = self._run_epoch(self.model_train, train[:nb * B, 0].contiguous(),
                                           train[:nb * B, 1].contiguous(), B)
                torch.cuda.synchronize()
                self.train_time.append(time.time() - t_start)
            ps = self._epoch_report(epoch, it, cost, eval_scheme, use_async_eval, ps)
        self._finish()


class GroupNegSharedTrainer(TrainerBase):
    """models/train_group_neg_shared.py — item-stratified batches via group_shuffle_train(by='item', chop)."""

    def __init__(self, model_dict, conf, data_helper):
        super().__init__(model_dict, conf, data_helper)
        self.model_train = model_dict['model_group_neg_shared']
        if conf.neg_dist != 'unigram':
            print('[WARNING] Only unigram neg_dist is currently supported for group_neg_shared training. '
                  'Set neg_dist = unigram.')
        try:
            group_shuffling_trick = conf.group_shuffling_trick
        except AttributeError:
            group_shuffling_trick = False
        assert group_shuffling_trick, 'GroupSampler (group_shuffling_trick=False) is outside this round (SURVEY.md §8a5)'
        _num_in_train = np.max(data_helper.data['train'], axis=0) + 1
        self._iidx = {'user': np.arange(_num_in_train[0]), 'item': np.arange(_num_in_train[1])}

    def train(self, eval_scheme=None, use_async_eval=True):
        conf = self.conf
        train = self._device_train()
        n, B = train.shape[0], conf.batch_size_p
        ps = None
        for epoch in range(conf.max_epoch + 1):
            train = group_shuffle_train(train, by='item', chop=conf.chop_size, iidx=self._iidx['item'])
            self._train_dev = train
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = n // B
                cost, it = self._run_epoch(self.model_train, train[:nb * B, 0].contiguous(),
                                           train[:nb * B, 1].contiguous(), B)
                torch.cuda.synchronize()
                self.train_time.append(time.time() - t_start)
            ps = self._epoch_report(epoch, it, cost, eval_scheme, use_async_eval, ps)
        self._finish()


class _PairsTrainer(TrainerBase):
    """Common part of 'original' and 'group_sample': positives + k sampled negatives per positive, 'mul' view."""

    neg_col = 1

    def __init__(self, model_dict, conf, data_helper):
        super().__init__(model_dict, conf, data_helper)
        self.model_train = model_dict['model']
        self.neg_sign = -1 if conf.loss == 'skip-gram' else 0
        key = 'sample_batch' if self.neg_col == 1 else 'sample_batch_u'
        self.sample_batch = data_helper.sampler_dict[key]                 # bound MultinomialSampler.sample_batch
        self.sampler = self.sample_batch.__self__

    def _shuffle(self, train):
        raise NotImplementedError

    def train(self, eval_scheme=None, use_async_eval=True):
        conf = self.conf
        k, B = conf.num_negatives, conf.batch_size_p
        train = self._device_train()
        n = train.shape[0]
        ps = None
        for epoch in range(conf.max_epoch + 1):
            train = self._shuffle(train)
            self._train_dev = train
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = n // B
                # the whole epoch's negatives in one device draw (k*B per batch, consecutive counters = same stream
                # as nb successive sample_batch(k*B) calls), then every (1+k)B batch assembled on the device
                negs = self.sampler.sample_batch_device(nb * k * B)
                rows = (1 + k) * B
                uid = torch.empty(nb * rows, dtype=torch.int32, device='cuda')
                cid = torch.empty(nb * rows, dtype=torch.int32, device='cuda')
                for b in range(nb):
                    batch = ops.assemble_pairs_batch(train[b * B:(b + 1) * B], k, negs[b * k * B:(b + 1) * k * B],
                                                     self.neg_col, self.neg_sign)
                    uid[b * rows:(b + 1) * rows] = batch[:, 0]
                    cid[b * rows:(b + 1) * rows] = batch[:, 1]
                cost, it = self._run_epoch(self.model_train, uid, cid, rows)
                torch.cuda.synchronize()
                self.train_time.append(time.time() - t_start)
            ps = self._epoch_report(epoch, it, cost, eval_scheme, use_async_eval, ps)
        self._finish()


class OriginalTrainer(_PairsTrainer):
    """models/train_original.py — IID positives + k sampled negative ITEMS each."""
    neg_col = 1

    def _shuffle(self, train):
        perm = np.arange(train.shape[0])
        np.random.shuffle(perm)
        return ops.permute_rows(train, torch.from_numpy(perm).cuda())


class GroupSampleTrainer(_PairsTrainer):
    """models/train_group_sample.py — item-grouped positives + k sampled negative USERS each (by == 'item');
    pointwise losses only (:14-15).  The reference's by == 'user' branch writes the wrong variable (:82) and is not
    replicated: shuffle_st must start with 'by_item'."""
    neg_col = 0

    def __init__(self, model_dict, conf, data_helper):
        assert conf.loss not in ['log-loss', 'max-margin'], "[ERROR] group_sample does not support pairwise losses"
        assert conf.shuffle_st.startswith('by_item'), 'group_sample supports shuffle_st = by_item* only'
        try:
            assert conf.group_shuffling_trick
        except (AttributeError, AssertionError):
            assert False, 'GroupSampler (group_shuffling_trick=False) is outside this round (SURVEY.md §8a5)'
        super().__init__(model_dict, conf, data_helper)
        if conf.neg_dist == 'uniform':
            print('[WARNING] group_shuffling_trick in group_sample does not fully support uniform neg_dist (no_correction).')
        _num_in_train = np.max(data_helper.data['train'], axis=0) + 1
        self._iidx = {'user': np.arange(_num_in_train[0]), 'item': np.arange(_num_in_train[1])}
        print('[INFO] sampling group based on item')

    def _shuffle(self, train):
        return group_shuffle_train(train, by='item', chop=self.conf.chop_size, iidx=self._iidx['item'])


TRAINERS = {'original': OriginalTrainer, 'neg_shared': NegSharedTrainer,
            'group_neg_shared': GroupNegSharedTrainer, 'group_sample': GroupSampleTrainer}


def get_trainer(train_scheme):
    if train_scheme not in TRAINERS:
        assert False, '[ERROR] Unknown train_scheme {}'.format(train_scheme)
    return TRAINERS[train_scheme]
