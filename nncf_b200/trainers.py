"""Trainers: original | neg_shared | group_sample | group_neg_shared (the scoped list) and, widening per SURVEY.md §8f,
presample | reverse | sampled_neg_shared.

ref: models/train_base.py:6-41 (TrainerBase protocol), models/train_neg_shared.py:22-78,
     models/train_group_neg_shared.py:29-76, models/train_original.py:21-88, models/train_group_sample.py:37-108,
     models/train_presample.py:28-114, models/train_reverse.py:36-83, models/train_sampled_neg_shared.py:19-67.
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
from .data_utils import GroupSampler, group_shuffle_train
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

    def _run_epoch(self, model, uid, cid, rows_per_batch, responses=None):
        """uid/cid (and optional y_true `responses`): device int32 arrays holding the epoch's batches back to back.
        Returns (cost_sum, iterations)."""
        state = model.state
        R = state.conf.replicas if state.item_table is not None else 1
        n_batches = uid.numel() // rows_per_batch
        if n_batches == 0:
            return 0.0, 0
        if state.item_table is not None:
            n_steps = n_batches // R                          # a super-step needs R whole batches
            if n_steps == 0:
                return 0.0, 0
            if responses is not None:
                losses = model.train_on_batches(uid, cid, n_steps, responses=responses)
            else:
                losses = model.train_on_batches(uid, cid, n_steps)
            return float(losses.double().sum().item()), n_steps * R
        if responses is None and hasattr(model, 'train_tower_batches'):
            return model.train_tower_batches(uid, cid, rows_per_batch)         # content towers: no host sync per batch
        cost = 0.0
        for b in range(n_batches):
            s = slice(b * rows_per_batch, (b + 1) * rows_per_batch)
            cost += model.train_on_batch([uid[s], cid[s]], [None if responses is None else responses[s]])
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
        self.group_shuffling_trick = group_shuffling_trick
        if group_shuffling_trick:
            _num_in_train = np.max(data_helper.data['train'], axis=0) + 1
            self._iidx = {'user': np.arange(_num_in_train[0]), 'item': np.arange(_num_in_train[1])}
        else:
            # models/train_group_neg_shared.py:33-37
            self.group_sampler = GroupSampler(data_helper.data['train'], group_by='item', chop=conf.chop_size,
                                              seed=getattr(conf, 'seed', 0) + 11)
            self.group_sample = self.group_sampler.sample

    def train(self, eval_scheme=None, use_async_eval=True):
        conf = self.conf
        train = self._device_train()
        n, B = train.shape[0], conf.batch_size_p
        ps = None
        for epoch in range(conf.max_epoch + 1):
            if self.group_shuffling_trick:
                train = group_shuffle_train(train, by='item', chop=conf.chop_size, iidx=self._iidx['item'])
                self._train_dev = train
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = n // B
                if self.group_shuffling_trick:
                    batches = train[:nb * B]
                else:
                    # one GroupSampler.sample(B) per iteration (models/train_group_neg_shared.py:55-56), the whole
                    # epoch drawn in one launch
                    batches = self.group_sampler.sample_device(B, nb).reshape(nb * B, 3)
                cost, it = self._run_epoch(self.model_train, batches[:, 0].contiguous(), batches[:, 1].contiguous(), B)
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
        try:
            self.group_shuffling_trick = bool(conf.group_shuffling_trick)
        except AttributeError:
            self.group_shuffling_trick = False
        if self.group_shuffling_trick:
            assert conf.shuffle_st.startswith('by_item'), 'group_sample supports shuffle_st = by_item* only'
        super().__init__(model_dict, conf, data_helper)
        if self.group_shuffling_trick:
            if conf.neg_dist == 'uniform':
                print('[WARNING] group_shuffling_trick in group_sample does not fully support uniform neg_dist (no_correction).')
            _num_in_train = np.max(data_helper.data['train'], axis=0) + 1
            self._iidx = {'user': np.arange(_num_in_train[0]), 'item': np.arange(_num_in_train[1])}
        else:
            # models/train_group_sample.py:31-36
            self.group_sampler = GroupSampler(data_helper.data['train'], group_by='item', chop=conf.chop_size,
                                              neg_dist=conf.neg_dist, neg_sign=self.neg_sign,
                                              seed=getattr(conf, 'seed', 0) + 12)
            self.group_sample_with_negs = self.group_sampler.sample_with_negs
        print('[INFO] sampling group based on item')

    def _shuffle(self, train):
        if not self.group_shuffling_trick:
            return train
        return group_shuffle_train(train, by='item', chop=self.conf.chop_size, iidx=self._iidx['item'])

    def train(self, eval_scheme=None, use_async_eval=True):
        if self.group_shuffling_trick:
            return super().train(eval_scheme, use_async_eval)
        # GroupSampler.sample_with_negs(B, k) per iteration (models/train_group_sample.py:86-92); the epoch's batches are
        # drawn in one launch and the y_true column travels with them (positives are not always the first B rows)
        conf = self.conf
        k, B = conf.num_negatives, conf.batch_size_p
        n = self._device_train().shape[0]
        ps = None
        for epoch in range(conf.max_epoch + 1):
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = n // B
                rows = (1 + k) * B
                batches, _ = self.group_sampler.sample_with_negs_device(B, k, nb)
                self.group_sampler._dev.check()
                flat = batches.reshape(nb * rows, 3)
                cost, it = self._run_epoch(self.model_train, flat[:, 0].contiguous(), flat[:, 1].contiguous(), rows,
                                           responses=flat[:, 2].contiguous())
                torch.cuda.synchronize()
                self.train_time.append(time.time() - t_start)
            ps = self._epoch_report(epoch, it, cost, eval_scheme, use_async_eval, ps)
        self._finish()


class ReverseTrainer(_PairsTrainer):
    """models/train_reverse.py — IID positives + k sampled negative USERS each ("naively reverse original training");
    pointwise losses only (:15-16).  The per-user weights cu (:19-34) are computed by the reference but never used."""
    neg_col = 0

    def __init__(self, model_dict, conf, data_helper):
        assert conf.loss not in ['log-loss', 'max-margin'], "[ERROR] revrese does not support pairwise losses"
        super().__init__(model_dict, conf, data_helper)

    def _shuffle(self, train):
        perm = np.arange(train.shape[0])
        np.random.shuffle(perm)
        return ops.permute_rows(train, torch.from_numpy(perm).cuda())


class PresampleTrainer(TrainerBase):
    """models/train_presample.py — the whole epoch's negatives are sampled up front, the (1+k)N rows are arranged by
    `shuffle_st` (original | reverse | random | by_user | by_item | by_user_chop* | by_item_chop* | by_useritem_chop*)
    and cut into batches of (1+k)B consecutive rows whose y_true column says which rows are positives."""

    def __init__(self, model_dict, conf, data_helper):
        super().__init__(model_dict, conf, data_helper)
        self.model_train = model_dict['model']
        self.neg_sign = -1 if conf.loss == 'skip-gram' else 0
        self.sample_batch = data_helper.sampler_dict['sample_batch']
        self.sample_batch_u = data_helper.sampler_dict['sample_batch_u']
        _num_in_train = np.max(data_helper.data['train'], axis=0) + 1
        self._iidx = {'user': np.arange(_num_in_train[0]), 'item': np.arange(_num_in_train[1])}
        if conf.shuffle_st == 'reverse' or conf.shuffle_st.startswith('by_item'):
            assert conf.loss not in ['log-loss', 'max-margin'], \
                "[ERROR] shuffle_st %s does not support pairwise losses" % conf.shuffle_st

    def _presample(self, train_p, epoch):
        """-> (train_p after its in-place shuffle, the epoch's (1+k)N rows), both on the device"""
        conf = self.conf
        k, st, chop = conf.num_negatives, conf.shuffle_st, conf.chop_size
        n = train_p.shape[0]

        def shuffled(t):
            perm = np.arange(t.shape[0])
            np.random.shuffle(perm)
            return ops.permute_rows(t, torch.from_numpy(perm).cuda())

        if st == 'original' or st == 'reverse':
            train_p = shuffled(train_p)                                              # np.random.shuffle(train_p)
            sampler = (self.sample_batch if st == 'original' else self.sample_batch_u).__self__
            negs = sampler.sample_batch_device(n * k)
            return train_p, ops.presample_assemble(train_p, k, negs, 1 if st == 'original' else 0, self.neg_sign, 0)
        negs = self.sample_batch.__self__.sample_batch_device(n * k)
        train = ops.presample_assemble(train_p, k, negs, 1, self.neg_sign, 1)         # vstack((train_p, train_n))
        if st == 'random':
            train = shuffled(train)
        elif st == 'by_user':
            train = group_shuffle_train(train, by='user', iidx=self._iidx['user'])
        elif st == 'by_item':
            train = group_shuffle_train(train, by='item', iidx=self._iidx['item'])
        elif st.startswith('by_user_chop'):
            train = group_shuffle_train(train, by='user', chop=chop, iidx=self._iidx['user'])
        elif st.startswith('by_item_chop'):
            train = group_shuffle_train(train, by='item', chop=chop, iidx=self._iidx['item'])
        elif st.startswith('by_useritem_chop'):
            by = 'user' if epoch % 2 == 0 else 'item'
            train = group_shuffle_train(train, by=by, chop=chop, iidx=self._iidx[by])
        else:
            assert False, 'ERROR: unknown shuffle strategy {}'.format(st)
        return train_p, train

    def train(self, eval_scheme=None, use_async_eval=True):
        conf = self.conf
        rows = conf.batch_size_p * (1 + conf.num_negatives)
        train_p = self._device_train()
        ps = None
        for epoch in range(conf.max_epoch + 1):
            train_p, train = self._presample(train_p, epoch)                         # also in epoch 0, like the reference
            self._train_dev = train_p
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = train.shape[0] // rows
                flat = train[:nb * rows]
                cost, it = self._run_epoch(self.model_train, flat[:, 0].contiguous(), flat[:, 1].contiguous(), rows,
                                           responses=flat[:, 2].contiguous())
                torch.cuda.synchronize()
                self.train_time.append(time.time() - t_start)
            del train
            ps = self._epoch_report(epoch, it, cost, eval_scheme, use_async_eval, ps)
        self._finish()


class SampledNegSharedTrainer(TrainerBase):
    """models/train_sampled_neg_shared.py — item-stratified positives (group_shuffle_train by item, chop) plus k
    sampled items per batch shared as negatives by all B positives; the k extra rows carry user id 0."""

    def __init__(self, model_dict, conf, data_helper):
        super().__init__(model_dict, conf, data_helper)
        self.model_train = model_dict['model_sampled_neg_shared']
        self.sample_batch = data_helper.sampler_dict['sample_batch']
        _num_in_train = np.max(data_helper.data['train'], axis=0) + 1
        self._iidx = {'user': np.arange(_num_in_train[0]), 'item': np.arange(_num_in_train[1])}

    def train(self, eval_scheme=None, use_async_eval=True):
        conf = self.conf
        k, B = conf.num_negatives, conf.batch_size_p
        train = self._device_train()
        n = train.shape[0]
        ps = None
        for epoch in range(conf.max_epoch + 1):
            train = group_shuffle_train(train, by='item', chop=conf.chop_size, iidx=self._iidx['item'])
            self._train_dev = train
            cost, it = 0.0, 0
            torch.cuda.synchronize()
            t_start = time.time()
            if epoch > 0:
                nb = n // B
                if nb > 0:
                    negs = self.sample_batch.__self__.sample_batch_device(nb * k)    # sample_batch(k) per batch
                    uid, cid = ops.assemble_sns_batches(train, nb, B, k, negs)
                    cost, it = self._run_epoch(self.model_train, uid, cid, B + k)
                torch.cuda.synchronize()
                self.train_time.append(time.time() - t_start)
            ps = self._epoch_report(epoch, it, cost, eval_scheme, use_async_eval, ps)
        self._finish()


TRAINERS = {'original': OriginalTrainer, 'neg_shared': NegSharedTrainer,
            'group_neg_shared': GroupNegSharedTrainer, 'group_sample': GroupSampleTrainer,
            'presample': PresampleTrainer, 'reverse': ReverseTrainer, 'sampled_neg_shared': SampledNegSharedTrainer}


def get_trainer(train_scheme):
    if train_scheme not in TRAINERS:
        assert False, '[ERROR] Unknown train_scheme {}'.format(train_scheme)
    return TRAINERS[train_scheme]
