"""Data layer: the reference's array contract, the sampler factory and the batch-index builders.

ref: configs/data_utils.py:15-70 (get_data, DataSpec, the four samplers), :193-215 (get_sampler),
     :218-241 (group_shuffle_train); data/readme.txt:1-9 (pickle keys and array layouts).
Array contract kept: train/test/test_seen = int [N,3] rows (user, item, label); C = int [items, L] left-zero-padded.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from .sampler import MultinomialSampler

data_root = './data'


class DataHelper(object):
    """Plain attribute holder, as used on the reference's run path (configs/data_utils.py:74,103)."""

    def __init__(self):
        self.data = None
        self.data_spec = None
        self.sampler_dict = None


class DataSpec(object):
    def __init__(self, user_count, word_count, item_count, max_content_len):
        self.user_count = user_count
        self.word_count = word_count
        self.item_count = item_count
        self.max_content_len = max_content_len
        self.W_pretrain = None
        self.C_pretrain = None


# ------------------------------------------------------------------------------------------------
# synthetic power-law interaction data of the named shapes
# ------------------------------------------------------------------------------------------------
def powerlaw_ids(n_ids, n_draws, exponent, perm_seed, draw_rng, offset=10.0):
    """ids ~ p(rank) ∝ (rank + offset)^-exponent over a seeded random permutation of the id space."""
    p = np.power(np.arange(n_ids, dtype=np.float64) + offset, -exponent)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, draw_rng.random_sample(n_draws), side='right')
    ranks = np.minimum(ranks, n_ids - 1)
    perm = np.random.RandomState(perm_seed).permutation(n_ids)
    return perm[ranks]


def make_synthetic(user_count=5551, item_count=16980, n_links=204986, content_len=300, vocab=8000,
                   cold_fraction=0.2, seed=2017, n_test_negatives=4):
    """CiteULike-shaped stand-in (the data blob is absent from the reference tree, .MISSING_LARGE_BLOBS):
    power-law links, 20% of the items held out as cold-start test items, 0-left-padded content matrix."""
    rng = np.random.RandomState(seed)
    users = powerlaw_ids(user_count, n_links, 0.8, 124, rng)
    items = powerlaw_ids(item_count, n_links, 1.0, 123, rng)
    links = np.stack([users, items, np.ones(n_links, dtype=np.int64)], 1)
    cold = np.random.RandomState(5).uniform(size=item_count) < cold_fraction
    is_test = cold[links[:, 1]]
    train = links[~is_test]
    test_pos = links[is_test]
    test_items = np.nonzero(cold)[0]
    neg_u = np.repeat(test_pos[:, 0], n_test_negatives)
    neg_i = test_items[rng.randint(0, test_items.size, size=neg_u.size)]
    test = np.vstack([test_pos, np.stack([neg_u, neg_i, np.zeros(neg_u.size, dtype=np.int64)], 1)])
    sub = rng.choice(train.shape[0], size=min(train.shape[0], 20000), replace=False)
    seen_pos = train[sub]
    train_items = np.unique(train[:, 1])
    neg_u = np.repeat(seen_pos[:, 0], n_test_negatives)
    neg_i = train_items[rng.randint(0, train_items.size, size=neg_u.size)]
    test_seen = np.vstack([seen_pos, np.stack([neg_u, neg_i, np.zeros(neg_u.size, dtype=np.int64)], 1)])
    # content: random length in [5, L], words ~ Zipf over the vocabulary, zero padding in the beginning
    C = np.zeros((item_count, content_len), dtype=np.int32)
    lens = rng.randint(5, content_len + 1, size=item_count)
    words = 1 + powerlaw_ids(vocab - 1, int(lens.sum()), 1.0, 77, rng, offset=2.0)
    pos = 0
    for i in range(item_count):
        C[i, content_len - lens[i]:] = words[pos:pos + lens[i]]
        pos += lens[i]
    return {'C': C, 'train': train.astype(np.int32), 'test': test.astype(np.int32),
            'test_seen': test_seen.astype(np.int32), 'train_items': train_items.tolist(), 'test_items': test_items.tolist()}


def create_val(data, val_ratio=0.1, rng=None):
    """Validation split of a `data_split_cold_item` dictionary, in place: a random `val_ratio` of the TRAIN items becomes
    the new cold-start test set (their links leave `train`), `test_seen` = the first len(test) remaining train rows.
    ref: data/create_val.py:8,16-40 (same steps, same np.random.shuffle of train_items, same sanity checks)."""
    rng = np.random if rng is None else rng
    train_items = list(data['train_items'])
    train = data['train']
    num_train_items, num_train = len(train_items), train.shape[0]
    rng.shuffle(train_items)
    test_len = int(len(train_items) * val_ratio)
    test_items = train_items[:test_len]
    train_items = train_items[test_len:]
    tt_idx = np.isin(train[:, 1], np.asarray(test_items, dtype=train.dtype))
    test = train[tt_idx]
    train = train[~tt_idx]
    data['train_items'], data['test_items'] = train_items, test_items
    data['train'], data['test'], data['test_seen'] = train, test, train[:test.shape[0]]
    assert len(train_items) + len(test_items) == num_train_items
    assert data['train'].shape[0] + data['test'].shape[0] == num_train
    assert len(set(test_items).intersection(train_items)) == 0
    return data


def _get_data(data_name):
    """Loads the reference's `data_split_cold_item.pkl` (a Python-2 pickle of NumPy arrays, keys per
    data/readme.txt:3-9) when present under ./data, else builds the synthetic stand-in of the same shape."""
    data_helper = DataHelper()
    import re
    sub_folder = ''
    fold = re.findall(r'fold(\d+)', data_name)
    if len(fold) == 1:
        sub_folder = 'fold%d' % int(fold[0])
    split_file = None
    for family, prefix in (('citeulike', 'citeulike_'), ('news', 'news_')):
        for kind in ('title_only', 'title_and_abstract'):
            if data_name.startswith(prefix + kind):
                split_file = '%s/%s/%s/%s/data_split_cold_item.pkl' % (data_root, family, kind, sub_folder)
    if split_file is not None and os.path.exists(split_file):
        with open(split_file, 'rb') as fp:
            data_helper.data = pickle.load(fp, encoding='latin1')
    elif data_name.startswith('synthetic_small'):
        data_helper.data = make_synthetic(600, 1500, 20000, content_len=40, vocab=500, seed=11)
        if data_name.endswith('_val'):
            create_val(data_helper.data, rng=np.random.RandomState(3))
    elif split_file is not None or data_name.startswith('synthetic'):
        print('[INFO] %s: data blob not found, using the synthetic CiteULike-shaped stand-in' % data_name)
        data_helper.data = make_synthetic()
    else:
        assert False, '[ERROR] unseen data_name %s' % data_name
    return data_helper


def get_pretrain_folder(data_name, aug=True):
    """configs/data_utils.py:107-126: <data_root>/<family>/<kind>/pretrain/<aug|''>/ ; None for a data_name outside the
    four families (the synthetic stand-ins), where the reference asserts."""
    for family, prefix in (('citeulike', 'citeulike_'), ('news', 'news_')):
        for kind in ('title_only', 'title_and_abstract'):
            if data_name.startswith(prefix + kind):
                return data_root + '/%s/%s/pretrain/%s/' % (family, kind, 'aug' if aug else '')
    return None


def get_pretrained_vectors(conf, data_spec, data_helper):
    """configs/data_utils.py:129-185: (W_pretrain [word_count, dw] | None, C_pretrain [item_count, dc] | None) from
    `conf.pretrain`.  A '.pkl' file holds the matrix itself (Python-2 pickle: latin1); the text formats are word2vec's
    (first line skipped, `word v0 v1 ... ` with a trailing blank column; words resolved through `data_helper.word2id`)
    and doc2vec's (`_*<item id> v0 v1 ... `; ids >= item_count dropped).  Engine addition: the arrays may be handed
    over directly as conf.pretrain['W_pretrain'] / ['C_pretrain']."""
    pre = getattr(conf, 'pretrain', None) or {}
    W_pretrain = C_pretrain = None
    wordvec_filepath, sentvec_filepath = pre.get('wordvec_filepath'), pre.get('sentvec_filepath')
    if pre.get('W_pretrain') is not None:
        W_pretrain = np.asarray(pre['W_pretrain'])
    elif wordvec_filepath and os.path.exists(wordvec_filepath):
        if wordvec_filepath.endswith('pkl'):
            with open(wordvec_filepath, 'rb') as fp:
                W_pretrain = np.asarray(pickle.load(fp, encoding='latin1'))
        else:
            word2id = getattr(data_helper, 'word2id', None) or {}
            with open(wordvec_filepath) as fp:
                fp.readline()
                for line in fp:
                    tok = line.rstrip().split(' ')
                    if W_pretrain is None:
                        W_pretrain = np.zeros((data_spec.word_count, len(tok) - 1))
                    wid = word2id.get(tok[0])
                    if wid is not None and wid < data_spec.word_count:
                        W_pretrain[wid] = [float(x) for x in tok[1:]]
    if pre.get('C_pretrain') is not None:
        C_pretrain = np.asarray(pre['C_pretrain'])
    elif sentvec_filepath and os.path.exists(sentvec_filepath):
        if sentvec_filepath.endswith('pkl'):
            with open(sentvec_filepath, 'rb') as fp:
                C_pretrain = np.asarray(pickle.load(fp, encoding='latin1'))
        else:
            with open(sentvec_filepath) as fp:
                for line in fp:
                    tok = line.rstrip().split(' ')
                    if C_pretrain is None:
                        C_pretrain = np.zeros((data_spec.item_count, len(tok) - 1))
                    iid = int(tok[0][2:])                       # '_*<id>'
                    if iid < data_spec.item_count:
                        C_pretrain[iid] = [float(x) for x in tok[1:]]
    if W_pretrain is not None:
        assert W_pretrain.shape[0] == data_spec.word_count, 'W_pretrain rows != word_count'
    if C_pretrain is not None:
        assert C_pretrain.shape[0] == data_spec.item_count, 'C_pretrain rows != item_count'
    return W_pretrain, C_pretrain


def get_data(data_name, conf, reverse_samping=False):
    """configs/data_utils.py:15-70 (the misspelt keyword `reverse_samping` is the reference's, main.py:67)."""
    data_helper = _get_data(data_name)
    train = data_helper.data['train']
    test = data_helper.data['test']
    C = data_helper.data['C']
    user_count = int(max(np.max(train[:, 0]), np.max(test[:, 0])) + 1)
    word_count = int(np.max(C) + 1)
    item_count = C.shape[0]
    max_content_len = C.shape[1]
    data_spec = DataSpec(user_count, word_count, item_count, max_content_len)
    if conf is not None:
        data_spec.W_pretrain, data_spec.C_pretrain = get_pretrained_vectors(conf, data_spec, data_helper)   # :37
        neg_dist = conf.neg_dist
        try:
            neg_sampling_power = conf.neg_sampling_power
        except AttributeError:
            neg_sampling_power = 0.75
        seed = getattr(conf, 'seed', 0)
        sampler_dict = {}
        if reverse_samping:
            train_r = train[:, [1, 0, 2]]
            sampler_dict['sample_u'] = get_sampler(train_r, neg_dist, neg_sampling_power, rand_seed=seed + 1,
                                                   batch_mode=False)
            sampler_dict['sample_batch_u'] = get_sampler(train_r, neg_dist, neg_sampling_power, rand_seed=seed + 2,
                                                         batch_mode=True)
        sampler_dict['sample'] = get_sampler(train, neg_dist, neg_sampling_power, rand_seed=seed + 3, batch_mode=False)
        sampler_dict['sample_batch'] = get_sampler(train, neg_dist, neg_sampling_power, rand_seed=seed + 4,
                                                   batch_mode=True)
        data_helper.sampler_dict = sampler_dict
    data_helper.data_spec = data_spec
    return data_helper


def get_sampler(ratings, neg_dist='unigram', neg_sampling_power=0.75, column=1, rand_seed=0, batch_mode=True):
    """configs/data_utils.py:193-215: degree histogram of `column` -> MultinomialSampler; returns the bound
    sample_batch / sample method.  'uniform' => dist[dist > 0] = 1; a suffix after '_' in neg_dist is ignored."""
    neg_dist = neg_dist.split('_')[0]
    assert neg_dist == 'uniform' or neg_dist == 'unigram', [neg_dist]
    dist = np.bincount(ratings[:, column], minlength=int(np.max(ratings[:, column])) + 1).astype(float)
    if neg_dist == 'uniform':
        dist[dist > 0] = 1
    s = MultinomialSampler(dist, dist.size, neg_sampling_power, rand_seed)
    return s.sample_batch if batch_mode else s.sample


def group_shuffle_train(train, by='item', chop=0, iidx=None, rng=None):
    """group_shuffle_train (configs/data_utils.py:218-241) with the work on the device: the host draws the three
    permutations from the shared legacy stream in the reference's order (np.random by default, as the reference),
    the device applies them with a stable radix sort.  `train` may be a NumPy array (returned as NumPy) or a CUDA
    int32 tensor (returned as a CUDA tensor).  iidx is shuffled IN PLACE and persists across epochs."""
    import torch
    from . import ops
    rng = np.random if rng is None else rng
    col = 0 if by == 'user' else 1
    is_np = isinstance(train, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(train, dtype=np.int32)).cuda() if is_np else train
    n = t.shape[0]
    if iidx is None:
        iidx = np.arange(int(t[:, col].max().item()) + 1)
    rng.shuffle(iidx)
    row_perm = np.arange(n)
    rng.shuffle(row_perm)
    block_perm = None
    if chop > 0:
        block_perm = np.arange(n // chop)
        rng.shuffle(block_perm)
    dev = t.device
    out = ops.group_shuffle(t, col, torch.from_numpy(iidx).to(dev), torch.from_numpy(row_perm).to(dev),
                            None if block_perm is None else torch.from_numpy(block_perm).to(dev), chop)
    if is_np:
        # the reference also shuffles its input rows in place (np.random.shuffle(train))
        train[:] = train[row_perm]
        return out.cpu().numpy().astype(train.dtype)
    return out


class GroupSampler(object):
    """First sample a group/item, then sample its positive members/users, followed by sampling negative
    members/users (whose number is decided).

    ref: configs/data_utils.py:244-408.  Same constructor keywords, `sample(batch_size_p, strict_return_shape)` and
    `sample_with_negs(batch_size_p, k, strict_return_shape)` returning host int arrays [rows, 3] like the reference;
    the `*_device` variants return CUDA tensors holding many batches from one launch (what the trainers use).
    Kept: group ~ degree^1, `chop` members with replacement, random_rounding(k * chop * p_n/p_d) negatives per group,
    the top-up rounds, positives-then-negatives order, truncation, column swap for group_by='user'.
    Declared: the reference reads an undefined global `conf` for neg_sampling_power and therefore ALWAYS uses 0.75
    (:274-280); here 0.75 is the default and the power is a keyword.  Randomness comes from the device Philox stream
    (the reference uses np.random + a time-seeded LCG), so parity is distributional.
    """

    def __init__(self, train, group_by='item', chop=1, neg_dist='unigram', neg_sign=0, neg_sampling_power=0.75, seed=0):
        from . import ops
        self.train = train
        self.group_by = group_by
        self.chop = chop
        self.neg_dist = neg_dist
        self.neg_sign = int(np.asarray(neg_sign).reshape(-1)[0])
        self._dev = ops.DeviceGroupSampler(train, group_by, chop, neg_dist, self.neg_sign, neg_sampling_power, seed)

    # ---- device variants (many batches per launch)
    def sample_device(self, batch_size_p, n_batches=1):
        return self._dev.sample(batch_size_p, n_batches)

    def sample_with_negs_device(self, batch_size_p, k, n_batches=1):
        return self._dev.sample_with_negs(batch_size_p, k, n_batches)

    # ---- the reference's methods
    def sample(self, batch_size_p, strict_return_shape=True):
        if strict_return_shape:
            return self._dev.sample(batch_size_p, 1)[0].cpu().numpy().astype(np.int64)
        # non-strict: batch_size_p // chop whole groups, nothing truncated (:308-311, :334-335)
        n = (batch_size_p // self.chop) * self.chop
        return self._dev.sample(n, 1)[0].cpu().numpy().astype(np.int64)

    def sample_with_negs(self, batch_size_p, k, strict_return_shape=True):
        # the reference's non-strict branch returns an undefined name (`sample_batch`, :408) -> NameError there
        assert strict_return_shape, 'sample_with_negs(strict_return_shape=False) is broken in the reference ' \
                                    '(configs/data_utils.py:408) and not provided'
        out, _ = self._dev.sample_with_negs(batch_size_p, k, 1)
        self._dev.check()
        return out[0].cpu().numpy().astype(np.int64)
