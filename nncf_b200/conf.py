"""Configuration surface of the reference, kept key for key.

ref: configs/basic_embedding_conf.py:10-86 (Conf defaults), :98-142 (get_conf_best, incl. the `reset_after_getconf`
re-apply rule :137-141), :170-179 (get_conf); configs/pretrained_conf.py:10-68 ('mf').  Engine-only keys added here
(`precision`, `replicas`, `optimizer_kind`, `seed`) have defaults that reproduce the reference's behaviour.
"""
from __future__ import annotations


class Conf(object):
    def __init__(self, data_name, param_dict=None):
        self.data_name = data_name
        # loss function related                       (basic_embedding_conf.py:15-27)
        self.max_epoch = 40
        self.batch_size_p = 512
        self.num_negatives = 10
        self.loss = 'skip-gram'
        # NOTE (reference quirk kept): these three defaults are evaluated BEFORE param_dict is applied
        # (basic_embedding_conf.py:22-24 vs :48-49), so overriding `loss` alone does not switch them.
        self.learn_rate = 0.001 if self.loss == 'mse' else 0.01
        self.loss_gamma = 0.1 if self.loss == 'max-margin' else 10
        self.neg_loss_weight = 8 if self.loss == 'mse' else 128
        self.neg_dist = 'unigram'
        self.neg_sampling_power = 1
        self.emb_normalization = None  # set below
        # train scheme related                        (:29-32)
        self.shuffle_st = 'by_item_chop'
        self.chop_size = 2
        self.group_shuffling_trick = True
        # embedding related                           (:34-40)
        self.user_dim = self.item_dim = 50
        self.word_dim = 50
        self.u_reg = 1e-6
        self.c_reg = 0
        self.word_emb_dropout_rate = 0.
        self.pooling = 'average'
        # uninterested                                (:42-46)
        self.eval_topk = 50
        self.interaction_bias = None
        self.use_content_id = False
        self.v_reg = 0
        # pretrained content embedding                (:68-76).  The reference always points at its vector blobs; they are
        # absent here (.MISSING_LARGE_BLOBS), so the dict is only created when the files exist or param_dict supplies one
        # ({'wordvec_filepath', 'sentvec_filepath', 'pretrain_combine_dropout', '..._mode', '..._actv'}; the arrays may
        # also be handed over as 'W_pretrain' / 'C_pretrain').  None = the reference's conf_var 'sup' (supervised only).
        self.pretrain = None
        from .data_utils import get_pretrain_folder
        import os as _os
        folder = get_pretrain_folder(data_name, aug=True)
        if folder is not None and _os.path.exists(folder + 'sentence_vectors_50d.pkl'):
            self.pretrain = {'wordvec_filepath': folder + 'word_vectors_50d.pkl',
                             'sentvec_filepath': folder + 'sentence_vectors_50d.pkl'}
        # engine keys (not in the reference)
        self.precision = 'bf16'          # 'bf16' = tcgen05 tensor cores, 'fp32' = CUDA-core exact mode
        self.replicas = 1                # independent batches per device step (1 = the reference's sequential loop)
        self.optimizer_kind = None       # None -> 'lazy_adam' (the reference trains with Adam); 'sgd' | 'lazy_adam'
        self.seed = 0

        if param_dict is not None:
            self.__dict__.update(param_dict)
        self._post_init()

    def _post_init(self):
        if self.pretrain:
            for key, val in (('pretrain_combine_dropout', 0.5), ('pretrain_combine_mode', 'concat'),
                             ('pretrain_combine_actv', 'relu')):
                self.pretrain.setdefault(key, val)
        if self.emb_normalization is None:
            self.emb_normalization = True \
                if self.loss == 'max-margin' or self.loss == 'log-loss' else False
        # optimizer: the reference builds Adam(lr) for lr > 0 and its own lazy/sparse AdamOptimizer(-lr) for lr < 0
        # (basic_embedding_conf.py:59-66).  Both map to the sparse lazy Adam kernel here (declared deviation from
        # dense Keras Adam); `optimizer_kind='sgd'` selects the atomic scatter-add SGD.
        if self.optimizer_kind is None:
            self.optimizer_kind = 'lazy_adam'
        self.optimizer = (self.optimizer_kind, abs(self.learn_rate))
        self.item_dense_transform = \
            {'dense_hidden_dim': self.user_dim,
             'dense_hidden_dropout': 0.,
             'dense_hidden_actv': 'relu'}
        self.contextual_spatial_gated_input = None
        self.contextual_temporal_gated_input = None


class MfConf(Conf):
    """configs/pretrained_conf.py:10-41, the Conf main.py gives `--model_choice mf` (main.py:47-48): its own defaults
    (64-link batches, 5 negatives, neg_loss_weight 1, uniform negatives, item interaction bias ...).  get_conf_best then
    applies pretrained_conf.py:121-142 (30 epochs, 10 negatives, no bias, per-dataset u_reg)."""

    def __init__(self, data_name, param_dict=None):
        pd = {'max_epoch': 20, 'num_negatives': 5, 'batch_size_p': 64, 'neg_loss_weight': 1, 'interaction_bias': 'item',
              'learn_rate': 0.01, 'u_reg': 1e-5, 'neg_dist': 'uniform', 'neg_sampling_power': 1, 'chop_size': 1,
              'shuffle_st': 'by_item', 'interaction_multiplier': False, 'evaluation_mode': False}
        pd.update(param_dict or {})
        super().__init__(data_name, pd)


class CnnConf(Conf):
    """configs/cnn_embedding_conf.py:10-60: the basic keys plus the convolutional content model's."""

    def __init__(self, data_name, param_dict=None):
        self.num_filters = [50]
        self.filter_lengths = [[1, 3, 5]]
        self.poolings = ['average']
        self.pool_lengths = [-1]
        self.conv_dropout_rate = 0.3
        self.conv_activation = 'relu'
        self.conv_batch_normalization = True
        pd = {'u_reg': 1e-5, 'word_emb_dropout_rate': 0}
        pd.update(param_dict or {})
        super().__init__(data_name, pd)
        for k in ('num_filters', 'filter_lengths', 'poolings', 'pool_lengths', 'conv_dropout_rate', 'conv_activation',
                  'conv_batch_normalization'):
            if param_dict and k in param_dict:
                setattr(self, k, param_dict[k])


class RnnConf(Conf):
    """configs/rnn_embedding_conf.py:10-60: the basic keys plus the recurrent content model's."""

    def __init__(self, data_name, param_dict=None):
        pd = {'u_reg': 1e-5, 'word_emb_dropout_rate': 0., 'rnn': 'lstm', 'bidirection': True, 'lstm_dims': [64],
              'lstm_w_dropout_rate': 0., 'lstm_u_dropout_rate': 0., 'lstm_o_dropout_rate': 0.3, 'pooling': 'average',
              'use_seq_for_dnn': True}
        pd.update(param_dict or {})
        super().__init__(data_name, pd)


class PretrainedConf(MfConf):
    """configs/pretrained_conf.py:10-65 for `--model_choice pretrained`: the 'mf' keys plus the pretrain dict (doc2vec
    item vectors of the un-augmented corpus, `transform` False = use them as the item embedding as they are)."""

    def __init__(self, data_name, param_dict=None):
        super().__init__(data_name, param_dict)
        from .data_utils import get_pretrain_folder
        given = dict(self.pretrain) if isinstance(self.pretrain, dict) else {}
        folder = get_pretrain_folder(data_name, aug=False)
        self.pretrain = {'wordvec_filepath': None,
                         'sentvec_filepath': (folder + 'sentence_vectors_50d.txt') if folder else None,
                         'transform': False, 'pretrain_combine_dropout': 0.5, 'pretrain_combine_actv': 'relu',
                         'pretrain_combine_mode': None}
        self.pretrain.update(given)


CONF_CLASSES = {'pretrained': PretrainedConf, 'mf': MfConf, 'basic_embedding': Conf, 'cnn_embedding': CnnConf, 'rnn_embedding': RnnConf}


def get_conf_default(data_name, param_dict=None, model_choice='basic_embedding'):
    return CONF_CLASSES[model_choice](data_name, param_dict=param_dict)


def get_conf_best(data_name, param_dict=None, model_choice='basic_embedding'):
    """basic_embedding_conf.py:98-142: tuned per-dataset settings, then param_dict re-applied when it carries the key
    `reset_after_getconf` (the demo scripts always pass it, scripts/demos/run_neg_shared.sh:37).  As in the
    reference the re-apply is a plain __dict__.update: derived fields (emb_normalization, optimizer) keep the values
    computed in Conf.__init__, where param_dict had already been applied once."""
    conf = CONF_CLASSES[model_choice](data_name, param_dict=param_dict)
    conf.c_reg = 0
    conf.num_negatives = 10
    if model_choice in ('mf', 'pretrained'):               # pretrained_conf.py:128-142
        conf.max_epoch = 30
        conf.interaction_bias = None
        conf.u_reg = 1e-5 if data_name.startswith('news_title_only') else 1e-6
    elif data_name.startswith('news'):
        conf.max_epoch = 20
        conf.u_reg = 1e-5 if data_name.startswith('news_title_only') else 1e-6
        conf.word_emb_dropout_rate = 0.3
        if conf.pretrain:
            conf.pretrain['pretrain_combine_dropout'] = 0.1
    else:  # citeulike_* and the synthetic stand-ins of the same shape
        conf.max_epoch = 30
        conf.u_reg = 1e-6
        conf.word_emb_dropout_rate = 0.5
        if conf.pretrain:
            conf.pretrain['pretrain_combine_dropout'] = 0.3
    conf_var = (param_dict or {}).get('conf_var')          # basic_embedding_conf.py:131-136
    if conf_var == 'sup':
        conf.pretrain = None
    elif isinstance(conf_var, str) and conf_var.startswith('unsup_dropout') and conf.pretrain:
        conf.pretrain['pretrain_combine_dropout'] = float(conf_var[conf_var.find('=') + 1:])
    try:
        param_dict['reset_after_getconf']
        conf.__dict__.update(param_dict)
    except (TypeError, KeyError):
        pass
    return conf


def get_conf(data_name, conf_choice, param_dict=None, model_choice='basic_embedding'):
    if conf_choice == 'best':
        return get_conf_best(data_name, param_dict, model_choice)
    elif conf_choice in ('default', 'evaluation'):
        return get_conf_default(data_name, param_dict, model_choice)
    else:
        assert False, '[ERROR] conf_choice %s unknown' % conf_choice
