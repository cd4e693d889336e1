"""CNN and RNN item towers as plain torch modules (the north_star keeps them in the framework; they feed the fused
score kernels through the same contract as the mean-pool tower: forward(item_ids) -> float32 [n, d]).

ref: modules/content/cnn_model.py:11-147 (class CNN: embedding => CNN => dense), modules/content/rnn_model.py:11-139
(class RNN: embedding => RNN => pooling => dense), configs/cnn_embedding_conf.py:10-60, configs/rnn_embedding_conf.py:10-60.
Kept: word Embedding with row dropout, 'same' convolutions (he_normal) with one or several filter lengths per layer
concatenated, optional BatchNorm + activation + dropout, average / max pooling (pool_length <= 0 = whole sequence),
Flatten -> Dense -> BatchNorm -> activation -> Dropout; LSTM / GRU stacks, optional second stack reading the sequence
backwards, concatenation, pooling over time when `use_seq_for_dnn`.
Declared (third-party Keras-1.2.2 arithmetic, not under /root/reference): torch's LSTM/GRU gates use sigmoid where the
reference asks Keras for `hard_sigmoid`; recurrent dropout (dropout_W / dropout_U) is not applied.
The contextual gating layers of modules/shared/gatings.py (off in every Conf) are `SpatialGate` / `TemporalGate` below,
applied to the word embeddings in the reference's order (temporal, then spatial: cnn_model.py:53-59, rnn_model.py:46-52).

Two wrappers complete the item side of the reference graph for every content model (mean-pool, CNN, RNN):
`ContentIdTower` (`use_content_id`: + Emb_Cid[cid], ref: modules/content/mean_pool.py:102-108, cnn_model.py:134-140,
rnn_model.py:126-132) and `PretrainCombinedTower` (the supervised / pretrained sentence-vector combination, ref:
modules/shared/vec2vec.py:17-64, applied at models/model_framework.py:99-100).
"""
from __future__ import annotations

import torch


def _actv(name):
    return {'relu': torch.relu, 'tanh': torch.tanh, 'linear': (lambda x: x), 'sigmoid': torch.sigmoid}[name]


class SpatialGate(torch.nn.Module):
    """X' = X * sigmoid(Dense(mean_steps(f(Dense(X))))): one gate per embedding dimension, shared by all steps
    (ref: modules/shared/gatings.py:63-80; conf_dict keys gating_hidden_dim, gating_hidden_actv)"""

    def __init__(self, x_dim, conf_dict):
        super().__init__()
        self.hidden = torch.nn.Linear(x_dim, conf_dict['gating_hidden_dim'])
        self.actv = _actv(conf_dict['gating_hidden_actv'])
        self.out = torch.nn.Linear(conf_dict['gating_hidden_dim'], x_dim)

    def forward(self, x):                                   # [n, steps, x_dim]
        g = torch.sigmoid(self.out(self.actv(self.hidden(x)).mean(dim=1)))
        return x * g[:, None, :]


class TemporalGate(torch.nn.Module):
    """X' = X * c * sigmoid(<X_t, q>), q = Dense(mean_steps(f(Dense(X)))) (optionally after BatchNorm): one gate per step
    (ref: modules/shared/gatings.py:83-118; conf_dict keys gating_hidden_dim, gating_hidden_actv, scale, nl_choice in
    'nl' | 'bn+nl' | 'bn+l'; the reference's softmax variant is dead code there)"""

    def __init__(self, x_dim, conf_dict):
        super().__init__()
        hd = conf_dict['gating_hidden_dim']
        self.hidden = torch.nn.Linear(x_dim, hd)
        self.actv = _actv(conf_dict['gating_hidden_actv'])
        nl = conf_dict['nl_choice']
        assert nl in ('nl', 'bn+nl', 'bn+l'), 'nonononon'
        self.bn = torch.nn.BatchNorm1d(hd, eps=1e-3, momentum=0.01) if nl.startswith('bn') else None
        self.out = torch.nn.Linear(hd, x_dim)
        self.out_actv = torch.relu if nl.endswith('+nl') or nl == 'nl' else (lambda t: t)
        self.c = torch.nn.Parameter(torch.ones(1)) if conf_dict.get('scale') else None

    def forward(self, x):                                   # [n, steps, x_dim]
        q = self.actv(self.hidden(x)).mean(dim=1)
        if self.bn is not None:
            q = self.bn(q)
        q = self.out_actv(self.out(q))                      # [n, x_dim]
        g = torch.sigmoid((x * q[:, None, :]).sum(dim=-1, keepdim=True))      # [n, steps, 1]
        y = x * g
        return y * self.c if self.c is not None else y


class _ContentTower(torch.nn.Module):
    """shared head and tail: word embedding (Keras-1 row dropout) ... Dense -> BN -> actv -> dropout"""

    def __init__(self, data_spec, conf, content, generator, flat_dim):
        super().__init__()
        dev = content.device
        self.temporal_gate = TemporalGate(conf.word_dim, conf.contextual_temporal_gated_input) \
            if conf.contextual_temporal_gated_input else None
        self.spatial_gate = SpatialGate(conf.word_dim, conf.contextual_spatial_gated_input) \
            if conf.contextual_spatial_gated_input else None
        w = (torch.rand((data_spec.word_count, conf.word_dim), generator=generator, device=dev) - 0.5) * 0.1
        if getattr(data_spec, 'W_pretrain', None) is not None:
            w = torch.as_tensor(data_spec.W_pretrain, dtype=torch.float32, device=dev)
        self.word_embedding = torch.nn.Parameter(w)
        self.word_dropout = float(conf.word_emb_dropout_rate)
        self.content = content
        idt = conf.item_dense_transform
        self.dense = self.bn = None
        if idt:
            self.dense = torch.nn.Linear(flat_dim, idt['dense_hidden_dim'])
            if not getattr(conf, 'no_BN', False):
                self.bn = torch.nn.BatchNorm1d(idt['dense_hidden_dim'], eps=1e-3, momentum=0.01)
            self.dense_actv = _actv(idt['dense_hidden_actv'])
            self.dense_dropout = torch.nn.Dropout(idt['dense_hidden_dropout'])

    def embed(self, item_ids):
        W = self.word_embedding
        if self.training and self.word_dropout > 0:
            keep = (torch.rand((W.shape[0], 1), device=W.device) >= self.word_dropout).float() / (1.0 - self.word_dropout)
            W = W * keep
        x = W[self.content[item_ids.long()].long()]                     # [n, L, word_dim]
        if self.temporal_gate is not None:
            x = self.temporal_gate(x)
        if self.spatial_gate is not None:
            x = self.spatial_gate(x)
        return x

    def head(self, h):
        if self.dense is None:
            return h.reshape(h.shape[0], -1)
        h = self.dense(h.reshape(h.shape[0], -1))
        if self.bn is not None:
            h = self.bn(h)
        return self.dense_dropout(self.dense_actv(h))


class CNNTower(_ContentTower):
    def __init__(self, data_spec, conf, content, generator):
        L = content.shape[1]
        layers, in_ch, steps = [], conf.word_dim, L
        for l, fl in enumerate(conf.filter_lengths):
            fls = fl if isinstance(fl, list) else [fl]
            layers.append((in_ch, conf.num_filters[l], fls))
            in_ch = conf.num_filters[l] * len(fls)
            pool = steps if conf.pool_lengths[l] <= 0 else conf.pool_lengths[l]
            steps = steps // pool
        super().__init__(data_spec, conf, content, generator, in_ch * steps)
        self.convs = torch.nn.ModuleList()
        self.bns = torch.nn.ModuleList()
        for cin, nf, fls in layers:
            group = torch.nn.ModuleList()
            for f in fls:
                assert f % 2 == 1, "'same' padding with an even filter length is not provided"
                c = torch.nn.Conv1d(cin, nf, f, padding=f // 2)
                torch.nn.init.kaiming_normal_(c.weight)                  # init='he_normal'
                torch.nn.init.zeros_(c.bias)
                group.append(c)
            self.convs.append(group)
            use_bn = conf.conv_batch_normalization and not getattr(conf, 'no_BN', False)
            self.bns.append(torch.nn.BatchNorm1d(nf * len(fls), eps=1e-3, momentum=0.01) if use_bn else torch.nn.Identity())
        self.conv_actv = _actv(conf.conv_activation)
        self.conv_dropout = torch.nn.Dropout(conf.conv_dropout_rate)
        self.poolings, self.pool_lengths = list(conf.poolings), list(conf.pool_lengths)

    def forward(self, item_ids):
        h = self.embed(item_ids).transpose(1, 2)                         # [n, channels, steps]
        for l, group in enumerate(self.convs):
            h = torch.cat([c(h) for c in group], dim=1)
            h = self.conv_dropout(self.conv_actv(self.bns[l](h)))
            pool = h.shape[2] if self.pool_lengths[l] <= 0 else self.pool_lengths[l]
            if self.poolings[l] == 'average':
                h = torch.nn.functional.avg_pool1d(h, pool)
            elif self.poolings[l] == 'max':
                h = torch.nn.functional.max_pool1d(h, pool)
            else:
                assert False, '[ERROR] unknown pooling %s' % self.poolings[l]
        return self.head(h.transpose(1, 2))                              # Flatten over (steps, channels) like Keras


class RNNTower(_ContentTower):
    def __init__(self, data_spec, conf, content, generator):
        ndir = 2 if conf.bidirection else 1
        out = conf.lstm_dims[-1] * ndir
        flat = out if conf.use_seq_for_dnn or True else out
        super().__init__(data_spec, conf, content, generator, flat)
        kind = conf.rnn.lower()
        assert kind in ('lstm', 'gru'), 'ERROR! check conf.rnn %s' % conf.rnn
        cell = torch.nn.LSTM if kind == 'lstm' else torch.nn.GRU
        self.stacks = torch.nn.ModuleList()
        for _ in range(ndir):
            stack, cin = torch.nn.ModuleList(), conf.word_dim
            for hdim in conf.lstm_dims:
                stack.append(cell(cin, hdim, batch_first=True))
                cin = hdim
            self.stacks.append(stack)
        self.out_dropout = torch.nn.Dropout(conf.lstm_o_dropout_rate)
        self.use_seq, self.pooling = bool(conf.use_seq_for_dnn), conf.pooling

    def forward(self, item_ids):
        x = self.embed(item_ids)
        outs = []
        for d, stack in enumerate(self.stacks):
            h = torch.flip(x, dims=[1]) if d == 1 else x                 # go_backwards: the second stack reads right to left
            for i, rnn in enumerate(stack):
                h, _ = rnn(h)
                if not self.use_seq and i == len(stack) - 1:
                    h = h[:, -1]                                         # return_sequences=False on the last layer
                h = self.out_dropout(h)
            outs.append(h)
        h = torch.cat(outs, dim=-1)
        if self.use_seq:
            if self.pooling == 'average':
                h = h.mean(dim=1)
            elif self.pooling == 'max':
                h = h.max(dim=1).values
            else:
                assert False, 'pooling %s not recognized.' % self.pooling
        return self.head(h)


class ContentIdTower(torch.nn.Module):
    """`use_content_id`: h + Emb_Cid[cid] with Emb_Cid [item_count, item_dim] (Keras-1 'uniform' init) and the activity
    regulariser `v_reg * sum_d mean_b E[b, d]^2` on its output (ref: modules/content/mean_pool.py:102-108;
    utils/utilities.py:122-135).  `reg_loss` holds the regulariser of the last training forward (None when v_reg = 0):
    the trainer adds it to the backward pass of the tower."""

    def __init__(self, tower, item_count, item_dim, v_reg, generator=None, device=None):
        super().__init__()
        self.tower = tower
        dev = device if device is not None else next(tower.parameters()).device
        self.emb_cid = torch.nn.Parameter((torch.rand((item_count, item_dim), generator=generator, device=dev) - 0.5) * 0.1)
        self.v_reg = float(v_reg)
        self.reg_loss = None

    def forward(self, item_ids):
        e = self.emb_cid[item_ids.long()]
        self.reg_loss = self.v_reg * (e * e).mean(dim=0).sum() if (self.training and self.v_reg > 0) else None
        h = self.tower(item_ids)
        assert h.shape[1] == e.shape[1], 'use_content_id: the content model must end in item_dim columns'
        return h + e


class PretrainCombinedTower(torch.nn.Module):
    """ItemCombination (ref: modules/shared/vec2vec.py:17-64): frozen pretrained item vectors `C_pretrain[cid]` (Keras-1
    Embedding dropout = whole table rows dropped, the rest rescaled by 1/(1-p); skipped entirely when the dropout is >= 1)
    merged with the content model's output (`pretrain_combine_mode`: concat | sum | mul | ave | max), then
    Dense(user_dim) -> activation.  No BatchNorm: the reference forces `conf.no_BN = True` right before its test
    (vec2vec.py:52-59).  `tower=None` is the 'pretrained' model's transform branch (models/model_framework.py:78-79)."""

    def __init__(self, tower, C_pretrain, conf, tower_dim=None):
        super().__init__()
        pre = conf.pretrain
        self.tower = tower
        self.p_drop = float(pre.get('pretrain_combine_dropout', 0.5))
        self.mode = pre.get('pretrain_combine_mode') or 'concat'       # (None in configs/pretrained_conf.py:63: nothing to merge there)
        self.actv = _actv(pre.get('pretrain_combine_actv', 'relu'))
        assert self.mode in ('concat', 'sum', 'mul', 'ave', 'max'), 'unknown pretrain_combine_mode %s' % self.mode
        C = torch.as_tensor(C_pretrain, dtype=torch.float32)
        self.register_buffer('c_pretrain', C)
        pdim = C.shape[1]
        if tower is None:
            assert self.p_drop < 1, 'pretrained transform without a content model needs the pretrained vectors'
            in_dim = pdim
        elif self.p_drop >= 1:
            in_dim = tower_dim
        elif self.mode == 'concat':
            in_dim = tower_dim + pdim
        else:
            assert tower_dim == pdim, 'merge mode %s needs equal widths (%d vs %d)' % (self.mode, tower_dim, pdim)
            in_dim = pdim
        self.dense = torch.nn.Linear(in_dim, conf.user_dim)
        self.reg_loss = None

    def forward(self, item_ids):
        h = self.tower(item_ids) if self.tower is not None else None
        self.reg_loss = getattr(self.tower, 'reg_loss', None)
        if self.p_drop < 1:
            p = self.c_pretrain[item_ids.long()]
            if self.training and self.p_drop > 0:
                # a mask per TABLE row: every occurrence of an item in the batch shares it (the batch holds unique ids anyway)
                keep = (torch.rand((self.c_pretrain.shape[0], 1), device=p.device) >= self.p_drop).float() / (1.0 - self.p_drop)
                p = p * keep[item_ids.long()]
            if h is None:
                h = p
            elif self.mode == 'concat':
                h = torch.cat([h, p], dim=1)
            elif self.mode == 'sum':
                h = h + p
            elif self.mode == 'mul':
                h = h * p
            elif self.mode == 'ave':
                h = 0.5 * (h + p)
            else:
                h = torch.maximum(h, p)
        return self.actv(self.dense(h))


class FrozenItemTable(torch.nn.Module):
    """model_choice 'pretrained' without transform: item embeddings are the frozen pretrained vectors themselves
    (ref: models/model_framework.py:80-83: Embedding(..., trainable=False, weights=[C_pretrain]))."""

    def __init__(self, C_pretrain):
        super().__init__()
        self.register_buffer('table', torch.as_tensor(C_pretrain, dtype=torch.float32))
        self.reg_loss = None

    def forward(self, item_ids):
        return self.table[item_ids.long()]
