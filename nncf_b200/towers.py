"""CNN and RNN item towers as plain torch modules (the north_star keeps them in the framework; they feed the fused
score kernels through the same contract as the mean-pool tower: forward(item_ids) -> float32 [n, d]).

ref: modules/content/cnn_model.py:11-147 (class CNN: embedding => CNN => dense), modules/content/rnn_model.py:11-139
(class RNN: embedding => RNN => pooling => dense), configs/cnn_embedding_conf.py:10-60, configs/rnn_embedding_conf.py:10-60.
Kept: word Embedding with row dropout, 'same' convolutions (he_normal) with one or several filter lengths per layer
concatenated, optional BatchNorm + activation + dropout, average / max pooling (pool_length <= 0 = whole sequence),
Flatten -> Dense -> BatchNorm -> activation -> Dropout; LSTM / GRU stacks, optional second stack reading the sequence
backwards, concatenation, pooling over time when `use_seq_for_dnn`.
Declared (third-party Keras-1.2.2 arithmetic, not under /root/reference): torch's LSTM/GRU gates use sigmoid where the
reference asks Keras for `hard_sigmoid`; recurrent dropout (dropout_W / dropout_U) is not applied; the contextual gating
layers (modules/shared/gatings.py, off by default) and `use_content_id` are not provided.
"""
from __future__ import annotations

import torch


def _actv(name):
    return {'relu': torch.relu, 'tanh': torch.tanh, 'linear': (lambda x: x), 'sigmoid': torch.sigmoid}[name]


class _ContentTower(torch.nn.Module):
    """shared head and tail: word embedding (Keras-1 row dropout) ... Dense -> BN -> actv -> dropout"""

    def __init__(self, data_spec, conf, content, generator, flat_dim):
        super().__init__()
        dev = content.device
        assert not conf.contextual_temporal_gated_input and not conf.contextual_spatial_gated_input, \
            'contextual gating (modules/shared/gatings.py) is not provided'
        assert not conf.use_content_id, 'use_content_id is not provided'
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
        return W[self.content[item_ids.long()].long()]                  # [n, L, word_dim]

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
