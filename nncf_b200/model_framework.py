"""get_model(conf, data_helper, model_name) -> model_dict, the reference's model surface.

ref: models/model_framework.py:18-206.  The reference builds ONE shared Keras graph and nine compiled views of it;
here one `SharedState` (embedding tables + optional content tower, all torch CUDA tensors) is shared by view objects
that expose the Keras methods the trainers and the Evaluator call:

  model_dict['model']                   .train_on_batch([uid, cid], [resp])      'mul' view, (1+k)B listed pairs
  model_dict['model_neg_shared']        .train_on_batch / .predict_on_batch([users, items]) -> [U, M]
  model_dict['model_group_neg_shared']  .train_on_batch
  model_dict['model_user_emb'] / ['model_item_emb']   .predict_on_batch([ids]) -> embeddings
  model_dict['model_pred_pairs']        .predict([Uemb, Vemb, uid, cid], batch_size) -> [n, 1]

All arithmetic is done by libnncf_b200.so kernels; the Dense/BatchNorm/ReLU of the mean-pool tower (and the CNN/RNN
towers) stay torch modules feeding the fused score kernel, as the north_star prescribes.
  model_dict['model_sampled_neg_shared'] .train_on_batch([uid, cid])  B positives + k shared sampled negatives
Keys absent here: the two monitor dicts (the reference compiles them with the same loss, models/model_framework.py:179-187).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from .ops import FusedStep, SparseUpdater, StepSpec


def _dev_i32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x).reshape(-1), dtype=np.int32)).to(device)


def _tower_backward(tower, rows, grad_rows):
    """backward through the item tower from the kernel's dL/d(rows); an activity regulariser the tower carries
    (`reg_loss`, e.g. use_content_id's v_reg term) joins the same pass.  Returns the regulariser's value (added to the
    reported loss like Keras adds regularisers)."""
    reg = getattr(tower, 'reg_loss', None)
    if reg is None:
        rows.backward(grad_rows)
        return 0.0
    torch.autograd.backward([rows, reg], [grad_rows, None])
    return float(reg.item())


class MeanPoolTower(torch.nn.Module):
    """word Embedding -> mean over all L positions -> Dense -> BatchNorm -> relu   (modules/content/mean_pool.py:46-110).
    The gather-mean is the nncf_meanpool kernel; BN uses Keras-1 defaults (epsilon 1e-3, momentum 0.99)."""

    def __init__(self, data_spec, conf, content, generator):
        super().__init__()
        dw, d = conf.word_dim, conf.item_dense_transform['dense_hidden_dim']
        w = (torch.rand((data_spec.word_count, dw), generator=generator, device='cuda') - 0.5) * 0.1   # Keras-1 'uniform'
        self.word_embedding = torch.nn.Parameter(w)
        self.dense = torch.nn.Linear(dw, d)
        self.use_bn = not getattr(conf, 'no_BN', False)
        self.bn = torch.nn.BatchNorm1d(d, eps=1e-3, momentum=0.01) if self.use_bn else None
        if self.use_bn:
            # A Dense bias that feeds a BatchNorm has NO gradient (the batch mean removes it); what autograd returns is the
            # rounding error of a zero sum, which Adam's normalisation turns into random +-lr steps per batch - harmless in
            # training mode, but the test phase normalises with running statistics that lag ~100 batches behind the random
            # walk.  The gradient is set to its exact value, zero.  (Declared difference: the reference lets TF do the walk.)
            self.dense.bias.register_hook(torch.zeros_like)
        self.dropout_rate = float(conf.word_emb_dropout_rate)
        self.content = content            # int32 [items, L] on device
        self.actv = conf.item_dense_transform['dense_hidden_actv']

    def forward(self, item_ids):
        W = self.word_embedding
        if self.training and self.dropout_rate > 0:
            # Keras-1 Embedding(dropout=p): whole word rows are dropped and the rest rescaled by 1/(1-p)
            keep = (torch.rand((W.shape[0], 1), device=W.device) >= self.dropout_rate).float() / (1.0 - self.dropout_rate)
            W = W * keep
        h = ops.MeanPoolFunction.apply(W, self.content, item_ids)
        h = self.dense(h)
        if self.bn is not None:
            h = self.bn(h)
        if self.actv == 'relu':
            h = torch.relu(h)
        elif self.actv == 'tanh':
            h = torch.tanh(h)
        return h


class BiasedTower(torch.nn.Module):
    """item tower output [n, d] -> [n, d + 2] = (embedding, 1, cbias[id]): the interaction-bias columns of the item side
    (ref: modules/interaction/interaction_dot.py:64-69,104-107; init 'zero').  cbias is trained by the tower's optimizer."""

    def __init__(self, tower, item_count, item_bias):
        super().__init__()
        self.tower = tower
        self.cbias = torch.nn.Parameter(torch.zeros(item_count, device='cuda'), requires_grad=item_bias)

    def forward(self, item_ids):
        h = self.tower(item_ids)
        self.reg_loss = getattr(self.tower, 'reg_loss', None)
        return torch.cat([h, torch.ones_like(h[:, :1]), self.cbias[item_ids.long()][:, None]], dim=1)


class SharedState(object):
    """Parameters shared by every view (the reference's single Keras graph)."""

    def __init__(self, conf, data_helper, model_name):
        assert torch.cuda.is_available(), '[ERROR] nncf_b200 needs a CUDA device (there is no CPU fallback)'
        self.conf = conf
        self.model_name = model_name
        spec = data_helper.data_spec
        self.data_spec = spec
        dev = torch.device('cuda')
        self.device = dev
        g = torch.Generator(device='cuda').manual_seed(int(getattr(conf, 'seed', 0)) + 7)
        d = conf.user_dim
        assert conf.user_dim == conf.item_dim, 'dot-product interaction needs user_dim == item_dim'
        assert conf.interaction_bias in ['user', 'item', 'both', None], \
            "ERROR! Unknown interation bias {}".format(conf.interaction_bias)
        # interaction bias (modules/interaction/interaction_dot.py:96-107): two extra columns per row, users (ubias, 1),
        # items (1, cbias), biases initialised to zero; the kernels treat them as bias plumbing (include/nncf_b200.h)
        self.bias = conf.interaction_bias
        self.emb_dim = d
        self.dim = d + 2 if self.bias else d

        def with_bias_cols(t, ones_first):
            if not self.bias:
                return t
            z, o = torch.zeros_like(t[:, :1]), torch.ones_like(t[:, :1])
            return torch.cat([t, o, z] if ones_first else [t, z, o], dim=1).contiguous()
        # Keras-1 Embedding init 'uniform' = U(-0.05, 0.05)
        self.user_table = with_bias_cols((torch.rand((spec.user_count, d), generator=g, device=dev) - 0.5) * 0.1, False)
        self.opt_kind, self.lr = conf.optimizer
        self.tower = None
        self.item_table = None
        if model_name == 'mf':
            self.item_table = with_bias_cols((torch.rand((spec.item_count, d), generator=g, device=dev) - 0.5) * 0.1, True)
        elif model_name in ('basic_embedding', 'cnn_embedding', 'rnn_embedding'):
            content = torch.from_numpy(np.ascontiguousarray(data_helper.data['C'], dtype=np.int32)).to(dev)
            if model_name == 'basic_embedding':
                self.tower = MeanPoolTower(spec, conf, content, g).to(dev)
            else:
                from .towers import CNNTower, RNNTower          # plain torch modules (models/model_framework.py:89-96)
                self.tower = (CNNTower if model_name == 'cnn_embedding' else RNNTower)(spec, conf, content, g).to(dev)
            if conf.use_content_id:                 # ref: modules/content/mean_pool.py:102-108 (same in the CNN / RNN models)
                from .towers import ContentIdTower
                self.tower = ContentIdTower(self.tower, spec.item_count, conf.item_dim, conf.v_reg, g, dev)
            if getattr(spec, 'C_pretrain', None) is not None and getattr(conf, 'pretrain', None):
                # ref: models/model_framework.py:99-100 -> modules/shared/vec2vec.py:17-64
                from .towers import PretrainCombinedTower
                self.tower.eval()
                with torch.no_grad():
                    tower_dim = int(self.tower(torch.zeros(2, dtype=torch.int32, device=dev)).shape[1])
                self.tower = PretrainCombinedTower(self.tower, spec.C_pretrain, conf, tower_dim).to(dev)
            if self.bias:
                self.tower = BiasedTower(self.tower, spec.item_count, self.bias in ('item', 'both'))
            # Keras Adam(lr) on the tower's dense parameters: one launch for all tensors, step count on the device, so the
            # captured mean-pool step (below) and the eager paths share one optimizer state
            self.tower_opt = ops.KerasAdam([p for p in self.tower.parameters() if p.requires_grad], lr=self.lr, eps=1e-8)
        elif model_name == 'pretrained':
            # ref: models/model_framework.py:69-84 + configs/pretrained_conf.py:57-65,77-105
            from .towers import FrozenItemTable, PretrainCombinedTower
            pre = conf.pretrain or {}
            if getattr(conf, 'evaluation_mode', False):        # given user / item embeddings, nothing trainable
                ue, ie = np.asarray(pre['user_emb'], dtype=np.float32), np.asarray(pre['item_emb'], dtype=np.float32)
                assert ue.shape == (spec.user_count, d) and ie.shape[1] == d, 'evaluation_mode: embedding shapes'
                self.user_table = with_bias_cols(torch.from_numpy(ue).to(dev), False)
                self.tower = FrozenItemTable(ie).to(dev)
                self.opt_kind, self.lr = 'sgd', 0.0                                          # conf.optimizer = SGD(0)
            else:
                assert spec.C_pretrain is not None, \
                    '[ERROR] model_choice pretrained needs pretrained item vectors (conf.pretrain sentvec_filepath / C_pretrain)'
                if pre.get('transform'):
                    self.tower = PretrainCombinedTower(None, spec.C_pretrain, conf).to(dev)
                else:
                    assert spec.C_pretrain.shape[1] == d, 'pretrained item vectors must have item_dim columns without transform'
                    self.tower = FrozenItemTable(spec.C_pretrain).to(dev)
            if self.bias:
                self.tower = BiasedTower(self.tower, spec.item_count, self.bias in ('item', 'both'))
            params = [p for p in self.tower.parameters() if p.requires_grad]
            self.tower_opt = ops.KerasAdam(params, lr=self.lr, eps=1e-8) if params else None
        else:
            assert False, '[ERROR] Model name {} unknown'.format(model_name)
        self.norm_u = bool(conf.emb_normalization)
        # reference quirk: the item-side l2_normalize sits inside the content-model `else:` branch
        # (models/model_framework.py:85-111), so neither 'mf' nor 'pretrained' items are normalised
        self.norm_v = bool(conf.emb_normalization) and model_name not in ('mf', 'pretrained')
        self.adam = None
        if self.opt_kind == 'lazy_adam':
            z = torch.zeros_like
            self.adam = [z(self.user_table), z(self.user_table),
                         z(self.item_table) if self.item_table is not None else None,
                         z(self.item_table) if self.item_table is not None else None]
        self._steps = {}
        self._user_updater = None

    def step(self, scheme):
        if scheme not in self._steps:
            c = self.conf
            self._steps[scheme] = FusedStep(StepSpec(
                scheme=scheme, loss=c.loss, precision=c.precision, batch_size_p=c.batch_size_p,
                num_negatives=c.num_negatives, dim=self.dim,
                # reference quirk: the sampled_neg_shared view scores U_emb_front = Emb_U(uid_front) WITHOUT l2_normalize
                # (models/model_framework.py:64-65,138-141); the item side is normalised as everywhere else
                norm_u=self.norm_u and scheme != 'sampled_neg_shared', norm_v=self.norm_v,
                optimizer=self.opt_kind, replicas=(c.replicas if self.item_table is not None else 1),
                neg_loss_weight=float(c.neg_loss_weight), loss_gamma=float(c.loss_gamma), u_reg=float(c.u_reg),
                learn_rate=float(self.lr), interaction_bias=self.bias))
        return self._steps[scheme]

    def _norm_emb(self, rows):
        """l2-normalise the embedding columns only (the bias columns pass through)"""
        if not self.bias:
            return torch.nn.functional.normalize(rows, dim=-1, eps=1e-6)
        e = self.emb_dim
        return torch.cat([torch.nn.functional.normalize(rows[:, :e], dim=-1, eps=1e-6), rows[:, e:]], dim=1)

    # ---- embeddings as the views see them (test phase: dropout off, BN running statistics)
    def user_emb(self, ids):
        rows = ops.gather_rows(self.user_table, ids)
        return self._norm_emb(rows) if self.norm_u else rows

    def item_emb(self, ids):
        if self.item_table is not None:
            rows = ops.gather_rows(self.item_table, ids)
        else:
            self.tower.eval()
            with torch.no_grad():
                rows = self.tower(ids)
        return self._norm_emb(rows) if self.norm_v else rows


class MeanPoolGraphStep(object):
    """One neg_shared / group_neg_shared training step of the mean-pool content model (`basic_embedding`) as ONE replayed
    CUDA graph: tf.unique, the segmented gather-mean kernel, Dense -> BatchNorm over the n_u unique items -> relu, the fused
    score kernels, the tower's backward, both optimizers.  Nothing of a step touches the host: shapes are sized by B with the
    n_u valid rows masked (n_u stays on the device), forward and backward are written out explicitly (no autograd tape), the
    lazy-Adam step clock of the user table lives in device memory (nncf_trainer_set_device_clock), the batch losses
    accumulate on the device and are read once per epoch.  The eager path it replaces spent 2.15 ms per step in host syncs
    (`n_u.item()`, `loss.item()`), the autograd tape and ~60 launches.
    ref: modules/content/mean_pool.py:46-110 (tower), models/model_framework.py:45-56,98-111 (unique items, re-expansion)."""

    def __init__(self, state, scheme):
        self.state, self.scheme = state, scheme
        st, conf = state, state.conf
        B = conf.batch_size_p
        dev = st.device
        self.B = B
        self.ids = torch.zeros((2, B), dtype=torch.int32, device=dev)    # the graph reads its batch here: one copy per step
        self.uid, self.cid = self.ids[0], self.ids[1]
        self.arange = torch.arange(B, device=dev, dtype=torch.int32)
        self.loss_sum = torch.zeros((), dtype=torch.float64, device=dev)
        self.graph = None
        self.calls = 0
        self.step = st.step(scheme)
        if st.opt_kind == 'lazy_adam':
            self.step.set_device_clock(True)

    @staticmethod
    def eligible(state):
        t = state.tower
        return (type(t) is MeanPoolTower and t.dropout_rate == 0.0 and t.actv in ('relu', 'tanh', 'linear')
                and not state.bias and state.tower_opt is not None)

    def _body(self):
        st, t = self.state, self.state.tower
        W, Wd, bd = t.word_embedding, t.dense.weight, t.dense.bias
        uq, inv, nuq = ops.unique_first_occurrence(self.cid)
        inv64 = inv.long()
        h0 = ops.meanpool_fwd_n(W, t.content, uq, nuq)                   # [B, dw] mean of the word rows; zero rows beyond n_u
        h1 = torch.addmm(bd, h0, Wd.t())                                 # Dense
        # BatchNorm over the n_u unique items (Keras eps 1e-3, momentum 0.99) + activation, one kernel; rows beyond n_u: zeros
        h3, xhat, rstd = ops.tower_bn_act_fwd(h1, nuq, t.bn, t.actv)
        if self.scheme == 'group_neg_shared':
            out = self.step.run(st.user_table, None, self.uid, self.cid, 1, adam_state=st.adam, want_grads=True,
                                item_rows=h3, inverse=inv, n_unique=nuq)
            g3 = out['grad_item_rows']                                   # (the backward kernel ignores rows beyond n_u)
        else:
            rows = h3[inv64]                                             # C_emb = C_emb_compact[cid_x]
            out = self.step.run(st.user_table, None, self.uid, self.cid, 1, adam_state=st.adam, want_grads=True, item_rows=rows)
            g3 = torch.zeros_like(h3).index_add_(0, inv64, out['grad_item_rows'])
        # ---- backward of the tower, written out (no autograd tape)
        dh1, dgamma, dbeta = ops.tower_bn_act_bwd(g3, h3, xhat, rstd, nuq, t.bn, t.actv)
        if t.bn is not None:
            t.bn.weight.grad, t.bn.bias.grad = dgamma, dbeta
        t.dense.weight.grad = dh1.t() @ h0
        t.dense.bias.grad = dh1.sum(0) if t.bn is None else None     # (exactly zero in front of a BatchNorm, see MeanPoolTower: Adam skips it)
        dh0 = dh1 @ Wd
        dW = torch.zeros_like(W)
        ops.meanpool_bwd_n(dW, t.content, uq, nuq, dh0)
        W.grad = dW
        st.tower_opt.step()
        self.loss_sum += out['loss'][0].double()

    def run(self, uid, cid=None):
        """one step on the B ids of `uid` / `cid` (device int32), or on a [2, B] tensor holding both; returns nothing - see
        take_loss()"""
        if cid is None:
            self.ids.copy_(uid, non_blocking=True)
        else:
            self.uid.copy_(uid, non_blocking=True)
            self.cid.copy_(cid, non_blocking=True)
        self.state.tower.train()
        self.calls += 1
        if self.graph is not None:
            self.graph.replay()
        elif self.calls <= 3:
            with torch.no_grad():
                self._body()                                            # warm-up (real steps): optimizer state, workspaces, attributes
        else:
            import gc
            gc.collect()                                                # (a CUDAGraph freed by the collector DURING a capture invalidates the capture)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                self._body()
            self.graph = g
            g.replay()

    def take_loss(self):
        """sum of the batch losses since the last call (one device -> host read)"""
        v = float(self.loss_sum.item())
        self.loss_sum.zero_()
        return v


class _View(object):
    def __init__(self, state):
        self.state = state


class MatmulView(_View):
    """model_neg_shared / model_group_neg_shared  (models/model_framework.py:126-136,151-161)."""

    def __init__(self, state, scheme):
        super().__init__(state)
        self.scheme = scheme

    def train_on_batches(self, user_ids, item_ids, n_steps, loss_out=None):
        """Device-resident fast path: `n_steps` consecutive steps over ids already in HBM (embedding-table models)."""
        st = self.state
        assert st.item_table is not None
        return st.step(self.scheme).run(st.user_table, st.item_table, user_ids, item_ids, n_steps, adam_state=st.adam,
                                        loss_out=loss_out)['loss']

    def train_on_batch(self, x, y=None):
        """Keras signature: x = [user_batch, item_batch] (host or device int arrays), y = [response] (ignored: the
        matmul views define positives by position).  Returns the python float loss, like Keras."""
        st = self.state
        uid, cid = _dev_i32(x[0], st.device), _dev_i32(x[1], st.device)
        if st.item_table is not None:
            loss = self.train_on_batches(uid, cid, 1)
            return float(loss.mean().item())
        return self._train_with_tower(uid, cid)

    def train_tower_batches(self, user_ids, item_ids, rows_per_batch):
        """all batches of `user_ids` / `item_ids` (device int32, back to back) through the content tower; returns
        (sum of the batch losses, batches).  The mean-pool model runs each batch as one replayed CUDA graph."""
        st = self.state
        nb = user_ids.numel() // rows_per_batch
        fast = (rows_per_batch == st.conf.batch_size_p and MeanPoolGraphStep.eligible(st)
                and os.environ.get('NNCF_TOWER_GRAPH', '1') != '0')
        if not fast:
            cost = 0.0
            for b in range(nb):
                s = slice(b * rows_per_batch, (b + 1) * rows_per_batch)
                cost += self._train_with_tower(user_ids[s], item_ids[s])
            return cost, nb
        key = ('graph', self.scheme)
        if key not in st._steps:
            st._steps[key] = MeanPoolGraphStep(st, self.scheme)
        gs = st._steps[key]
        both = torch.stack([user_ids[:nb * rows_per_batch].view(nb, rows_per_batch),
                            item_ids[:nb * rows_per_batch].view(nb, rows_per_batch)], dim=1)     # [nb, 2, B]: one copy per batch
        for b in range(nb):
            gs.run(both[b])
        return gs.take_loss(), nb

    def _train_with_tower(self, uid, cid):
        st = self.state
        step = st.step(self.scheme)
        st.tower.train()
        uq, inv, nuq = ops.unique_first_occurrence(cid)
        n_u = int(nuq.item())
        compact = st.tower(uq[:n_u])                                   # item tower runs once per UNIQUE item
        if self.scheme == 'group_neg_shared':
            rows = compact
            out = step.run(st.user_table, None, uid, cid, 1, adam_state=st.adam, want_grads=True,
                           item_rows=rows.detach(), inverse=inv, n_unique=nuq)
            g = out['grad_item_rows'][:n_u]
        else:
            rows = compact[inv.long()]                                 # C_emb = C_emb_compact[cid_x]
            out = step.run(st.user_table, None, uid, cid, 1, adam_state=st.adam, want_grads=True, item_rows=rows.detach())
            g = out['grad_item_rows']
        reg = 0.0
        if st.tower_opt is not None:
            st.tower_opt.zero_grad(set_to_none=True)
            reg = _tower_backward(st.tower, rows, g)
            st.tower_opt.step()
        return float(out['loss'][0].item()) + reg

    def predict_on_batch(self, x):
        """[users, items] -> float32 [len(users), len(items)] score matrix (host), for API compatibility with
        utils/objectives.py:310.  The evaluator's fast path never materialises this (see Evaluator)."""
        st = self.state
        U = st.user_emb(_dev_i32(x[0], st.device))
        V = st.item_emb(_dev_i32(x[1], st.device))
        nu, ni = U.shape[0], V.shape[0]
        # all (user, item) pairs of the block through the pair-scoring kernel in one launch
        uidx = torch.arange(nu, device=st.device, dtype=torch.int32).repeat_interleave(ni)
        cidx = torch.arange(ni, device=st.device, dtype=torch.int32).repeat(nu)
        return ops.score_pairs(U, V, uidx, cidx).reshape(nu, ni).cpu().numpy()


class PairsView(_View):
    """model: the row-wise 'mul' view used by train_original / train_group_sample (model_framework.py:123,147-149)."""

    def train_on_batches(self, user_ids, item_ids, n_steps, loss_out=None, responses=None):
        """responses: optional int32 y_true per row (1 = positive).  Pointwise losses weight by it (ref:
        utils/objectives.py:59-70), so batches whose positives are not the first B rows (presample's shuffles,
        GroupSampler's ragged batches) train correctly; pairwise losses are positional in the reference too."""
        st = self.state
        assert st.item_table is not None
        if st.conf.loss not in ('skip-gram', 'mse'):
            responses = None
        return st.step('pairs').run(st.user_table, st.item_table, user_ids, item_ids, n_steps, adam_state=st.adam,
                                    loss_out=loss_out, responses=responses)['loss']

    def train_on_batch(self, x, y=None):
        st = self.state
        uid, cid = _dev_i32(x[0], st.device), _dev_i32(x[1], st.device)
        if st.item_table is not None:
            resp = None
            if y is not None and y[0] is not None:
                resp = _dev_i32(y[0], st.device)
            return float(self.train_on_batches(uid, cid, 1, responses=resp).mean().item())
        # content tower: the unique items' embeddings act as a temporary item table indexed by tf.unique's inverse
        c = st.conf
        st.tower.train()
        uq, inv, nuq = ops.unique_first_occurrence(cid)
        n_u = int(nuq.item())
        compact = st.tower(uq[:n_u])
        key = ('pairs_tower',)
        if key not in st._steps:
            st._steps[key] = FusedStep(StepSpec(
                scheme='pairs', loss=c.loss, precision='fp32', batch_size_p=c.batch_size_p, num_negatives=c.num_negatives,
                dim=st.dim, norm_u=st.norm_u, norm_v=st.norm_v, optimizer='none', neg_loss_weight=float(c.neg_loss_weight),
                loss_gamma=float(c.loss_gamma), u_reg=float(c.u_reg), interaction_bias=st.bias))
            st._user_updater = SparseUpdater(st.opt_kind, st.lr)
        # y_true per row for the pointwise losses (ref: utils/objectives.py:59-70 weights by it): presample's shuffles and
        # GroupSampler's ragged batches do not keep the positives in rows [0, B)
        resp = None
        if y is not None and y[0] is not None and c.loss in ('skip-gram', 'mse'):
            resp = _dev_i32(y[0], st.device)
        out = st._steps[key].run(st.user_table, compact.detach().contiguous(), uid, inv, 1, want_grads=True, responses=resp)
        st._user_updater.begin_step()
        st._user_updater.apply(st.user_table, uid, out['grad_user_rows'], *(st.adam[:2] if st.adam else (None, None)))
        g = torch.zeros_like(compact)
        g.index_add_(0, inv.long(), out['grad_item_rows'])
        reg = 0.0
        if st.tower_opt is not None:
            st.tower_opt.zero_grad(set_to_none=True)
            reg = _tower_backward(st.tower, compact, g)
            st.tower_opt.step()
        return float(out['loss'][0].item()) + reg


class SampledNegSharedView(_View):
    """model_sampled_neg_shared: B positive rows + k shared sampled negative items per batch, scores [B, 1+k]
    (models/model_framework.py:138-143,163-167; loss utils/objectives.py:120-161).  Embedding-table models."""

    def train_on_batches(self, user_ids, item_ids, n_steps, loss_out=None):
        st = self.state
        assert st.item_table is not None, 'sampled_neg_shared runs on embedding-table models (mf)'
        return st.step('sampled_neg_shared').run(st.user_table, st.item_table, user_ids, item_ids, n_steps,
                                                 adam_state=st.adam, loss_out=loss_out)['loss']

    def train_on_batch(self, x, y=None):
        st = self.state
        uid, cid = _dev_i32(x[0], st.device), _dev_i32(x[1], st.device)
        return float(self.train_on_batches(uid, cid, 1).mean().item())


class UserEmbView(_View):
    def predict_on_batch(self, x):
        return self.state.user_emb(_dev_i32(x[0], self.state.device)).cpu().numpy()


class ItemEmbView(_View):
    def predict_on_batch(self, x):
        return self.state.item_emb(_dev_i32(x[0], self.state.device)).cpu().numpy()


class PredPairsView(_View):
    def predict(self, x, batch_size=4096):
        """[U_emb_given, C_emb_given, uid, cid] -> [n, 1]: row-wise dot of the given embeddings
        (model_framework.py:125,174-176; utils/objectives.py:245-247)."""
        dev = self.state.device
        U = torch.as_tensor(np.asarray(x[0], dtype=np.float32)).to(dev)
        V = torch.as_tensor(np.asarray(x[1], dtype=np.float32)).to(dev)
        idx = torch.arange(U.shape[0], device=dev, dtype=torch.int32)
        return ops.score_pairs(U, V, idx, idx).cpu().numpy().reshape(-1, 1)


def get_model(conf, data_helper, model_name):
    state = SharedState(conf, data_helper, model_name)
    model_dict = {'model': PairsView(state),
                  'model_neg_shared': MatmulView(state, 'neg_shared'),
                  'model_group_neg_shared': MatmulView(state, 'group_neg_shared'),
                  'model_sampled_neg_shared': SampledNegSharedView(state),
                  'model_user_emb': UserEmbView(state),
                  'model_item_emb': ItemEmbView(state),
                  'model_pred_pairs': PredPairsView(state),
                  '_state': state}
    return model_dict
