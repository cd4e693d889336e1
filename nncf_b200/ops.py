"""Torch-tensor level wrappers over the C-ABI.  torch is used for device memory and streams only; every
computation below happens in libnncf_b200.so's CUDA kernels.  No CPU fallback: tensors must live on a CUDA device."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from ._lib import (BIASES, LOSSES, OPTIMIZERS, PRECISIONS, SCHEMES, NNCFError, StepConfig, StepIO, Tables, check, lib)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NNCFError("nncf_b200 kernels need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t.to(dtype=torch.int32).contiguous()


def launch_count() -> int:
    return int(lib.nncf_launch_count())


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
class DeviceSampler:
    """Alias-table + Philox sampler handle (nncf_sampler_*)."""

    def __init__(self, dist: np.ndarray, power: float = 0.75, seed: int = 0):
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        self.n = int(dist.size)
        h = C.c_void_p()
        check(lib.nncf_sampler_create(dist.ctypes.data_as(C.c_void_p), self.n, float(power), int(seed) & (2**64 - 1),
                                      C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:
            lib.nncf_sampler_destroy(h)
            self._h = None

    def seek(self, counter: int) -> None:
        check(lib.nncf_sampler_seek(self._h, int(counter)))

    def sample_device(self, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(int(n), dtype=torch.int32, device="cuda")
        _need_cuda(out)
        check(lib.nncf_sampler_sample_batch_dev(self._h, int(n), _ptr(out), _stream()))
        return out

    def sample_host(self, n: int) -> np.ndarray:
        out = np.zeros(int(n), dtype=np.int32)
        check(lib.nncf_sampler_sample_batch_host(self._h, int(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def export_table(self):
        prob = np.zeros(self.n, dtype=np.float32)
        alias = np.zeros(self.n, dtype=np.int32)
        check(lib.nncf_sampler_export_table(self._h, prob.ctypes.data_as(C.c_void_p), alias.ctypes.data_as(C.c_void_p)))
        return prob, alias


class DeviceGroupSampler:
    """Handle on nncf_group_sampler_*: the reference's GroupSampler (configs/data_utils.py:244-408) on the device."""

    NEG_DISTS = {"unigram": 0, "uniform": 1, "uniform_no_correction": 2}

    def __init__(self, train: np.ndarray, group_by: str = "item", chop: int = 1, neg_dist: str = "unigram",
                 neg_sign: int = 0, neg_sampling_power: float = 0.75, seed: int = 0):
        if group_by not in ("item", "user"):
            raise AssertionError("[ERROR] Illegal group_by {}".format(group_by))
        if neg_dist not in self.NEG_DISTS:
            raise AssertionError("[ERROR] Illegal neg_dist {}".format(neg_dist))
        if not torch.cuda.is_available():
            raise NNCFError("nncf_b200 needs a CUDA device (no CPU fallback)")
        train = np.ascontiguousarray(train, dtype=np.int32)
        assert train.ndim == 2 and train.shape[1] == 3
        h = C.c_void_p()
        check(lib.nncf_group_sampler_create(train.ctypes.data_as(C.c_void_p), train.shape[0], 0 if group_by == "item" else 1,
                                            int(chop), self.NEG_DISTS[neg_dist], int(neg_sign), float(neg_sampling_power),
                                            int(seed) & (2**64 - 1), C.byref(h)))
        self._h = h
        self.chop = int(chop)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:
            lib.nncf_group_sampler_destroy(h)
            self._h = None

    def sample(self, batch_size_p: int, n_batches: int = 1) -> torch.Tensor:
        """int32 CUDA tensor [n_batches, batch_size_p, 3]"""
        out = torch.empty((int(n_batches), int(batch_size_p), 3), dtype=torch.int32, device="cuda")
        check(lib.nncf_group_sampler_sample(self._h, int(batch_size_p), int(n_batches), _ptr(out), _stream()))
        return out

    def sample_with_negs(self, batch_size_p: int, k: int, n_batches: int = 1):
        """(int32 CUDA tensor [n_batches, batch_size_p * (1 + k), 3], int32 [n_batches] positives per batch)"""
        out = torch.empty((int(n_batches), int(batch_size_p) * (1 + int(k)), 3), dtype=torch.int32, device="cuda")
        n_pos = torch.empty(int(n_batches), dtype=torch.int32, device="cuda")
        check(lib.nncf_group_sampler_sample_with_negs(self._h, int(batch_size_p), int(k), int(n_batches), _ptr(out),
                                                      _ptr(n_pos), _stream()))
        return out, n_pos

    def check(self) -> None:
        """raises where the reference asserts: a batch was not filled within 10 top-up rounds (data_utils.py:368-370)"""
        failed = C.c_int(0)
        check(lib.nncf_group_sampler_check(self._h, C.byref(failed)))
        assert not failed.value, "[WARNING] the code here should be optimized if this is shown."


# ------------------------------------------------------------------------------------------------
# batch builders
# ------------------------------------------------------------------------------------------------
def permute_rows(train: torch.Tensor, row_perm: torch.Tensor) -> torch.Tensor:
    _need_cuda(train, row_perm)
    train = _i32(train)
    row_perm = row_perm.to(torch.int64).contiguous()
    out = torch.empty_like(train)
    check(lib.nncf_permute_rows(_ptr(train), train.shape[0], _ptr(row_perm), _ptr(out), _stream()))
    return out


def group_shuffle(train: torch.Tensor, col: int, iidx: torch.Tensor, row_perm: torch.Tensor,
                  block_perm: Optional[torch.Tensor], chop: int) -> torch.Tensor:
    _need_cuda(train, iidx, row_perm, block_perm)
    train = _i32(train)
    iidx = iidx.to(torch.int64).contiguous()
    row_perm = row_perm.to(torch.int64).contiguous()
    if block_perm is not None:
        block_perm = block_perm.to(torch.int64).contiguous()
    n = train.shape[0]
    out = torch.empty_like(train)
    wsb = int(lib.nncf_group_shuffle_workspace_bytes(n, iidx.numel()))
    ws = torch.empty(wsb, dtype=torch.uint8, device=train.device)
    check(lib.nncf_group_shuffle(_ptr(train), n, int(col), _ptr(iidx), iidx.numel(), _ptr(row_perm), _ptr(block_perm),
                                 int(chop), _ptr(out), _ptr(ws), wsb, _stream()))
    return out


def assemble_pairs_batch(pos: torch.Tensor, k: int, negs: torch.Tensor, neg_col: int, neg_sign: int) -> torch.Tensor:
    _need_cuda(pos, negs)
    pos = _i32(pos)
    negs = _i32(negs)
    B = pos.shape[0]
    out = torch.empty(((1 + k) * B, 3), dtype=torch.int32, device=pos.device)
    check(lib.nncf_assemble_pairs_batch(_ptr(pos), B, int(k), _ptr(negs), int(neg_col), int(neg_sign), _ptr(out), _stream()))
    return out


def presample_assemble(pos: torch.Tensor, k: int, negs: torch.Tensor, neg_col: int, neg_sign: int, layout: int) -> torch.Tensor:
    """(1+k)N rows of the presample trainer: layout 0 = interleaved, 1 = positives then negatives"""
    _need_cuda(pos, negs)
    pos, negs = _i32(pos), _i32(negs)
    n = pos.shape[0]
    assert negs.numel() >= n * k
    out = torch.empty(((1 + k) * n, 3), dtype=torch.int32, device=pos.device)
    check(lib.nncf_presample_assemble(_ptr(pos), n, int(k), _ptr(negs), int(neg_col), int(neg_sign), int(layout), _ptr(out),
                                      _stream()))
    return out


def assemble_sns_batches(train: torch.Tensor, n_batches: int, B: int, k: int, negs: torch.Tensor):
    """id arrays [n_batches * (B + k)] of the sampled_neg_shared trainer"""
    _need_cuda(train, negs)
    train, negs = _i32(train), _i32(negs)
    assert train.shape[0] >= n_batches * B and negs.numel() >= n_batches * k
    uid = torch.empty(n_batches * (B + k), dtype=torch.int32, device=train.device)
    cid = torch.empty_like(uid)
    check(lib.nncf_assemble_sns_batches(_ptr(train), int(n_batches), int(B), int(k), _ptr(negs), _ptr(uid), _ptr(cid), _stream()))
    return uid, cid


def unique_first_occurrence(ids: torch.Tensor):
    _need_cuda(ids)
    ids = _i32(ids)
    n = ids.numel()
    uniq = torch.empty(n, dtype=torch.int32, device=ids.device)
    inv = torch.empty(n, dtype=torch.int32, device=ids.device)
    cnt = torch.empty(1, dtype=torch.int32, device=ids.device)
    check(lib.nncf_unique_first_occurrence(_ptr(ids), n, _ptr(uniq), _ptr(inv), _ptr(cnt), _stream()))
    return uniq, inv, cnt


# ------------------------------------------------------------------------------------------------
# fused training step
# ------------------------------------------------------------------------------------------------
@dataclass
class StepSpec:
    scheme: str = "neg_shared"          # neg_shared | group_neg_shared | pairs | sampled_neg_shared
    loss: str = "skip-gram"
    precision: str = "bf16"             # fp32 (CUDA cores) | bf16 (tcgen05)
    batch_size_p: int = 512
    num_negatives: int = 10
    dim: int = 50
    norm_u: bool = False
    norm_v: bool = False
    optimizer: str = "sgd"              # none | sgd | lazy_adam
    replicas: int = 1
    neg_loss_weight: float = 128.0
    loss_gamma: float = 10.0
    u_reg: float = 0.0
    learn_rate: float = 0.01
    beta1: float = 0.9
    beta2: float = 0.999
    epsilon: float = 1e-8
    interaction_bias: Optional[str] = None   # None | user | item | both: `dim` then counts the two bias columns

    def to_c(self) -> StepConfig:
        if self.loss not in LOSSES:
            raise AssertionError("[ERROR!] loss %s not specified." % self.loss)
        return StepConfig(SCHEMES[self.scheme], LOSSES[self.loss], PRECISIONS[self.precision], self.batch_size_p,
                          self.num_negatives, self.dim, int(self.norm_u), int(self.norm_v), OPTIMIZERS[self.optimizer],
                          self.replicas, self.neg_loss_weight, self.loss_gamma, self.u_reg, self.learn_rate, self.beta1,
                          self.beta2, self.epsilon, BIASES[self.interaction_bias])


class FusedStep:
    """Handle on nncf_trainer_*: owns the staging workspace of one (scheme, loss, B, d, replicas) configuration."""

    def __init__(self, spec: StepSpec):
        if not torch.cuda.is_available():
            raise NNCFError("nncf_b200 needs a CUDA device (no CPU fallback)")
        self.spec = spec
        cfg = spec.to_c()
        h = C.c_void_p()
        check(lib.nncf_trainer_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.rows = ((1 + spec.num_negatives) * spec.batch_size_p if spec.scheme == "pairs" else
                     spec.batch_size_p + spec.num_negatives if spec.scheme == "sampled_neg_shared" else spec.batch_size_p)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:                    # (module globals are gone when this runs at interpreter exit)
            lib.nncf_trainer_destroy(h)
            self._h = None

    def set_device_clock(self, enable: bool) -> None:
        """keep the lazy-Adam step count / lr_t on the device so that a captured step can be replayed as a CUDA graph"""
        check(lib.nncf_trainer_set_device_clock(self._h, int(bool(enable))))

    def set_profile(self, enable: bool) -> None:
        check(lib.nncf_trainer_set_profile(self._h, int(bool(enable))))

    def get_profile(self):
        """(ms_gather, ms_score, ms_finalize) accumulated over `steps` profiled steps."""
        ms = (C.c_double * 3)()
        steps = C.c_int64()
        check(lib.nncf_trainer_get_profile(self._h, ms, C.byref(steps)))
        return [ms[0], ms[1], ms[2]], int(steps.value)

    def run(self, user_table: torch.Tensor, item_table: Optional[torch.Tensor], user_ids: torch.Tensor,
            item_ids: torch.Tensor, n_steps: int = 1, *, adam_state=None, want_grads: bool = False,
            item_rows: Optional[torch.Tensor] = None, inverse: Optional[torch.Tensor] = None,
            n_unique: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None,
            responses: Optional[torch.Tensor] = None):
        """Runs n_steps steps (each over `replicas` batches).  Returns dict(loss=[n_steps*R] tensor, and if
        want_grads: grad_user_rows, grad_item_rows (+ unique_ids, inverse, n_unique for group_neg_shared))."""
        sp = self.spec
        _need_cuda(user_table, item_table, user_ids, item_ids, item_rows)
        assert user_table.dtype == torch.float32 and user_table.is_contiguous()
        assert user_ids.dtype == torch.int32 and item_ids.dtype == torch.int32
        need = n_steps * sp.replicas * self.rows
        assert user_ids.numel() >= need and item_ids.numel() >= need, "not enough ids for n_steps x replicas batches"
        dev = user_table.device
        tb = Tables()
        tb.user_table = user_table.data_ptr()
        tb.n_users = user_table.shape[0]
        if item_table is not None:
            assert item_table.dtype == torch.float32 and item_table.is_contiguous()
            tb.item_table = item_table.data_ptr()
            tb.n_items = item_table.shape[0]
        if adam_state is not None:
            um, uv, im, iv = adam_state
            tb.user_m, tb.user_v = um.data_ptr(), uv.data_ptr()
            if im is not None:
                tb.item_m, tb.item_v = im.data_ptr(), iv.data_ptr()
        io = StepIO()
        if loss_out is None:
            loss_out = torch.empty(n_steps * sp.replicas, dtype=torch.float32, device=dev)
        io.loss_out_dev = loss_out.data_ptr()
        out = {"loss": loss_out}
        keep = []
        if want_grads:
            gu = torch.zeros((self.rows, sp.dim), dtype=torch.float32, device=dev)
            gv = torch.zeros((self.rows, sp.dim), dtype=torch.float32, device=dev)
            io.grad_user_rows_dev, io.grad_item_rows_dev = gu.data_ptr(), gv.data_ptr()
            out["grad_user_rows"], out["grad_item_rows"] = gu, gv
            if sp.scheme == "group_neg_shared" and item_table is not None:
                uq = torch.zeros(self.rows, dtype=torch.int32, device=dev)
                iv_ = torch.zeros(self.rows, dtype=torch.int32, device=dev)
                nu = torch.zeros(1, dtype=torch.int32, device=dev)
                io.unique_ids_dev, io.inverse_dev, io.n_unique_dev = uq.data_ptr(), iv_.data_ptr(), nu.data_ptr()
                out["unique_ids"], out["inverse"], out["n_unique"] = uq, iv_, nu
        if responses is not None:
            # PAIRS, pointwise losses: y_true per row (1 = positive), ref utils/objectives.py:59-70
            _need_cuda(responses)
            assert sp.scheme == "pairs" and responses.dtype == torch.int32 and responses.numel() >= need
            responses = responses.contiguous()
            keep.append(responses)
            io.response_dev = responses.data_ptr()
        if sp.scheme in ("neg_shared", "group_neg_shared") and item_table is not None and user_ids.numel() >= need + sp.replicas * self.rows \
                and item_ids.numel() >= need + sp.replicas * self.rows and user_ids.is_contiguous() and item_ids.is_contiguous():
            # the caller's id arrays go on beyond this call: the rows of its next step are an L2 prefetch hint for the last one
            io.next_user_ids_dev = user_ids.data_ptr() + 4 * need
            io.next_item_ids_dev = item_ids.data_ptr() + 4 * need
        if item_rows is not None:
            item_rows = item_rows.contiguous()
            keep.append(item_rows)
            io.item_rows_dev = item_rows.data_ptr()
            if inverse is not None:
                io.inverse_dev = inverse.data_ptr()
                io.n_unique_dev = n_unique.data_ptr()
        check(lib.nncf_train_steps(self._h, C.byref(tb), _ptr(user_ids), _ptr(item_ids), int(n_steps), C.byref(io), _stream()))
        return out


    def run_host(self, user_table: torch.Tensor, item_table: torch.Tensor, user_ids_host: torch.Tensor,
                 item_ids_host: torch.Tensor, n_steps: int, loss_out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host-fed loop (nncf_train_steps_host): ids are HOST int32 tensors (pinned memory lets the copies overlap the
        kernels), every step's ids are copied H2D and every step's losses D2H, both in chunks of 1, 4, 16, 16, ... steps
        on two copy streams beside the compute stream; returns the host
        loss tensor [n_steps * R] after everything has completed.   ref: models/train_neg_shared.py:46-50"""
        sp = self.spec
        _need_cuda(user_table, item_table)
        assert not user_ids_host.is_cuda and not item_ids_host.is_cuda, "run_host takes HOST id tensors"
        assert user_ids_host.dtype == torch.int32 and item_ids_host.dtype == torch.int32
        assert user_ids_host.is_contiguous() and item_ids_host.is_contiguous()
        need = n_steps * sp.replicas * self.rows
        assert user_ids_host.numel() >= need and item_ids_host.numel() >= need, "not enough ids for n_steps x replicas batches"
        if loss_out_host is None:
            loss_out_host = torch.empty(n_steps * sp.replicas, dtype=torch.float32).pin_memory()
        assert loss_out_host.numel() >= n_steps * sp.replicas and loss_out_host.dtype == torch.float32
        tb = Tables()
        tb.user_table, tb.n_users = user_table.data_ptr(), user_table.shape[0]
        tb.item_table, tb.n_items = item_table.data_ptr(), item_table.shape[0]
        check(lib.nncf_train_steps_host(self._h, C.byref(tb), C.c_void_p(user_ids_host.data_ptr()),
                                        C.c_void_p(item_ids_host.data_ptr()), int(n_steps),
                                        C.c_void_p(loss_out_host.data_ptr()), _stream()))
        return loss_out_host


# ------------------------------------------------------------------------------------------------
# mean-pool encoder
# ------------------------------------------------------------------------------------------------
def meanpool_fwd(word_table: torch.Tensor, content: torch.Tensor, item_ids: Optional[torch.Tensor], n: int) -> torch.Tensor:
    _need_cuda(word_table, content, item_ids)
    out = torch.empty((n, word_table.shape[1]), dtype=torch.float32, device=word_table.device)
    check(lib.nncf_meanpool_fwd(_ptr(word_table), word_table.shape[1], _ptr(content), content.shape[1], _ptr(item_ids), n,
                                _ptr(out), _stream()))
    return out


def meanpool_bwd(grad_word_table: torch.Tensor, content: torch.Tensor, item_ids: Optional[torch.Tensor], n: int,
                 grad_out: torch.Tensor) -> None:
    _need_cuda(grad_word_table, content, item_ids, grad_out)
    grad_out = grad_out.contiguous()
    check(lib.nncf_meanpool_bwd(_ptr(grad_word_table), grad_word_table.shape[1], _ptr(content), content.shape[1],
                                _ptr(item_ids), n, _ptr(grad_out), _stream()))


def meanpool_fwd_n(word_table: torch.Tensor, content: torch.Tensor, item_ids: torch.Tensor, n_valid: torch.Tensor) -> torch.Tensor:
    """mean-pool of `item_ids.numel()` slots of which only the first *n_valid (device int32) exist; the rest come out as zero rows"""
    _need_cuda(word_table, content, item_ids, n_valid)
    n = item_ids.numel()
    out = torch.empty((n, word_table.shape[1]), dtype=torch.float32, device=word_table.device)
    check(lib.nncf_meanpool_fwd_n(_ptr(word_table), word_table.shape[1], _ptr(content), content.shape[1], _ptr(item_ids), n,
                                  _ptr(n_valid), _ptr(out), _stream()))
    return out


def meanpool_bwd_n(grad_word_table: torch.Tensor, content: torch.Tensor, item_ids: torch.Tensor, n_valid: torch.Tensor,
                   grad_out: torch.Tensor) -> None:
    _need_cuda(grad_word_table, content, item_ids, n_valid, grad_out)
    grad_out = grad_out.contiguous()
    check(lib.nncf_meanpool_bwd_n(_ptr(grad_word_table), grad_word_table.shape[1], _ptr(content), content.shape[1],
                                  _ptr(item_ids), item_ids.numel(), _ptr(n_valid), _ptr(grad_out), _stream()))


class KerasAdam(object):
    """Adam for the towers' dense parameters in Keras-1 form (nncf_dense_adam_step: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),
    p -= lr_t m / (sqrt(v) + eps); ref: configs/*_conf.py `optimizer = Adam(lr)`), one launch for all tensors (plus a
    one-thread step-clock kernel), step count on the device: safe inside a captured CUDA graph.  Same two methods the
    trainers use of a torch optimizer: zero_grad() and step().  Parameters without a gradient are skipped, like torch does."""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params = [p for p in params]
        assert all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params), "KerasAdam: contiguous fp32 CUDA parameters"
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(beta1), float(beta2), float(eps)
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        dev = self.params[0].device if self.params else torch.device("cuda")
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_t_dev = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self):
        idx = [i for i, p in enumerate(self.params) if p.grad is not None]
        n = len(idx)
        if n == 0:
            return
        grads = [self.params[i].grad.contiguous() for i in idx]
        arr = lambda ptrs: (C.c_void_p * n)(*ptrs)                      # noqa: E731
        sizes = (C.c_int64 * n)(*[self.params[i].numel() for i in idx])
        check(lib.nncf_dense_adam_step(n, arr([self.params[i].data_ptr() for i in idx]), arr([g.data_ptr() for g in grads]),
                                       arr([self.m[i].data_ptr() for i in idx]), arr([self.v[i].data_ptr() for i in idx]), sizes,
                                       self.lr, self.beta1, self.beta2, self.eps, _ptr(self.step_dev), _ptr(self.lr_t_dev), _stream()))


_ACTS = {"linear": 0, "relu": 1, "tanh": 2}


def tower_bn_act_fwd(h: torch.Tensor, n_valid: torch.Tensor, bn, activation: str):
    """BatchNorm over the first *n_valid rows of h (bn: a torch BatchNorm1d whose parameters / running statistics are used
    and updated, or None) + activation.  Returns (y, xhat, rstd) - what tower_bn_act_bwd needs."""
    _need_cuda(h, n_valid)
    h = h.contiguous()
    rows, d = h.shape
    y, xhat = torch.empty_like(h), torch.empty_like(h)
    rstd = torch.empty(d, dtype=torch.float32, device=h.device)
    if bn is not None:
        check(lib.nncf_tower_bn_act_fwd(_ptr(h), rows, d, _ptr(n_valid), 1, _ACTS[activation], _ptr(bn.weight), _ptr(bn.bias),
                                        float(bn.eps), float(bn.momentum), _ptr(bn.running_mean), _ptr(bn.running_var),
                                        _ptr(y), _ptr(xhat), _ptr(rstd), _stream()))
    else:
        check(lib.nncf_tower_bn_act_fwd(_ptr(h), rows, d, _ptr(n_valid), 0, _ACTS[activation], None, None, 0.0, 0.0, None, None,
                                        _ptr(y), _ptr(xhat), _ptr(rstd), _stream()))
    return y, xhat, rstd


def tower_bn_act_bwd(dy: torch.Tensor, y: torch.Tensor, xhat: torch.Tensor, rstd: torch.Tensor, n_valid: torch.Tensor, bn,
                     activation: str):
    """Returns (dh, dgamma, dbeta); dgamma / dbeta are None without BatchNorm."""
    _need_cuda(dy, y, xhat, rstd, n_valid)
    dy = dy.contiguous()
    rows, d = dy.shape
    dh = torch.empty_like(dy)
    if bn is not None:
        dg = torch.empty(d, dtype=torch.float32, device=dy.device)
        db = torch.empty(d, dtype=torch.float32, device=dy.device)
        check(lib.nncf_tower_bn_act_bwd(_ptr(dy), _ptr(y), _ptr(xhat), _ptr(rstd), rows, d, _ptr(n_valid), 1, _ACTS[activation],
                                        _ptr(bn.weight), _ptr(dh), _ptr(dg), _ptr(db), _stream()))
        return dh, dg, db
    check(lib.nncf_tower_bn_act_bwd(_ptr(dy), _ptr(y), _ptr(xhat), _ptr(rstd), rows, d, _ptr(n_valid), 0, _ACTS[activation], None,
                                    _ptr(dh), None, None, _stream()))
    return dh, None, None


class MeanPoolFunction(torch.autograd.Function):
    """Segmented gather-mean over word rows with a dense-gradient backward (scatter-add kernel)."""

    @staticmethod
    def forward(ctx, word_table, content, item_ids):
        n = item_ids.numel() if item_ids is not None else content.shape[0]
        ctx.save_for_backward(content, item_ids)
        ctx.shape = word_table.shape
        ctx.n = n
        return meanpool_fwd(word_table, content, item_ids, n)

    @staticmethod
    def backward(ctx, grad_out):
        content, item_ids = ctx.saved_tensors
        dW = torch.zeros(ctx.shape, dtype=torch.float32, device=grad_out.device)
        meanpool_bwd(dW, content, item_ids, ctx.n, grad_out)
        return dW, None, None


# ------------------------------------------------------------------------------------------------
# evaluation
# ------------------------------------------------------------------------------------------------
def eval_topk(user_rows: torch.Tensor, item_rows: torch.Tensor, k: int, precision: str = "bf16"):
    _need_cuda(user_rows, item_rows)
    user_rows = user_rows.contiguous().float()
    item_rows = item_rows.contiguous().float()
    nu, d = user_rows.shape
    ni = item_rows.shape[0]
    prec = PRECISIONS[precision]
    wsb = int(lib.nncf_eval_topk_workspace_bytes(nu, ni, d, int(k), prec))
    if wsb == 0:
        raise NNCFError("eval_topk: %s" % lib.nncf_last_error().decode())
    ws = torch.empty(wsb, dtype=torch.uint8, device=user_rows.device)
    ids = torch.empty((nu, k), dtype=torch.int32, device=user_rows.device)
    sc = torch.empty((nu, k), dtype=torch.float32, device=user_rows.device)
    check(lib.nncf_eval_topk(_ptr(user_rows), nu, _ptr(item_rows), ni, d, int(k), prec, _ptr(ids), _ptr(sc), _ptr(ws), wsb,
                             _stream()))
    return ids, sc


def eval_metrics(topk_ids: torch.Tensor, indptr: torch.Tensor, cols: torch.Tensor):
    """Returns (per_user [n,3] float32, sums [4] float64 = sum AP, sum recall, sum precision, users kept)."""
    _need_cuda(topk_ids, indptr, cols)
    nu, k = topk_ids.shape
    per_user = torch.empty((nu, 3), dtype=torch.float32, device=topk_ids.device)
    sums = torch.zeros(4, dtype=torch.float64, device=topk_ids.device)
    indptr = indptr.to(torch.int64).contiguous()
    cols = _i32(cols)
    check(lib.nncf_eval_metrics(_ptr(topk_ids.contiguous()), nu, k, _ptr(indptr), _ptr(cols), _ptr(per_user), _ptr(sums),
                                _stream()))
    return per_user, sums


def score_pairs(user_table: torch.Tensor, item_table: torch.Tensor, uid: torch.Tensor, cid: torch.Tensor) -> torch.Tensor:
    _need_cuda(user_table, item_table, uid, cid)
    uid, cid = _i32(uid), _i32(cid)
    out = torch.empty(uid.numel(), dtype=torch.float32, device=user_table.device)
    check(lib.nncf_score_pairs(_ptr(user_table), _ptr(item_table), user_table.shape[1], _ptr(uid), _ptr(cid), uid.numel(),
                               _ptr(out), _stream()))
    return out


def eval_given(scores: torch.Tensor, truth: torch.Tensor, indptr: torch.Tensor, topk: int = -1) -> torch.Tensor:
    """per user group (AP@k, AUC, recall@k, precision@k); topk = -1 ranks the whole list (given@-1)"""
    _need_cuda(scores, truth, indptr)
    ng = indptr.numel() - 1
    out = torch.empty((ng, 4), dtype=torch.float32, device=scores.device)
    check(lib.nncf_eval_given(_ptr(scores.contiguous()), _ptr(_i32(truth)), _ptr(indptr.to(torch.int64).contiguous()), ng,
                              int(topk), _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# stand-alone gather / sparse update
# ------------------------------------------------------------------------------------------------
def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    _need_cuda(table, ids)
    ids = _i32(ids)
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    check(lib.nncf_gather_rows(_ptr(table), table.shape[1], _ptr(ids), ids.numel(), _ptr(out), _stream()))
    return out


class SparseUpdater:
    """nncf_updater_*: applies per-position row gradients to an embedding table (SGD or lazy Adam)."""

    def __init__(self, optimizer: str, learn_rate: float, beta1: float = 0.9, beta2: float = 0.999, epsilon: float = 1e-8):
        h = C.c_void_p()
        check(lib.nncf_updater_create(OPTIMIZERS[optimizer], learn_rate, beta1, beta2, epsilon, C.byref(h)))
        self._h = h
        self.optimizer = optimizer

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:
            lib.nncf_updater_destroy(h)
            self._h = None

    def begin_step(self) -> None:
        check(lib.nncf_updater_begin_step(self._h))

    def apply(self, table: torch.Tensor, ids: torch.Tensor, grads: torch.Tensor, m: Optional[torch.Tensor] = None,
              v: Optional[torch.Tensor] = None) -> None:
        """grads [n, d] is clobbered (lazy Adam sums duplicates in place)."""
        _need_cuda(table, ids, grads, m, v)
        ids = _i32(ids)
        assert grads.dtype == torch.float32 and grads.is_contiguous()
        check(lib.nncf_updater_apply(self._h, _ptr(table), _ptr(m), _ptr(v), table.shape[0], table.shape[1], _ptr(ids),
                                     ids.numel(), _ptr(grads), _stream()))
