"""GPU parity for the batch sources and views added after the scoped four: the device GroupSampler (statistically,
against histograms produced by the reference's own class), sampled_neg_shared and the label-driven 'original' loss
(against the oracle), the presample / sampled_neg_shared batch assembly (bit-exact), and the trainers end to end."""
import numpy as np
import pytest
import torch

from oracle import nncf_oracle as O
from test_oracle_sampling import G, META, check_group_sampler_histograms, group_sampler_histograms

pytestmark = pytest.mark.gpu
LOSSES = ["skip-gram", "mse", "log-loss", "max-margin"]


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


# ------------------------------------------------------------------------------------------------ GroupSampler
@pytest.mark.parametrize("ci", range(len(META["gs_cases"])))
def test_group_sampler_matches_reference_distribution(ci):
    from nncf_b200.data_utils import GroupSampler
    c = META["gs_cases"][ci]
    train = G["gs_train"]
    n_users, n_items = int(train[:, 0].max()) + 1, int(train[:, 1].max()) + 1
    gs = GroupSampler(train, group_by="item", chop=c["chop"], neg_dist=c["neg_dist"], neg_sign=-1, seed=5 + ci)
    nb = 3000
    h = group_sampler_histograms(lambda B, n: gs.sample_device(B, n).cpu().numpy(),
                                 lambda B, k, n: gs.sample_with_negs_device(B, k, n)[0].cpu().numpy(), c, n_items, n_users, nb)
    gs._dev.check()
    check_group_sampler_histograms(h, ci, nb)


def test_group_sampler_structure_and_api():
    from nncf_b200.data_utils import GroupSampler
    train = G["gs_train"]
    links = set(map(tuple, train[:, :2]))
    gs = GroupSampler(train, group_by="item", chop=4, neg_dist="unigram", neg_sign=0, seed=1)
    b = gs.sample(30)                                      # host API, like the reference: int array [B, 3]
    assert b.shape == (30, 3) and np.all(b[:, 2] == 1)
    assert all((u, i) in links for u, i, _ in b)
    assert all(len(set(b[s:s + 4, 1])) == 1 for s in range(0, 28, 4))        # runs of `chop` rows share the group
    assert gs.sample(30, strict_return_shape=False).shape == (28, 3)
    bn = gs.sample_with_negs(30, 3)
    assert bn.shape == (120, 3)
    npos = int((bn[:, 2] == 1).sum())
    assert npos == 32 and np.all(bn[:npos, 2] == 1) and np.all(bn[npos:, 2] == 0)      # ceil(30/4)*4 positives first
    assert all((u, i) in links for u, i, y in bn if y == 1)
    # negatives of a group follow the positives' group order: k * chop = 12 per group for unigram
    assert np.array_equal(bn[npos:npos + 12, 1], np.repeat(bn[0, 1], 12))
    # two calls draw different batches; the same seed replays the same stream
    assert not np.array_equal(gs.sample(30), gs.sample(30))
    g1 = GroupSampler(train, chop=4, seed=9); g2 = GroupSampler(train, chop=4, seed=9)
    assert np.array_equal(g1.sample(64), g2.sample(64))
    # group_by user swaps the columns
    gu = GroupSampler(train, group_by="user", chop=2, seed=3)
    bu = gu.sample(16)
    assert all((u, i) in links for u, i, _ in bu) and all(bu[s, 0] == bu[s + 1, 0] for s in range(0, 16, 2))
    with pytest.raises(AssertionError):
        GroupSampler(train, group_by="nobody")


# ------------------------------------------------------------------------------------------------ sampled_neg_shared
@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("B,k,d,norm", [(64, 10, 50, False), (130, 7, 128, True), (40, 20, 256, True)])
def test_sampled_neg_shared_step_matches_oracle(loss, B, k, d, norm):
    from nncf_b200.ops import FusedStep, StepSpec
    rng = np.random.RandomState(B + k)
    nu, ni = 200, 90
    EU = rng.uniform(-0.5, 0.5, size=(nu, d)).astype(np.float32)
    EV = rng.uniform(-0.5, 0.5, size=(ni, d)).astype(np.float32)
    uid = np.r_[rng.randint(0, nu, B), np.zeros(k, dtype=np.int64)].astype(np.int32)
    cid = rng.randint(0, ni, B + k).astype(np.int32)
    lam, gamma = (8.0 if loss == "mse" else 128.0), (0.1 if loss == "max-margin" else 10.0)
    u_reg, lr = 1e-3, 2.0          # large step: (new - old) / lr must resolve gradients ~1e-3 against fp32 table values ~0.5
    ref = O.step_sampled_neg_shared(EU.astype(np.float64), EV.astype(np.float64), uid, cid, B, k, loss, lam, gamma, u_reg, norm, norm)
    step = FusedStep(StepSpec(scheme="sampled_neg_shared", loss=loss, precision="fp32", batch_size_p=B, num_negatives=k,
                              dim=d, norm_u=norm, norm_v=norm, optimizer="sgd", learn_rate=lr, neg_loss_weight=lam,
                              loss_gamma=gamma, u_reg=u_reg))
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    out = step.run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1, want_grads=True)
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert _rel(out["grad_user_rows"].cpu().numpy(), ref["dU_rows"]) <= 1e-4
    assert _rel(out["grad_item_rows"].cpu().numpy(), ref["dV_rows"]) <= 1e-4
    assert _rel((tU.cpu().numpy().astype(np.float64) - EU) / -lr, ref["dEU"]) <= 5e-4      # sparse SGD applied, duplicates summed
    assert _rel((tV.cpu().numpy().astype(np.float64) - EV) / -lr, ref["dEV"]) <= 5e-4


# ------------------------------------------------------------------------------------------------ labels in PAIRS
@pytest.mark.parametrize("loss", ["skip-gram", "mse"])
def test_pairs_step_follows_response_labels(loss):
    from nncf_b200.ops import FusedStep, StepSpec
    rng = np.random.RandomState(11)
    nu, ni, d, B, k = 150, 80, 64, 48, 3
    n = (1 + k) * B
    EU = rng.uniform(-0.5, 0.5, size=(nu, d)).astype(np.float32)
    EV = rng.uniform(-0.5, 0.5, size=(ni, d)).astype(np.float32)
    uid = rng.randint(0, nu, n).astype(np.int32); cid = rng.randint(0, ni, n).astype(np.int32)
    y = np.full(n, -1 if loss == "skip-gram" else 0, dtype=np.int32); y[:B] = 1
    perm = rng.permutation(n)
    y = y[perm]                                                  # positives anywhere in the batch (presample 'random')
    lam = 8.0 if loss == "mse" else 128.0
    ref = O.step_mul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, B, k, loss, lam, 10.0, u_reg=1e-3, y_true=y)
    step = FusedStep(StepSpec(scheme="pairs", loss=loss, precision="fp32", batch_size_p=B, num_negatives=k, dim=d,
                              optimizer="sgd", learn_rate=2.0, neg_loss_weight=lam, u_reg=1e-3))
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    out = step.run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1, responses=torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    assert _rel((tU.cpu().numpy().astype(np.float64) - EU) / -2.0, ref["dEU"]) <= 5e-4
    assert _rel((tV.cpu().numpy().astype(np.float64) - EV) / -2.0, ref["dEV"]) <= 5e-4


# ------------------------------------------------------------------------------------------------ batch assembly
@pytest.mark.parametrize("layout,neg_col", [(0, 1), (0, 0), (1, 1)])
def test_presample_assemble_bit_exact(layout, neg_col):
    from nncf_b200 import ops
    rng = np.random.RandomState(2)
    n, k = 1003, 5
    tp = np.stack([rng.randint(0, 99, n), rng.randint(0, 77, n), np.ones(n, dtype=np.int64)], 1).astype(np.int32)
    negs = rng.randint(0, 50, n * k).astype(np.int32)
    got = ops.presample_assemble(torch.from_numpy(tp).cuda(), k, torch.from_numpy(negs).cuda(), neg_col, -1, layout).cpu().numpy()
    assert np.array_equal(got, O.presample_rows(tp, k, negs, neg_col, -1, layout))


def test_assemble_sns_batches_bit_exact():
    from nncf_b200 import ops
    rng = np.random.RandomState(4)
    nb, B, k = 7, 33, 4
    train = np.stack([rng.randint(1, 99, nb * B + 5), rng.randint(0, 77, nb * B + 5), np.ones(nb * B + 5, dtype=np.int64)], 1).astype(np.int32)
    negs = rng.randint(0, 77, nb * k).astype(np.int32)
    uid, cid = ops.assemble_sns_batches(torch.from_numpy(train).cuda(), nb, B, k, torch.from_numpy(negs).cuda())
    ref = np.vstack([O.assemble_sns_batch(train[b * B:(b + 1) * B], k, negs[b * k:(b + 1) * k]) for b in range(nb)])
    assert np.array_equal(uid.cpu().numpy(), ref[:, 0]) and np.array_equal(cid.cpu().numpy(), ref[:, 1])


# ------------------------------------------------------------------------------------------------ trainers end to end
def _run(scheme, loss, extra=None):
    from nncf_b200.main import run
    pd = {'reset_after_getconf': True, 'max_epoch': 2, 'loss': loss, 'batch_size_p': 128, 'num_negatives': 3,
          'learn_rate': 0.01, 'neg_loss_weight': 8 if loss == 'mse' else 128, 'loss_gamma': 0.1 if loss == 'max-margin' else 10,
          'user_dim': 32, 'item_dim': 32, 'word_dim': 32, 'chop_size': 4, 'neg_sampling_power': 1}
    pd.update(extra or {})
    np.random.seed(0)
    return run(['--data_name', 'synthetic_small', '--model_choice', 'mf', '--conf_choice', 'best',
                '--train_scheme', scheme, '--eval_scheme', 'whole@10', '--param_dict', repr(pd)])


@pytest.mark.parametrize("scheme,loss,extra", [
    ("presample", "skip-gram", {"shuffle_st": "original"}),
    ("presample", "mse", {"shuffle_st": "reverse"}),
    ("presample", "skip-gram", {"shuffle_st": "random"}),
    ("presample", "skip-gram", {"shuffle_st": "by_user"}),
    ("presample", "mse", {"shuffle_st": "by_item_chop"}),
    ("presample", "skip-gram", {"shuffle_st": "by_useritem_chop"}),
    ("reverse", "skip-gram", None),
    ("sampled_neg_shared", "skip-gram", None),
    ("sampled_neg_shared", "log-loss", None),
    ("group_neg_shared", "log-loss", {"group_shuffling_trick": False}),
    ("group_sample", "skip-gram", {"group_shuffling_trick": False}),
    ("group_sample", "mse", {"group_shuffling_trick": False, "neg_dist": "uniform"}),
])
def test_widened_trainers_run_and_log(scheme, loss, extra, capsys):
    tr = _run(scheme, loss, extra)
    out = capsys.readouterr().out
    assert 'epoch 0 (0 it) cost -1.00000' in out
    assert 'epoch 2 (' in out and 'train recall/map' in out and 'test recall/map' in out
    assert len(tr.train_time) == 2


def test_presample_pairwise_guard():
    with pytest.raises(AssertionError):
        _run("presample", "log-loss", {"shuffle_st": "by_item"})
    with pytest.raises(AssertionError):
        _run("reverse", "max-margin")


def test_sampled_neg_shared_optimises(capsys):
    """the per-epoch mean cost the trainer prints falls steadily (the step itself is checked against the oracle above)"""
    import re
    _run('sampled_neg_shared', 'skip-gram', {'max_epoch': 5, 'precision': 'fp32', 'learn_rate': 0.05, 'num_negatives': 20})
    costs = [float(x) for x in re.findall(r"epoch [1-9]\d* \(\d+ it\) cost ([0-9.]+)", capsys.readouterr().out)]
    assert len(costs) == 5 and all(b < a for a, b in zip(costs, costs[1:])) and costs[-1] < 0.5 * costs[0], costs
