"""GPU parity: the fused training step (CUDA path, through the C-ABI) against the CPU oracle on the same seeded
inputs.  Tolerances are the north_star's: 1e-4 relative for the fp32 mode, 1e-2 for the bf16 tensor-core mode
(relative = max-abs error over the max-abs reference value of the same tensor)."""
import numpy as np
import pytest
import torch

from oracle import nncf_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}
LOSSES = ["skip-gram", "mse", "log-loss", "max-margin"]


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _rel_rows(a, b, floor=0.1):
    """row-wise check: for every row whose reference magnitude is at least `floor` of the tensor's largest, the row's
    max-abs error over the row's own max-abs reference value (a small row next to a large one cannot hide behind it)"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    mag = np.max(np.abs(b), axis=1)
    sel = mag >= floor * max(mag.max(), 1e-30)
    if not np.any(sel):
        return 0.0
    return float(np.max(np.max(np.abs(a - b)[sel], axis=1) / mag[sel]))


def _margin_T(ref, scheme, gamma, cid):
    """M - D of the max-margin losses on the oracle's scores (utils/objectives.py:91-94 / :194-197): an element's gradient
    is the indicator [T > 0], so only elements with |T| within the operand rounding of 0 can come out differently in bf16"""
    S = ref["S"]
    if scheme == "neg_shared":
        D = np.diagonal(S)[None, :] - S
        M = gamma * (1.0 - np.eye(S.shape[0]))
    else:
        _, cid_x = O.unique_first_occurrence(cid)
        D = S[np.arange(S.shape[0]), cid_x][:, None] - S
        M = gamma
    return M - D


def _tables(nu, ni, d, seed, scale=0.5):
    rng = np.random.RandomState(seed)
    EU = rng.uniform(-scale, scale, size=(nu, d)).astype(np.float32)
    EV = rng.uniform(-scale, scale, size=(ni, d)).astype(np.float32)
    return EU, EV


def _params(loss):
    lam = 8.0 if loss == "mse" else 128.0
    gamma = 0.1 if loss == "max-margin" else 10.0
    return lam, gamma


def _scatter(n, ids, rows):
    out = np.zeros((n, rows.shape[1]), dtype=np.float64)
    np.add.at(out, ids, rows.astype(np.float64))
    return out


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("scheme", ["neg_shared", "group_neg_shared"])
@pytest.mark.parametrize("B,d,norm", [(96, 50, False), (200, 64, True), (256, 128, True), (130, 256, True), (384, 192, False)])
def test_matmul_schemes_loss_and_grads(scheme, loss, precision, B, d, norm):
    from nncf_b200.ops import FusedStep, StepSpec
    nu, ni = 300, 120          # few items => duplicate items inside a batch (exercises tf.unique + dup columns)
    EU, EV = _tables(nu, ni, d, seed=B + d)
    if loss == "max-margin" and precision == "bf16" and not norm:
        # the max-margin gradient is an indicator (discontinuous in the score): feed bf16-representable tables so
        # that operand rounding cannot flip indicators that sit within bf16 error of the margin
        EU = torch.from_numpy(EU).bfloat16().float().numpy()
        EV = torch.from_numpy(EV).bfloat16().float().numpy()
    rng = np.random.RandomState(7)
    uid = rng.randint(0, nu, size=B).astype(np.int32)
    cid = rng.randint(0, ni, size=B).astype(np.int32)
    lam, gamma = _params(loss)
    u_reg = 1e-3
    ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, scheme, loss, lam, gamma, u_reg=u_reg,
                        norm_u=norm, norm_v=norm)
    spec = StepSpec(scheme=scheme, loss=loss, precision=precision, batch_size_p=B, dim=d, norm_u=norm, norm_v=norm,
                    optimizer="none", neg_loss_weight=lam, loss_gamma=gamma, u_reg=u_reg)
    step = FusedStep(spec)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    out = step.run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1, want_grads=True)
    torch.cuda.synchronize()
    tol = TOL[precision]
    loss_gpu = float(out["loss"][0].item())
    assert abs(loss_gpu - ref["loss"]) <= tol * max(abs(ref["loss"]), 1e-6), (loss_gpu, ref["loss"])
    flips_allowed = None
    if loss == "max-margin" and precision == "bf16" and norm:
        # indicator gradient + l2-normalised operands (not bf16-representable): a score within bf16 rounding of the margin
        # flips its indicator, which moves the gradient of ITS user row and ITS item row (and the positive's) by a whole
        # 1/(B*n) unit; the loss (continuous) keeps 1e-2.  Instead of a wider tolerance for the whole tensor: count the
        # elements whose margin term is within the rounding of its two scores (unit rows in bf16: each score is good to
        # ~2^-10, so 2^-9 for the pair) of zero, and allow at most that many ROWS (x3: row, column, positive) to miss
        # 1e-2 - by no more than 3e-2.
        T = _margin_T(ref, scheme, gamma, cid)
        flips_allowed = int(np.sum(np.abs(T) < 2.0 ** -9))
        assert flips_allowed <= 0.02 * T.size, flips_allowed          # the test data keep the near-margin set small
    gu = out["grad_user_rows"].cpu().numpy()
    gv = out["grad_item_rows"].cpu().numpy()
    dEU = _scatter(nu, uid, gu)
    if scheme == "neg_shared":
        dEV = _scatter(ni, cid, gv)
    else:
        n_u = int(out["n_unique"].item())
        cid_u, cid_x = O.unique_first_occurrence(cid)
        assert n_u == cid_u.size
        np.testing.assert_array_equal(out["unique_ids"].cpu().numpy()[:n_u], cid_u)     # integer work: bit-exact
        np.testing.assert_array_equal(out["inverse"].cpu().numpy(), cid_x)
        dEV = _scatter(ni, cid_u, gv[:n_u])
    for name, got, want in (("dEU", dEU, ref["dEU"]), ("dEV", dEV, ref["dEV"])):
        if flips_allowed is None:
            assert _rel(got, want) <= tol, (name, _rel(got, want))
            # and row by row, for the rows that carry at least a tenth of the largest gradient
            assert _rel_rows(got, want) <= 3 * tol, (name, "rows", _rel_rows(got, want))
        else:
            row_err = np.max(np.abs(got - want), axis=1) / np.max(np.abs(want))
            assert np.max(row_err) <= 3e-2, (name, float(np.max(row_err)))
            assert int(np.sum(row_err > tol)) <= 3 * flips_allowed, (name, int(np.sum(row_err > tol)), flips_allowed)
    # optimizer='none' must leave the tables untouched
    np.testing.assert_array_equal(tU.cpu().numpy(), EU)
    np.testing.assert_array_equal(tV.cpu().numpy(), EV)


@pytest.mark.parametrize("loss", LOSSES)
@pytest.mark.parametrize("norm", [False, True])
def test_pairs_scheme_loss_and_grads(loss, norm):
    from nncf_b200.ops import FusedStep, StepSpec
    B, k, d, nu, ni = 64, 5, 50, 200, 90
    EU, EV = _tables(nu, ni, d, seed=3)
    rng = np.random.RandomState(11)
    n = (1 + k) * B
    uid = np.concatenate([rng.randint(0, nu, size=B), np.zeros(k * B, dtype=np.int64)]).astype(np.int32)
    uid[B:] = np.repeat(uid[:B], k)                       # 'original': negatives keep the positive's user
    cid = rng.randint(0, ni, size=n).astype(np.int32)
    lam, gamma = _params(loss)
    ref = O.step_mul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, B, k, loss, lam, gamma, u_reg=1e-3,
                     norm_u=norm, norm_v=norm)
    spec = StepSpec(scheme="pairs", loss=loss, precision="fp32", batch_size_p=B, num_negatives=k, dim=d, norm_u=norm,
                    norm_v=norm, optimizer="none", neg_loss_weight=lam, loss_gamma=gamma, u_reg=1e-3)
    out = FusedStep(spec).run(torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda(), torch.from_numpy(uid).cuda(),
                              torch.from_numpy(cid).cuda(), 1, want_grads=True)
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - ref["loss"]) <= 1e-4 * max(abs(ref["loss"]), 1e-6)
    assert _rel(_scatter(nu, uid, out["grad_user_rows"].cpu().numpy()), ref["dEU"]) <= 1e-4
    assert _rel(_scatter(ni, cid, out["grad_item_rows"].cpu().numpy()), ref["dEV"]) <= 1e-4


@pytest.mark.parametrize("scheme", ["neg_shared", "group_neg_shared", "pairs"])
def test_sgd_update_matches_oracle(scheme):
    from nncf_b200.ops import FusedStep, StepSpec
    B, k, d, nu, ni, lr = 128, 3, 64, 150, 60, 0.05
    EU, EV = _tables(nu, ni, d, seed=5)
    rng = np.random.RandomState(2)
    rows = (1 + k) * B if scheme == "pairs" else B
    uid = rng.randint(0, nu, size=rows).astype(np.int32)
    cid = rng.randint(0, ni, size=rows).astype(np.int32)
    if scheme == "pairs":
        ref = O.step_mul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, B, k, "skip-gram", 128.0, 10.0)
    else:
        ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, scheme, "skip-gram", 128.0, 10.0)
    spec = StepSpec(scheme=scheme, loss="skip-gram", precision="fp32", batch_size_p=B, num_negatives=k, dim=d,
                    optimizer="sgd", learn_rate=lr)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1)
    torch.cuda.synchronize()
    assert _rel(tU.cpu().numpy() - EU, -lr * ref["dEU"]) <= 1e-3
    assert _rel(tV.cpu().numpy() - EV, -lr * ref["dEV"]) <= 1e-3


def test_lazy_adam_two_steps_match_oracle():
    from nncf_b200.ops import FusedStep, StepSpec
    B, d, nu, ni, lr = 128, 64, 100, 40, 0.01
    EU, EV = _tables(nu, ni, d, seed=9)
    rng = np.random.RandomState(4)
    uid = rng.randint(0, nu, size=2 * B).astype(np.int32)
    cid = rng.randint(0, ni, size=2 * B).astype(np.int32)
    U, V = EU.astype(np.float64), EV.astype(np.float64)
    mU, vU, mV, vV = np.zeros_like(U), np.zeros_like(U), np.zeros_like(V), np.zeros_like(V)
    for t in range(2):
        u, c = uid[t * B:(t + 1) * B], cid[t * B:(t + 1) * B]
        ref = O.step_matmul(U, V, u, c, "neg_shared", "skip-gram", 128.0, 10.0)
        U, mU, vU = O.lazy_adam_sparse(U, mU, vU, u, ref["dEU"], lr, t + 1)
        V, mV, vV = O.lazy_adam_sparse(V, mV, vV, c, ref["dEV"], lr, t + 1)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision="fp32", batch_size_p=B, dim=d,
                    optimizer="lazy_adam", learn_rate=lr)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    st = [torch.zeros_like(tU), torch.zeros_like(tU), torch.zeros_like(tV), torch.zeros_like(tV)]
    FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 2, adam_state=st)
    torch.cuda.synchronize()
    # Adam's first steps are sign-like (m/sqrt(v)), so compare the updated tables themselves
    assert np.max(np.abs(tU.cpu().numpy() - U)) <= 2e-4
    assert np.max(np.abs(tV.cpu().numpy() - V)) <= 2e-4
    assert _rel(st[0].cpu().numpy(), mU) <= 1e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_replicas_are_independent_batches_on_one_snapshot(precision):
    """R replicas in one call == R separate calls on the same table snapshot, updates summed (synchronous DP)."""
    from nncf_b200.ops import FusedStep, StepSpec
    B, d, nu, ni, R, lr = 128, 64, 500, 400, 3, 0.1
    EU, EV = _tables(nu, ni, d, seed=21)
    rng = np.random.RandomState(8)
    uid = rng.randint(0, nu, size=R * B).astype(np.int32)
    cid = rng.randint(0, ni, size=R * B).astype(np.int32)
    dU = np.zeros((nu, d)); dV = np.zeros((ni, d)); losses = []
    for r in range(R):
        ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid[r * B:(r + 1) * B], cid[r * B:(r + 1) * B],
                            "neg_shared", "skip-gram", 128.0, 10.0)
        dU += ref["dEU"]; dV += ref["dEV"]; losses.append(ref["loss"])
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision=precision, batch_size_p=B, dim=d, optimizer="sgd",
                    learn_rate=lr, replicas=R)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    out = FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1)
    torch.cuda.synchronize()
    tol = TOL[precision]
    np.testing.assert_allclose(out["loss"].cpu().numpy(), np.array(losses), rtol=tol)
    assert _rel(tU.cpu().numpy() - EU, -lr * dU) <= max(tol, 1e-3)
    assert _rel(tV.cpu().numpy() - EV, -lr * dV) <= max(tol, 1e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_reduces_loss_full_size_property(precision):
    """Size-independent property at the benchmark's batch shape (B=512, d=128): repeated steps on the same batch
    must monotonically reduce its loss, and fp32/bf16 must agree on the first loss."""
    from nncf_b200.ops import FusedStep, StepSpec
    B, d, nu, ni = 512, 128, 20000, 20000
    g = torch.Generator(device="cuda").manual_seed(0)
    tU = (torch.rand((nu, d), device="cuda", generator=g) - 0.5) * 0.1
    tV = (torch.rand((ni, d), device="cuda", generator=g) - 0.5) * 0.1
    uid = torch.randint(0, nu, (B,), device="cuda", generator=g, dtype=torch.int32).repeat(8)
    cid = torch.randint(0, ni, (B,), device="cuda", generator=g, dtype=torch.int32).repeat(8)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision=precision, batch_size_p=B, dim=d, optimizer="sgd",
                    learn_rate=5.0)
    out = FusedStep(spec).run(tU, tV, uid, cid, 8)
    losses = out["loss"].cpu().numpy()
    assert np.all(np.isfinite(losses))
    assert np.all(np.diff(losses) < 1e-3) and losses[-1] < losses[0] - 0.05, losses
    expected0 = 128.0 * np.log(2.0) + np.log(2.0)      # all scores ~ 0 at init: (lambda + 1) * log 2
    assert abs(losses[0] - expected0) / expected0 < 2e-2


def test_dense_item_side_for_framework_towers():
    """item_table = NULL: item rows come from a tower; gradient w.r.t. those rows is returned, users updated."""
    from nncf_b200.ops import FusedStep, StepSpec, unique_first_occurrence
    B, d, nu, ni = 128, 50, 100, 30
    EU, EV = _tables(nu, ni, d, seed=13)
    rng = np.random.RandomState(6)
    uid = rng.randint(0, nu, size=B).astype(np.int32)
    cid = rng.randint(0, ni, size=B).astype(np.int32)
    ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, "group_neg_shared", "log-loss", 128.0, 10.0,
                        norm_u=True, norm_v=True)
    tU = torch.from_numpy(EU).cuda()
    tc = torch.from_numpy(cid).cuda()
    uq, inv, nuq = unique_first_occurrence(tc)
    n_u = int(nuq.item())
    rows = torch.from_numpy(EV).cuda()[uq[:n_u].long()]
    spec = StepSpec(scheme="group_neg_shared", loss="log-loss", precision="fp32", batch_size_p=B, dim=d, norm_u=True,
                    norm_v=True, optimizer="none", loss_gamma=10.0)
    out = FusedStep(spec).run(tU, None, torch.from_numpy(uid).cuda(), tc, 1, want_grads=True, item_rows=rows, inverse=inv,
                              n_unique=nuq)
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    cid_u, _ = O.unique_first_occurrence(cid)
    dEV = _scatter(ni, cid_u, out["grad_item_rows"].cpu().numpy()[:n_u])
    assert _rel(dEV, ref["dEV"]) <= 1e-4
    assert _rel(_scatter(nu, uid, out["grad_user_rows"].cpu().numpy()), ref["dEU"]) <= 1e-4


def test_bad_arguments_raise():
    from nncf_b200.ops import FusedStep, StepSpec
    from nncf_b200._lib import NNCFError
    with pytest.raises(NNCFError):
        FusedStep(StepSpec(dim=1000))
    with pytest.raises(AssertionError):
        FusedStep(StepSpec(loss="hinge"))


@pytest.mark.parametrize("replicas,n_steps", [(1, 9), (3, 6), (2, 60)])
def test_host_fed_steps_match_device_fed(replicas, n_steps):
    """nncf_train_steps_host (ids in host memory, copied H2D in chunks of 1, 4, 16, 16, ... steps into a ring of 3 chunk
    buffers, every step's losses back to the host chunk by chunk, into pinned or pageable memory) must produce exactly what the device-fed loop produces on the same ids: same losses, same tables
    (the update is order-independent only up to atomic ordering, so tables are compared at 1e-6).  60 steps = 6 chunks:
    the ring wraps."""
    from nncf_b200.ops import FusedStep, StepSpec
    nu, ni, B, d = 500, 400, 128, 64
    EU, EV = _tables(nu, ni, d, seed=11)
    rng = np.random.RandomState(5)
    n = n_steps * replicas * B
    uid = rng.randint(0, nu, size=n).astype(np.int32)
    cid = rng.randint(0, ni, size=n).astype(np.int32)
    spec = StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="sgd",
                    learn_rate=0.05, replicas=replicas)
    a = FusedStep(spec)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    out = a.run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), n_steps)
    torch.cuda.synchronize()
    b = FusedStep(spec)
    hU, hV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    h_uid, h_cid = torch.from_numpy(uid).pin_memory(), torch.from_numpy(cid).pin_memory()
    loss_h = b.run_host(hU, hV, h_uid, h_cid, n_steps)
    assert not loss_h.is_cuda and loss_h.numel() == n_steps * replicas
    # (two runs of the same steps differ by the order of their float atomics; over 60 dependent steps that grows past 1e-6)
    lt, tt = (1e-5, 1e-6) if n_steps <= 10 else (1e-4, 2e-5)
    assert np.allclose(loss_h.numpy(), out["loss"].cpu().numpy(), rtol=lt, atol=1e-6)
    assert _rel(hU.cpu().numpy(), tU.cpu().numpy()) < tt and _rel(hV.cpu().numpy(), tV.cpu().numpy()) < tt
    # pageable loss array: must return the same numbers
    c = FusedStep(spec)
    cU, cV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    loss_p = c.run_host(cU, cV, h_uid, h_cid, n_steps, torch.zeros(n_steps * replicas, dtype=torch.float32))
    assert np.allclose(loss_p.numpy(), loss_h.numpy(), rtol=lt, atol=1e-6)
    assert _rel(cU.cpu().numpy(), hU.cpu().numpy()) < tt
    # a second call reuses the staging ring; pageable id arrays are accepted too
    loss_h2 = b.run_host(hU, hV, torch.from_numpy(uid), torch.from_numpy(cid), 2)
    assert np.all(np.isfinite(loss_h2.numpy()[:2 * replicas]))
    with pytest.raises(AssertionError):
        b.run_host(hU, hV, h_uid.cuda(), h_cid.cuda(), 1)


@pytest.mark.parametrize("self_gather", ["1", "0"])
@pytest.mark.parametrize("scheme", ["neg_shared", "group_neg_shared"])
@pytest.mark.parametrize("loss", ["skip-gram", "mse"])
@pytest.mark.parametrize("B,d,replicas", [(128, 64, 1), (200, 128, 2), (512, 128, 3), (130, 256, 1), (96, 52, 1)])
def test_fused_sgd_drain_matches_oracle(scheme, loss, B, d, replicas, self_gather, monkeypatch):
    monkeypatch.setenv("NNCF_SELF_GATHER", self_gather)     # neg_shared, dp <= 128: the score kernel gathers its own rows (or not)
    """bf16 + sparse SGD + pointwise loss is the fused mode: the score kernel's drain applies -lr * dX to the table rows
    with one bulk async reduction per row (duplicates, also across replicas, must sum) and publishes the loss itself.
    Checked against the oracle's table delta summed over the replicas, for three consecutive steps' worth of re-use."""
    from nncf_b200.ops import FusedStep, StepSpec
    nu, ni, lr = 300, 90, 0.05           # few items => many duplicate rows
    EU, EV = _tables(nu, ni, d, seed=B + d + replicas)
    rng = np.random.RandomState(B)
    uid = rng.randint(0, nu, size=B * replicas).astype(np.int32)
    cid = rng.randint(0, ni, size=B * replicas).astype(np.int32)
    lam, gamma = _params(loss)
    dU = np.zeros_like(EU, dtype=np.float64); dV = np.zeros_like(EV, dtype=np.float64); losses = []
    for r in range(replicas):
        ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid[r * B:(r + 1) * B], cid[r * B:(r + 1) * B],
                            scheme, loss, lam, gamma)
        dU += ref["dEU"]; dV += ref["dEV"]; losses.append(ref["loss"])
    spec = StepSpec(scheme=scheme, loss=loss, precision="bf16", batch_size_p=B, dim=d, optimizer="sgd", learn_rate=lr,
                    replicas=replicas, neg_loss_weight=lam, loss_gamma=gamma)
    step = FusedStep(spec)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    out = step.run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1)
    torch.cuda.synchronize()
    got = out["loss"].cpu().numpy()
    assert np.allclose(got, losses, rtol=1e-2), (got, losses)
    assert _rel(tU.cpu().numpy() - EU, -lr * dU) <= 1e-2
    assert _rel(tV.cpu().numpy() - EV, -lr * dV) <= 1e-2
    # the in-kernel loss hand-off re-arms itself: a second step on the same handle reports a fresh loss, not a sum
    out2 = step.run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1)
    torch.cuda.synchronize()
    assert np.all(out2["loss"].cpu().numpy() < 1.5 * got + 1.0) and np.all(np.isfinite(out2["loss"].cpu().numpy()))


@pytest.mark.parametrize("B,d,replicas,steps", [(512, 128, 37, 6), (512, 128, 1, 8), (300, 64, 5, 4)])
def test_self_gather_equals_separate_gather_over_consecutive_steps(B, d, replicas, steps, monkeypatch):
    """the self-gathering score kernel (no gather launch; image blocks handed between CTAs through flags; drains held back
    until every CTA has read its rows) against the two-launch step on the same ids, several dependent steps in one call:
    the same losses and the same tables up to the bf16 tolerance (the order of the fp32 reductions at the L2 is free, and a
    row that differs in its last fp32 bits can round to another bf16 operand in the next step: two runs of the SAME mode
    differ by ~2e-3 of the largest update after six steps on these hot items)"""
    from nncf_b200.ops import FusedStep, StepSpec
    nu, ni, lr = 4000, 1500, 0.05
    EU, EV = _tables(nu, ni, d, seed=B + replicas)
    rng = np.random.RandomState(steps)
    n = B * replicas * steps
    uid = torch.from_numpy(rng.randint(0, nu, size=n).astype(np.int32)).cuda()
    cid = torch.from_numpy((rng.zipf(1.3, size=n) % ni).astype(np.int32)).cuda()       # hot items: duplicates inside and across steps
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NNCF_SELF_GATHER", mode)
        step = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="sgd",
                                  learn_rate=lr, replicas=replicas, neg_loss_weight=128.0, loss_gamma=10.0))
        tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
        out = step.run(tU, tV, uid, cid, steps)
        torch.cuda.synchronize()
        res[mode] = (out["loss"].cpu().numpy().copy(), tU.cpu().numpy(), tV.cpu().numpy())
    assert np.all(np.isfinite(res["1"][0])) and res["1"][0].shape == res["0"][0].shape
    assert np.allclose(res["1"][0], res["0"][0], rtol=2e-3), (res["1"][0][:4], res["0"][0][:4])
    assert _rel(res["1"][1] - EU, res["0"][1] - EU) <= 1e-2
    assert _rel(res["1"][2] - EV, res["0"][2] - EV) <= 1e-2


@pytest.mark.parametrize("loss,norm", [("max-margin", True), ("skip-gram", False)])
def test_c5_full_size_loss_and_update_match_torch_fp32(loss, norm):
    """BASELINE config 5 at its full size (batch 16,384, dim 256): the oracle would need minutes, so the check is an
    independent fp32 torch evaluation of the same formulas on the GPU (score matrix materialised: 1 GiB) — loss within
    the bf16 tolerance, and the SGD update of a sample of rows against autograd."""
    from nncf_b200.ops import FusedStep, StepSpec
    B, d, nu, ni = 16384, 256, 50000, 40000
    lam, gamma, lr = 128.0, 0.1, 1.0
    g = torch.Generator(device="cuda").manual_seed(5)
    tU = (torch.rand((nu, d), device="cuda", generator=g) - 0.5)
    tV = (torch.rand((ni, d), device="cuda", generator=g) - 0.5)
    if not norm:
        tU *= 0.1; tV *= 0.1
    uid = torch.randperm(nu, device="cuda", generator=g)[:B].to(torch.int32)        # distinct ids: rows map 1:1 to updates
    cid = torch.randperm(ni, device="cuda", generator=g)[:B].to(torch.int32)
    U0 = tU[uid.long()].clone().requires_grad_(True)
    V0 = tV[cid.long()].clone().requires_grad_(True)
    Uh = torch.nn.functional.normalize(U0, dim=1, eps=1e-12) if norm else U0
    Vh = torch.nn.functional.normalize(V0, dim=1, eps=1e-12) if norm else V0
    S = Uh @ Vh.T
    if loss == "max-margin":                                                        # utils/objectives.py:91-94
        D = torch.diagonal(S)[None, :] - S
        M = gamma * (1.0 - torch.eye(B, device="cuda"))
        L = torch.relu(M - D).mean()
    else:                                                                           # :99-105
        w = lam / (B - 1)
        L = (-(torch.nn.functional.logsigmoid(torch.diagonal(S))).sum() - w * (torch.nn.functional.logsigmoid(-S).sum()
             - torch.nn.functional.logsigmoid(-torch.diagonal(S)).sum())) / B
    L.backward()
    EU_before, EV_before = tU.clone(), tV.clone()
    spec = StepSpec(scheme="neg_shared", loss=loss, precision="bf16", batch_size_p=B, dim=d, norm_u=norm, norm_v=norm,
                    optimizer="sgd", learn_rate=lr, neg_loss_weight=lam, loss_gamma=gamma)
    out = FusedStep(spec).run(tU, tV, uid, cid, 1)
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - float(L)) <= 1e-2 * abs(float(L)), (float(out["loss"][0]), float(L))
    gU = (EU_before[uid.long()] - tU[uid.long()]) / lr
    gV = (EV_before[cid.long()] - tV[cid.long()]) / lr
    if loss == "skip-gram":      # (max-margin gradients are indicators: bf16 operand rounding flips entries near the margin)
        assert float((gU - U0.grad).abs().max() / U0.grad.abs().max()) <= 2e-2
        assert float((gV - V0.grad).abs().max() / V0.grad.abs().max()) <= 2e-2
    else:
        cos = torch.nn.functional.cosine_similarity(gU.flatten(), U0.grad.flatten(), dim=0)
        assert float(cos) > 0.98, float(cos)
    # rows outside the batch are untouched
    mask = torch.ones(nu, dtype=torch.bool, device="cuda"); mask[uid.long()] = False
    assert torch.equal(tU[mask], EU_before[mask])


def test_zero_steps_is_a_no_op():
    from nncf_b200.ops import FusedStep, StepSpec
    tU = torch.rand((100, 64), device="cuda"); tV = torch.rand((100, 64), device="cuda")
    a, b = tU.clone(), tV.clone()
    ids = torch.zeros(128, dtype=torch.int32, device="cuda")
    out = FusedStep(StepSpec(batch_size_p=128, dim=64)).run(tU, tV, ids, ids, 0)
    torch.cuda.synchronize()
    assert out["loss"].numel() == 0 and torch.equal(tU, a) and torch.equal(tV, b)


def _augment(EU, EV, ub, cb):
    """tables with the two interaction-bias columns: users (ubias, 1), items (1, cbias)"""
    one_u, one_v = np.ones((EU.shape[0], 1), np.float32), np.ones((EV.shape[0], 1), np.float32)
    return (np.concatenate([EU, ub[:, None].astype(np.float32), one_u], 1),
            np.concatenate([EV, one_v, cb[:, None].astype(np.float32)], 1))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("scheme,loss,norm,bias", [("neg_shared", "skip-gram", False, "both"), ("neg_shared", "log-loss", True, "item"),
                                                   ("group_neg_shared", "mse", False, "user"), ("group_neg_shared", "max-margin", True, "both"),
                                                   ("pairs", "skip-gram", True, "item"), ("pairs", "log-loss", False, "both")])
def test_interaction_bias_matches_oracle(scheme, loss, norm, bias, precision):
    """InteractionDot(bias=...) (ref: modules/interaction/interaction_dot.py:96-107) through the augmented-column tables:
    loss, embedding gradients and bias gradients against the oracle; constant columns and unused biases stay put."""
    from nncf_b200.ops import FusedStep, StepSpec
    if scheme == "pairs" and precision == "bf16":
        pytest.skip("the PAIRS scheme is fp32 only")
    rng = np.random.RandomState(3)
    nu, ni, d, B, k, lr = 200, 90, 50, 96, 3, 2.0
    EU, EV = _tables(nu, ni, d, seed=1)
    ub = rng.uniform(-0.3, 0.3, nu) * (bias in ("user", "both"))
    cb = rng.uniform(-0.3, 0.3, ni) * (bias in ("item", "both"))
    if loss == "max-margin" and precision == "bf16":
        EU = torch.from_numpy(EU).bfloat16().float().numpy(); EV = torch.from_numpy(EV).bfloat16().float().numpy()
        ub = torch.from_numpy(ub).bfloat16().double().numpy(); cb = torch.from_numpy(cb).bfloat16().double().numpy()
    rows = (1 + k) * B if scheme == "pairs" else B
    uid = rng.randint(0, nu, rows).astype(np.int32); cid = rng.randint(0, ni, rows).astype(np.int32)
    lam, gamma = _params(loss)
    kw = dict(u_reg=1e-3, norm_u=norm, norm_v=norm, ubias=ub if bias in ("user", "both") else None,
              cbias=cb if bias in ("item", "both") else None)
    if scheme == "pairs":
        ref = O.step_mul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, B, k, loss, lam, gamma, **kw)
    else:
        ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, scheme, loss, lam, gamma, **kw)
    AU, AV = _augment(EU, EV, ub, cb)
    spec = StepSpec(scheme=scheme, loss=loss, precision=precision, batch_size_p=B, num_negatives=k, dim=d + 2, norm_u=norm,
                    norm_v=norm, optimizer="sgd", learn_rate=lr, neg_loss_weight=lam, loss_gamma=gamma, u_reg=1e-3,
                    interaction_bias=bias)
    tU, tV = torch.from_numpy(AU).cuda(), torch.from_numpy(AV).cuda()
    out = FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1)
    torch.cuda.synchronize()
    tol = TOL[precision]
    assert abs(float(out["loss"][0]) - ref["loss"]) <= tol * abs(ref["loss"])
    gU = (AU.astype(np.float64) - tU.cpu().numpy()) / lr
    gV = (AV.astype(np.float64) - tV.cpu().numpy()) / lr
    if not (loss == "max-margin" and precision == "bf16"):
        assert _rel(gU[:, :d], ref["dEU"]) <= max(tol, 1e-3)
        assert _rel(gV[:, :d], ref["dEV"]) <= max(tol, 1e-3)
        scale = np.max(np.abs(ref["dEU"]))

        def bias_grad_ok(got, want):
            # pairwise losses compare scores that share the bias (neg_shared: same item column, group: same user row), so
            # that bias has an exactly zero gradient: check it absolutely, against the scale of the embedding gradients
            if np.max(np.abs(want)) < 1e-9 * scale:
                return np.max(np.abs(got)) <= max(tol, 1e-3) * scale
            return _rel(got, want) <= max(tol, 1e-3)
        if bias in ("user", "both"):
            assert bias_grad_ok(gU[:, d], ref["dubias"])
        if bias in ("item", "both"):
            assert bias_grad_ok(gV[:, d + 1], ref["dcbias"])
    else:
        # max-margin + bf16 + l2-normalised rows: indicator gradients; rows may miss 1e-2 only where an element's margin term
        # sits within the operand rounding of zero (counted on the oracle's scores), and then by at most 3e-2
        T = _margin_T(ref, scheme, gamma, cid)
        flips_allowed = int(np.sum(np.abs(T) < 2.0 ** -9))
        for got, want in ((gU[:, :d], ref["dEU"]), (gV[:, :d], ref["dEV"])):
            row_err = np.max(np.abs(got - want), axis=1) / np.max(np.abs(want))
            assert np.max(row_err) <= 3e-2, float(np.max(row_err))
            assert int(np.sum(row_err > tol)) <= 3 * flips_allowed, (int(np.sum(row_err > tol)), flips_allowed)
    # the constant columns never move; an unused bias column stays zero
    assert np.array_equal(tU.cpu().numpy()[:, d + 1], np.ones(nu, np.float32)) and np.array_equal(tV.cpu().numpy()[:, d], np.ones(ni, np.float32))
    if bias == "item":
        assert np.all(tU.cpu().numpy()[:, d] == 0)
    if bias == "user":
        assert np.all(tV.cpu().numpy()[:, d + 1] == 0)


@pytest.mark.parametrize("split", ["1", "2", "4", "8"])
@pytest.mark.parametrize("optimizer,u_reg", [("sgd", 0.0), ("sgd", 1.0), ("lazy_adam", 1.0)])   # u_reg = 1: the regulariser's gradient is as large as the loss's
@pytest.mark.parametrize("scheme", ["neg_shared", "group_neg_shared"])
def test_split_sweep_and_folded_regulariser(scheme, optimizer, u_reg, split, monkeypatch):
    """Small steps (R = 1, the reference's sequential loop) split every owner block's sweep over several CTAs that ADD their
    partial gradient blocks (bulk reductions into the table in the fused SGD mode, red.global.add into the zeroed blocks for
    lazy Adam); the activity regulariser (ref: utils/utilities.py:129-135) is folded into the two-launch step: loss term
    in the gather, gradient in the drain.  Any split must reproduce the oracle: loss, table update / Adam first moment."""
    from nncf_b200.ops import FusedStep, StepSpec
    monkeypatch.setenv("NNCF_SPLIT", split)
    B, d, nu, ni, lr = 512, 128, 700, 260, 0.05
    EU, EV = _tables(nu, ni, d, seed=31)
    rng = np.random.RandomState(17)
    uid = rng.randint(0, nu, size=B).astype(np.int32)
    cid = rng.randint(0, ni, size=B).astype(np.int32)
    ref = O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid, cid, scheme, "skip-gram", 128.0, 10.0, u_reg=u_reg)
    spec = StepSpec(scheme=scheme, loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer=optimizer,
                    learn_rate=lr, neg_loss_weight=128.0, loss_gamma=10.0, u_reg=u_reg)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    st = [torch.zeros_like(tU), torch.zeros_like(tU), torch.zeros_like(tV), torch.zeros_like(tV)] if optimizer == "lazy_adam" else None
    out = FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), 1, adam_state=st)
    torch.cuda.synchronize()
    assert abs(float(out["loss"][0]) - ref["loss"]) <= 1e-2 * abs(ref["loss"])
    if optimizer == "sgd":
        assert _rel(tU.cpu().numpy() - EU, -lr * ref["dEU"]) <= 1e-2
        assert _rel(tV.cpu().numpy() - EV, -lr * ref["dEV"]) <= 1e-2
    else:
        # first Adam step: m = (1 - beta1) * g on the touched rows (duplicates summed)
        assert _rel(st[0].cpu().numpy(), 0.1 * ref["dEU"]) <= 1e-2
        assert _rel(st[2].cpu().numpy(), 0.1 * ref["dEV"]) <= 1e-2


@pytest.mark.parametrize("fold", ["1", "0"])
@pytest.mark.parametrize("scheme", ["neg_shared", "group_neg_shared"])
def test_lazy_adam_folded_matches_oracle_over_steps(scheme, fold, monkeypatch):
    """bf16 lazy Adam, three dependent steps of R = 3 replicas with many duplicate ids (inside a batch, across replicas,
    across steps): the folded form (the score kernel's drain sums gradient rows into per-table accumulators keyed by id,
    one apply launch claims each id once) and the owner / combine / apply form against the oracle's _apply_sparse rule
    with duplicates summed (ref: utils/optimizer.py:108-134)."""
    from nncf_b200.ops import FusedStep, StepSpec
    monkeypatch.setenv("NNCF_ADAM_FOLD", fold)
    B, d, nu, ni, lr, R, steps = 256, 64, 300, 90, 0.01, 3, 3
    EU, EV = _tables(nu, ni, d, seed=21)
    rng = np.random.RandomState(8)
    uid = rng.randint(0, nu, size=steps * R * B).astype(np.int32)
    cid = rng.randint(0, ni, size=steps * R * B).astype(np.int32)
    U, V = EU.astype(np.float64), EV.astype(np.float64)
    mU, vU, mV, vV = np.zeros_like(U), np.zeros_like(U), np.zeros_like(V), np.zeros_like(V)
    for t in range(steps):
        dU = np.zeros_like(U); dV = np.zeros_like(V)
        us, cs = [], []
        for r in range(R):
            sl = slice((t * R + r) * B, (t * R + r + 1) * B)
            ref = O.step_matmul(U, V, uid[sl], cid[sl], scheme, "skip-gram", 128.0, 10.0, u_reg=1e-3)
            dU += ref["dEU"]; dV += ref["dEV"]; us.append(uid[sl]); cs.append(cid[sl])
        U, mU, vU = O.lazy_adam_sparse(U, mU, vU, np.concatenate(us), dU, lr, t + 1)
        V, mV, vV = O.lazy_adam_sparse(V, mV, vV, np.concatenate(cs), dV, lr, t + 1)
    spec = StepSpec(scheme=scheme, loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="lazy_adam", learn_rate=lr,
                    replicas=R, neg_loss_weight=128.0, loss_gamma=10.0, u_reg=1e-3)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    st = [torch.zeros_like(tU), torch.zeros_like(tU), torch.zeros_like(tV), torch.zeros_like(tV)]
    out = FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), steps, adam_state=st)
    torch.cuda.synchronize()
    assert np.all(np.isfinite(out["loss"].cpu().numpy()))
    # Adam's first steps move every touched coordinate by ~lr whatever the gradient's size: compare the tables absolutely
    # (3 steps x lr = 0.03 of travel; bf16 gradients may flip a near-zero coordinate's direction for one step)
    assert np.mean(np.abs(tU.cpu().numpy() - U)) <= 2e-4 and np.max(np.abs(tU.cpu().numpy() - U)) <= 2.5 * lr
    assert np.mean(np.abs(tV.cpu().numpy() - V)) <= 2e-4 and np.max(np.abs(tV.cpu().numpy() - V)) <= 2.5 * lr
    assert _rel(st[0].cpu().numpy(), mU) <= 2e-2 and _rel(st[2].cpu().numpy(), mV) <= 2e-2
    assert _rel(st[1].cpu().numpy(), vU) <= 3e-2 and _rel(st[3].cpu().numpy(), vV) <= 3e-2
    # untouched rows did not move
    untouched = np.setdiff1d(np.arange(nu), uid)
    if untouched.size:
        np.testing.assert_array_equal(tU.cpu().numpy()[untouched], EU[untouched])


@pytest.mark.parametrize("gx", ["1", "0"])
@pytest.mark.parametrize("loss", ["skip-gram", "mse"])
@pytest.mark.parametrize("B,d,replicas,norm", [(128, 64, 1, False), (256, 128, 1, True), (256, 64, 2, False), (512, 128, 3, False), (640, 100, 2, False),
                                              (1024, 128, 2, False), (512, 32, 40, False)])
def test_g_exchange_two_sided_matches_oracle(loss, B, d, replicas, norm, gx, monkeypatch):
    """G' exchange (score_tc.cuh): the user-side CTAs compute every sigmoid once and hand the bf16 gradient tiles to the
    item-side CTAs, which only contract (dV_b += G'(a, b)^T U_a).  Both modes (NNCF_GX=0: the item side recomputes S^T)
    against the oracle: loss and both gradients of every replica with optimizer='none', then the fused sparse-SGD step
    (table deltas summed over the replicas) over two consecutive steps on one handle (the flags carry a sequence number)."""
    monkeypatch.setenv("NNCF_GX", gx)
    from nncf_b200.ops import FusedStep, StepSpec
    nu, ni, lr = 700, 260, 0.05
    EU, EV = _tables(nu, ni, d, seed=B + d + replicas)
    rng = np.random.RandomState(B + 1)
    uid = rng.randint(0, nu, size=B * replicas).astype(np.int32)
    cid = rng.randint(0, ni, size=B * replicas).astype(np.int32)
    lam, gamma = _params(loss)
    refs = [O.step_matmul(EU.astype(np.float64), EV.astype(np.float64), uid[r * B:(r + 1) * B], cid[r * B:(r + 1) * B],
                          "neg_shared", loss, lam, gamma, norm_u=norm, norm_v=norm) for r in range(replicas)]
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    d_uid, d_cid = torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda()
    if replicas == 1:
        spec = StepSpec(scheme="neg_shared", loss=loss, precision="bf16", batch_size_p=B, dim=d, norm_u=norm, norm_v=norm,
                        optimizer="none", neg_loss_weight=lam, loss_gamma=gamma)
        out = FusedStep(spec).run(tU, tV, d_uid, d_cid, 1, want_grads=True)
        torch.cuda.synchronize()
        assert abs(float(out["loss"][0]) - refs[0]["loss"]) <= 1e-2 * abs(refs[0]["loss"])
        for name, got, want in (("dEU", _scatter(nu, uid, out["grad_user_rows"].cpu().numpy()), refs[0]["dEU"]),
                                ("dEV", _scatter(ni, cid, out["grad_item_rows"].cpu().numpy()), refs[0]["dEV"])):
            assert _rel(got, want) <= 1e-2, (name, _rel(got, want))
            assert _rel_rows(got, want) <= 3e-2, (name, "rows", _rel_rows(got, want))
    if norm:
        return                                   # (the fused drain is the un-normalised pointwise step)
    spec = StepSpec(scheme="neg_shared", loss=loss, precision="bf16", batch_size_p=B, dim=d, optimizer="sgd", learn_rate=lr,
                    replicas=replicas, neg_loss_weight=lam, loss_gamma=gamma)
    step = FusedStep(spec)
    out = step.run(tU, tV, d_uid, d_cid, 1)
    torch.cuda.synchronize()
    assert np.allclose(out["loss"].cpu().numpy(), [r["loss"] for r in refs], rtol=1e-2)
    dU, dV = sum(r["dEU"] for r in refs), sum(r["dEV"] for r in refs)
    U1, V1 = tU.cpu().numpy().astype(np.float64), tV.cpu().numpy().astype(np.float64)
    assert _rel(U1 - EU, -lr * dU) <= 1e-2, _rel(U1 - EU, -lr * dU)
    assert _rel(V1 - EV, -lr * dV) <= 1e-2, _rel(V1 - EV, -lr * dV)
    assert _rel_rows(V1 - EV, -lr * dV) <= 3e-2
    # second step on the same handle, from the updated tables
    refs2 = [O.step_matmul(U1, V1, uid[r * B:(r + 1) * B], cid[r * B:(r + 1) * B], "neg_shared", loss, lam, gamma) for r in range(replicas)]
    out2 = step.run(tU, tV, d_uid, d_cid, 1)
    torch.cuda.synchronize()
    assert np.allclose(out2["loss"].cpu().numpy(), [r["loss"] for r in refs2], rtol=1e-2)
    dU2, dV2 = sum(r["dEU"] for r in refs2), sum(r["dEV"] for r in refs2)
    assert _rel(tU.cpu().numpy() - U1, -lr * dU2) <= 1e-2
    assert _rel(tV.cpu().numpy() - V1, -lr * dV2) <= 1e-2


@pytest.mark.parametrize("fold", ["1", "0"])
@pytest.mark.parametrize("scheme,loss,precision", [("group_neg_shared", "log-loss", "bf16"), ("neg_shared", "max-margin", "bf16"),
                                                   ("neg_shared", "skip-gram", "fp32")])
def test_lazy_adam_behind_the_finalize_pass_matches_oracle(scheme, loss, precision, fold, monkeypatch):
    """Lazy Adam on the steps that need a finalize pass (l2-normalised rows, pairwise losses; fp32 precision): folded form (the
    finalize kernel adds its finished rows into the per-table accumulators, one apply launch) and owner / combine / apply
    form against the oracle over three dependent steps of R = 2 replicas with duplicate ids everywhere."""
    from nncf_b200.ops import FusedStep, StepSpec
    monkeypatch.setenv("NNCF_ADAM_FOLD", fold)
    B, d, nu, ni, lr, R, steps = 256, 64, 300, 90, 0.01, 2, 3
    norm = loss != "skip-gram"
    lam, gamma = _params(loss)
    EU, EV = _tables(nu, ni, d, seed=33)
    rng = np.random.RandomState(9)
    uid = rng.randint(0, nu, size=steps * R * B).astype(np.int32)
    cid = rng.randint(0, ni, size=steps * R * B).astype(np.int32)
    U, V = EU.astype(np.float64), EV.astype(np.float64)
    mU, vU, mV, vV = np.zeros_like(U), np.zeros_like(U), np.zeros_like(V), np.zeros_like(V)
    for t in range(steps):
        dU = np.zeros_like(U); dV = np.zeros_like(V)
        us, cs = [], []
        for r in range(R):
            sl = slice((t * R + r) * B, (t * R + r + 1) * B)
            ref = O.step_matmul(U, V, uid[sl], cid[sl], scheme, loss, lam, gamma, u_reg=1e-3, norm_u=norm, norm_v=norm)
            dU += ref["dEU"]; dV += ref["dEV"]; us.append(uid[sl]); cs.append(cid[sl])
        U, mU, vU = O.lazy_adam_sparse(U, mU, vU, np.concatenate(us), dU, lr, t + 1)
        V, mV, vV = O.lazy_adam_sparse(V, mV, vV, np.concatenate(cs), dV, lr, t + 1)
    spec = StepSpec(scheme=scheme, loss=loss, precision=precision, batch_size_p=B, dim=d, optimizer="lazy_adam", learn_rate=lr,
                    replicas=R, neg_loss_weight=lam, loss_gamma=gamma, u_reg=1e-3, norm_u=norm, norm_v=norm)
    tU, tV = torch.from_numpy(EU).cuda(), torch.from_numpy(EV).cuda()
    st = [torch.zeros_like(tU), torch.zeros_like(tU), torch.zeros_like(tV), torch.zeros_like(tV)]
    out = FusedStep(spec).run(tU, tV, torch.from_numpy(uid).cuda(), torch.from_numpy(cid).cuda(), steps, adam_state=st)
    torch.cuda.synchronize()
    assert np.all(np.isfinite(out["loss"].cpu().numpy()))
    # (Adam's first steps move every touched coordinate by ~lr whatever the gradient's size: tables compared absolutely; a
    #  max-margin indicator or a near-zero coordinate may take one step the other way in bf16)
    assert np.mean(np.abs(tU.cpu().numpy() - U)) <= 3e-4 and np.max(np.abs(tU.cpu().numpy() - U)) <= 2.5 * lr
    assert np.mean(np.abs(tV.cpu().numpy() - V)) <= 3e-4 and np.max(np.abs(tV.cpu().numpy() - V)) <= 2.5 * lr
    tol_m = 1e-3 if precision == "fp32" else 3e-2
    assert _rel(st[0].cpu().numpy(), mU) <= tol_m and _rel(st[2].cpu().numpy(), mV) <= tol_m
    untouched = np.setdiff1d(np.arange(nu), uid)
    if untouched.size:
        np.testing.assert_array_equal(tU.cpu().numpy()[untouched], EU[untouched])
