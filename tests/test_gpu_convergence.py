"""GPU test of what the throughput modes do to CONVERGENCE: the reference's loop is strictly
sequential (ref: models/train_neg_shared.py:40-58); R > 1 replicas per step apply R stale-snapshot updates at once, and the
stratified multi-GPU schedule draws a batch's shared negatives from one item stratum.  Same planted C1-shaped problem,
same epochs, whole@50 on held-out links (tools/convergence.py; the full table is profiles/r02_convergence.md).  The
assertions are the caveats DESIGN.md 8.3 states next to the throughput numbers, with room for seed noise (+-0.004)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.fixture(scope="module")
def problem():
    import convergence as cv
    train, test, nu, ni = cv.planted_links()
    return cv, train, cv.csr_truth(test[0], test[1], nu, ni), nu, ni


def test_replicas_need_a_scaled_learning_rate_and_then_match_the_sequential_loop(problem):
    cv, train, truth, nu, ni = problem
    d, B, lr = 64, 512, 0.01
    seq = cv.run_single(train, truth, nu, ni, 1, 20, "lazy_adam", lr, d, B, 1)                    # R = 1: the reference's loop
    r8_scaled = cv.run_single(train, truth, nu, ni, 8, 20, "lazy_adam", lr * 8 ** 0.5, d, B, 1)
    r8_plain = cv.run_single(train, truth, nu, ni, 8, 20, "lazy_adam", lr, d, B, 1)
    r37_scaled = cv.run_single(train, truth, nu, ni, 37, 20, "lazy_adam", lr * 37 ** 0.5, d, B, 1)
    rec = lambda h: h[-1][1]                                                                        # noqa: E731
    assert rec(seq) > 0.05, rec(seq)                                   # the problem is learnable in 20 epochs (0.076 measured)
    assert rec(r8_scaled) >= 0.85 * rec(seq), (rec(r8_scaled), rec(seq))        # 0.096 vs 0.076 measured
    assert rec(r37_scaled) >= 0.85 * rec(seq), (rec(r37_scaled), rec(seq))      # 0.085 vs 0.076 measured
    assert rec(r8_plain) < 0.6 * rec(r8_scaled), (rec(r8_plain), rec(r8_scaled))   # summed stale updates at the R = 1 rate lag: 0.021
    assert all(torch.isfinite(torch.tensor(h[-1][0])) for h in (seq, r8_scaled, r8_plain, r37_scaled))


def test_stratified_schedule_converges_more_slowly_on_a_small_catalogue(problem):
    """N = 2 ranks as objects of this process (LocalPeerGroup: the schedule, not the transport, changes the statistics):
    4 strata of a 17k-item catalogue cost recall at equal epochs (0.060 vs 0.093 after 40), but it trains."""
    cv, train, truth, nu, ni = problem
    d, B, lr = 64, 512, 0.01
    seq = cv.run_single(train, truth, nu, ni, 1, 40, "lazy_adam", lr, d, B, 1)
    strat = cv.run_stratified(train, truth, nu, ni, 2, 1, 40, "lazy_adam", lr, d, B, 1)
    assert seq[-1][1] > 0.07, seq[-1][1]
    assert 0.45 * seq[-1][1] < strat[-1][1] < 1.05 * seq[-1][1], (strat[-1][1], seq[-1][1])
    assert strat[-1][0] < strat[0][0]                                  # the loss falls
