#!/bin/bash
# round 2, call J: full GPU test suite (vectorised PAIRS, mean-pool rewrite, folded Adam, tightened checks), timings, convergence v3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<PY
import sys, torch, time
sys.path.insert(0, '.')
from nncf_b200 import ops
from nncf_b200.ops import FusedStep, StepSpec
g = torch.Generator(device="cuda").manual_seed(0)
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
R, B, k, d = 37, 512, 10, 128
EU = (torch.rand((1_000_000, d), device="cuda", generator=g) - 0.5) * 0.1
EV = (torch.rand((1_000_000, d), device="cuda", generator=g) - 0.5) * 0.1
rows = (1 + k) * B
uid = torch.randint(0, 1_000_000, (R * rows,), device="cuda", generator=g, dtype=torch.int32)
cid = torch.randint(0, 1_000_000, (R * rows,), device="cuda", generator=g, dtype=torch.int32)
ps = FusedStep(StepSpec(scheme="pairs", loss="skip-gram", precision="fp32", batch_size_p=B, num_negatives=k, dim=d, optimizer="sgd", learn_rate=0.01, replicas=R, neg_loss_weight=8.0, u_reg=1e-6))
us = t(lambda: ps.run(EU, EV, uid, cid, 1))
print("PAIRS step R=37 B=512 k=10 d=128: %.1f us per step = %.3e positive links/s, %.0f GB/s at (1+k)(8+16d) B per positive" % (us, R * B / us * 1e6, R * B * (1 + k) * (8 + 16 * d) / us * 1e6 / 1e9))
for n_u, L, dw, V in ((512, 300, 50, 8000), (8192, 300, 128, 100_000)):
    W = torch.randn((V, dw), device="cuda", generator=g)
    content = torch.randint(0, V, (max(n_u, 20000), L), device="cuda", generator=g, dtype=torch.int32)
    ids = torch.randperm(content.shape[0], device="cuda", generator=g)[:n_u].to(torch.int32)
    y = ops.meanpool_fwd(W, content, ids, n_u)
    dW = torch.zeros_like(W); gy = torch.randn_like(y)
    f = t(lambda: ops.meanpool_fwd(W, content, ids, n_u)); b = t(lambda: ops.meanpool_bwd(dW, content, ids, n_u, gy))
    by = n_u * L * (4 + 4 * dw)
    print("meanpool n_u=%d L=%d dw=%d: fwd %.1f us (%.0f GB/s), bwd %.1f us (%.0f GB/s) at L(4+4dw) B per item" % (n_u, L, dw, f, by / f * 1e6 / 1e9, b, by / b * 1e6 / 1e9))
PY
timeout 1500 python tools/convergence.py --epochs 40 --out gpurun_out/r02_convergence.md > gpurun_out/r02j_convergence.log 2>&1; tail -22 gpurun_out/r02j_convergence.log
