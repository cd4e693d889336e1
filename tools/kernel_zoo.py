"""Runs every secondary kernel of the hot path once at a realistic size, for `ncu` (profiles/r02_kernels_*):
the alias sampler, the group-shuffle radix passes at 100M links, the GroupSampler, the PAIRS step, lazy Adam, the mean-pool
encoder and given@k.  Prints the algorithmic bytes of each so that the summary can turn durations into GB/s.

  python tools/kernel_zoo.py [scale]        scale < 1 shrinks the sizes (default 1)
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from nncf_b200 import ops
from nncf_b200.ops import FusedStep, StepSpec

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = torch.Generator(device="cuda").manual_seed(0)
alg = {}

# ---- alias sampler: 1M-entry table (degree^0.75), 50M draws: 8 B alias entry read + 4 B written per draw
n_items, n_draws = 1_000_000, int(50_000_000 * scale)
deg = np.random.RandomState(0).zipf(1.5, size=n_items).astype(np.float64)
s = ops.DeviceSampler(deg, 0.75, seed=1)
out = s.sample_device(n_draws)
alg["sample_kernel"] = {"bytes": 12 * n_draws, "what": "%d draws, 8 B alias entry + 4 B out" % n_draws}

# ---- group shuffle at N = 100M links: radix passes move (key 4 B + value 4 B) in and out per pass
N = int(100_000_000 * scale)
train = torch.stack([torch.randint(0, 1_000_000, (N,), device="cuda", generator=g, dtype=torch.int32),
                     torch.randint(0, 1_000_000, (N,), device="cuda", generator=g, dtype=torch.int32),
                     torch.ones(N, device="cuda", dtype=torch.int32)], 1).contiguous()
iidx = torch.randperm(1_000_000, device="cuda", generator=g)
row_perm = torch.randperm(N, device="cuda", generator=g)
chop = 4
block_perm = torch.randperm(N // chop, device="cuda", generator=g)
res = ops.group_shuffle(train, 1, iidx, row_perm, block_perm, chop)
alg["group_shuffle"] = {"bytes_per_pass": 16 * N + 4 * N, "rows": N,
                        "what": "%d links; one 8-bit pass reads keys for the histogram (4 B) and moves key + value (8 B in, 8 B out)" % N}
del train, row_perm, block_perm, res
torch.cuda.synchronize()

# ---- GroupSampler: one epoch of batches, sample_with_negs (B = 512, k = 10, chop 4) on a 2M-link graph
n_links = int(2_000_000 * scale)
rs = np.random.RandomState(1)
tr = np.stack([rs.randint(0, 100_000, n_links), (rs.zipf(1.3, n_links) % 50_000), np.ones(n_links, dtype=np.int64)], 1).astype(np.int32)
gs = ops.DeviceGroupSampler(tr, group_by="item", chop=4, neg_dist="unigram", neg_sign=0, neg_sampling_power=0.75, seed=3)
nb = n_links // 512
o1 = gs.sample(512, nb)
o2, npos = gs.sample_with_negs(512, 10, nb)
alg["group_sample_kernel"] = {"bytes": 24 * nb * 512, "what": "%d batches x 512 rows, ~12 B read + 12 B written per row" % nb}
alg["group_sample_negs_kernel"] = {"bytes": 24 * nb * 512 * 11, "what": "%d batches x 5,632 rows" % nb}

# ---- PAIRS step ('original' / 'group_sample'): B = 512, k = 10, d = 128, R = 37 batches per launch: (1+k)(8 + 16 d) B per positive
R, B, k, d = 37, 512, 10, 128
EU = (torch.rand((1_000_000, d), device="cuda", generator=g) - 0.5) * 0.1
EV = (torch.rand((1_000_000, d), device="cuda", generator=g) - 0.5) * 0.1
rows = (1 + k) * B
uid = torch.randint(0, 1_000_000, (3 * R * rows,), device="cuda", generator=g, dtype=torch.int32)
cid = torch.randint(0, 1_000_000, (3 * R * rows,), device="cuda", generator=g, dtype=torch.int32)
ps = FusedStep(StepSpec(scheme="pairs", loss="skip-gram", precision="fp32", batch_size_p=B, num_negatives=k, dim=d, optimizer="sgd",
                        learn_rate=0.01, replicas=R, neg_loss_weight=8.0, u_reg=1e-6))
ps.run(EU, EV, uid, cid, 3)
alg["pairs_score_kernel"] = {"bytes": R * rows * (8 + 8 * d), "what": "%d pairs: ids + two %d-float rows read" % (R * rows, d)}
alg["pairs_grad_kernel"] = {"bytes": R * rows * (8 + 16 * d), "what": "rows re-read, two gradient rows written"}
alg["rows_sgd_kernel"] = {"bytes": R * rows * (4 + 12 * d) , "what": "per table: gradient row read, table row read-modify-write"}

# ---- lazy Adam on the neg_shared step (the reference's optimizer family): R = 37, B = 512, d = 128
st = [torch.zeros_like(EU), torch.zeros_like(EU), torch.zeros_like(EV), torch.zeros_like(EV)]
ad = FusedStep(StepSpec(scheme="neg_shared", loss="skip-gram", precision="bf16", batch_size_p=B, dim=d, optimizer="lazy_adam",
                        learn_rate=0.001, replicas=R, neg_loss_weight=128.0, u_reg=1e-6))
ad.run(EU, EV, uid, cid, 3, adam_state=st)
alg["adam_apply_vec_kernel"] = {"bytes": 2 * R * B * 7 * 4 * d, "what": "both tables: g, m, v, p read and m, v, p written for %d rows" % (2 * R * B)}
alg["adam_combine_vec_kernel"] = {"bytes": None, "what": "duplicate rows only"}
del st

# ---- mean-pool encoder: 512 unique items x L = 300 word ids, dw = 50 (C1 shape) and a 16x larger case
for n_u, L, dw, V in ((512, 300, 50, 8000), (8192, 300, 128, 100_000)):
    W = torch.randn((V, dw), device="cuda", generator=g)
    content = torch.randint(0, V, (max(n_u, 20000), L), device="cuda", generator=g, dtype=torch.int32)
    ids = torch.randperm(content.shape[0], device="cuda", generator=g)[:n_u].to(torch.int32)
    y = ops.meanpool_fwd(W, content, ids, n_u)
    dW = torch.zeros_like(W)
    ops.meanpool_bwd(dW, content, ids, n_u, torch.randn_like(y))
    alg["meanpool n_u=%d dw=%d" % (n_u, dw)] = {"bytes": n_u * L * (4 + 4 * dw), "what": "L (4 + 4 dw) bytes per unique item, each way"}

# ---- given@k: 1M listed pairs, 20k users
npairs = int(1_000_000 * scale)
pu = torch.sort(torch.randint(0, 20_000, (npairs,), device="cuda", generator=g, dtype=torch.int32))[0]
pc = torch.randint(0, 1_000_000, (npairs,), device="cuda", generator=g, dtype=torch.int32)
sc = ops.score_pairs(EU, EV, pu, pc)
truth = (torch.rand(npairs, device="cuda", generator=g) < 0.2).to(torch.int32)
indptr = torch.zeros(20_001, dtype=torch.int64, device="cuda")
indptr[1:] = torch.cumsum(torch.bincount(pu.long(), minlength=20_000), 0)
ev = ops.eval_given(sc, truth, indptr, 50)
alg["score_pairs_kernel"] = {"bytes": npairs * (8 + 8 * d + 4), "what": "%d pairs" % npairs}
alg["eval_given_kernel"] = {"bytes": npairs * 8, "what": "scores + truth once (each group re-reads its list n times from L1/L2)"}
torch.cuda.synchronize()
print("ZOO_ALGORITHMIC " + json.dumps(alg))
