#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/prof_eval2 \
   python - > gpurun_out/ncu_eval2.log 2>&1 <<PY
import torch, sys
sys.path.insert(0, '.')
from nncf_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
U = torch.randn((37888, 128), device="cuda", generator=g) / 128 ** 0.5     # 296 user blocks = 2 waves of 148 CTAs
V = torch.randn((1000000, 128), device="cuda", generator=g) / 128 ** 0.5
ops.eval_topk(U, V, 50, "bf16"); torch.cuda.synchronize()
PY
tail -3 gpurun_out/ncu_eval2.log
