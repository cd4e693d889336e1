#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./build/score_bench 512 1 > gpurun_out/score_bench_512_1.txt
timeout 60 ./build/score_bench 2048 1 > gpurun_out/score_bench_2048_1.txt
# launch list of a short bench run (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 60 --warmup 10 --no-eval --cpu-steps 5 --links 4000000 > gpurun_out/ncu_bench.log 2>&1
# full capture of the two dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_grad_tc -s 20 -c 2 -o gpurun_out/prof_score \
   python bench.py --steps 30 --warmup 10 --no-eval --cpu-steps 5 --links 4000000 > gpurun_out/ncu_score.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/prof_eval \
   python - > gpurun_out/ncu_eval.log 2>&1 <<PY
import torch, sys
sys.path.insert(0, '.')
from nncf_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
U = torch.randn((32768, 128), device="cuda", generator=g) / 128 ** 0.5
V = torch.randn((500000, 128), device="cuda", generator=g) / 128 ** 0.5
ops.eval_topk(U, V, 50, "bf16"); torch.cuda.synchronize()
PY
ls -la gpurun_out/ | tail; cat gpurun_out/score_bench_2048_1.txt
