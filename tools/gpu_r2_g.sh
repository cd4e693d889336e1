#!/bin/bash
# round 2, call G: convergence table, eval (k = 100 in the pair kernel) parity + timing, bench short, ncu of the secondary kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval_sampler_batch.py tests/test_gpu_train_step.py -x -q 2>&1 | tail -3
for k in 50 100; do timeout 300 python tools/eval_bench.py 75776 2000000 $k 2>&1 | tail -2; done
NNCF_EVAL_GEN=2 timeout 300 python tools/eval_bench.py 75776 2000000 100 2>&1 | tail -1
timeout 900 python tools/convergence.py --epochs 12 --out gpurun_out/r02_convergence.md > gpurun_out/r02g_convergence.log 2>&1; tail -45 gpurun_out/r02g_convergence.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eval --cpu-steps 2 > gpurun_out/r02g_n1_short.json 2>/dev/null
python - <<PY
import json
j=json.load(open("gpurun_out/r02g_n1_short.json")); print("short: value=%.3e ms/step=%.4f e2e=%.3e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]), j["clocks"])
PY
KR='regex:sample_kernel|make_keys|radix_|scan_|group_emit|group_sample|pairs_|rows_sgd|adam_|meanpool_|score_pairs|eval_given'
timeout 1500 ncu --set full --clock-control none -k "$KR" -c 80 -o gpurun_out/r02_kernels python tools/kernel_zoo.py 1 > gpurun_out/r02_kernels.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_kernels.log
ls -la gpurun_out/r02_kernels.ncu-rep
