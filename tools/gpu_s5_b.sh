#!/bin/bash
# session 5, call B: warp-uniform MMA issue (eval g2/g3, score kernel): sweeps, step timings, parity tests
mkdir -p gpurun_out
echo "== sweep"; timeout 300 python tools/eval_sweep.py 37888 1000000 50 all 2>&1 | tee gpurun_out/s5b_sweep.txt | tail -40
echo "== config benches"
( timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 37 2000
  timeout 200 python tools/config_bench.py neg_shared skip-gram 512 128 74 1000
  timeout 200 python tools/config_bench.py neg_shared max-margin 16384 256 1 20 norm
  timeout 200 python tools/config_bench.py neg_shared skip-gram 4096 128 5 200 ) 2>&1 | grep -v Warning | tee gpurun_out/s5b_configs.txt
echo "== tests (train step, eval g2)"; timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_eval_sampler_batch.py -x -q 2>&1 | tail -6
echo "== tests (eval g3)"; NNCF_EVAL_GEN=3 timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -4
