#!/bin/bash
mkdir -p gpurun_out
echo "== tests (eval g2)"; timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -2
echo "== tests (eval g3)"; NNCF_EVAL_GEN=3 timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -2
echo "== sweep"; timeout 300 python tools/eval_sweep.py 37888 1000000 50 quick2 2>&1 | grep -v Warn | tee gpurun_out/s5n_sweep.txt | tail -40
