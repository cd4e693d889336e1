#!/bin/bash
# session 5, call C: per-lane filter (filter_tile3): parity tests for both generations, sweep
mkdir -p gpurun_out
echo "== tests (eval g2)"; timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -4
echo "== tests (eval g3)"; NNCF_EVAL_GEN=3 timeout 400 python -m pytest tests/test_gpu_eval_sampler_batch.py -k "topk or whole_eval" -x -q 2>&1 | tail -4
echo "== sweep"; timeout 300 python tools/eval_sweep.py 37888 1000000 50 all 2>&1 | tee gpurun_out/s5c_sweep.txt | tail -40
